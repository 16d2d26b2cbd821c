#!/bin/bash
# round 2, run 4: sharded handle + externals + C++ host on shards; u128 geometry probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "shard or cpp_host or heavy or externals or superkmer or fused" > gpurun_out/r02_t4.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/r02_t4.log
B="timeout 300 python bench.py --steps 5 --warmup 3"
Q="--no-cpu --no-probe --no-consumer --no-check --no-e2e"
run() { name=$1; shift; $B $Q "$@" > gpurun_out/r02_b4_$name.json 2>> gpurun_out/r02_b4.err; echo "$name rc=$?"; }
run c2_default
run c3k63_sub27 --workload c3k63 --sub-log2 27
run c3k63_sub27_v3 --workload c3k63 --sub-log2 27 --opt l2s_variant=3
run c3k63_sub28 --workload c3k63 --sub-log2 28
run c3_sub27 --workload c3 --sub-log2 27
timeout 300 python scripts/bench_multi_handle.py --devices 0,0 --workload c2 > gpurun_out/r02_mh_c2_00.json 2> gpurun_out/r02_mh.err; echo "mh rc=$?"; cat gpurun_out/r02_mh_c2_00.json
timeout 300 python scripts/bench_multi_handle.py --devices 0 --workload c2 --export > gpurun_out/r02_mh_c2_0.json 2>> gpurun_out/r02_mh.err; echo "mh rc=$?"; cat gpurun_out/r02_mh_c2_0.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_b4_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02_b4_")[1], round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms", d["table"]["sub_tables"],
              {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1})
    except Exception as e:
        print(f, "no result", e)
PY
