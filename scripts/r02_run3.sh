#!/bin/bash
# round 2, run 3: one handle over several shards (emulated on one GPU), scan fast path with two bins per thread
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "shard or cpp_host or heavy or superkmer or fused or paged" > gpurun_out/r02_t3.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_t3.log
B="timeout 300 python bench.py --steps 5 --warmup 3"
Q="--no-cpu --no-probe --no-consumer --no-check --no-e2e"
run() { name=$1; shift; $B $Q "$@" > gpurun_out/r02_b3_$name.json 2>> gpurun_out/r02_b3.err; echo "$name rc=$?"; }
run c2_default
run c3_default --workload c3
run c3_l2v1 --workload c3 --opt l2s_variant=1
run c3_sub25 --workload c3 --sub-log2 25
run c3k63_default --workload c3k63
run c3k63_sub26 --workload c3k63 --sub-log2 26
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_b3_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02_b3_")[1], round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms", d["table"]["sub_tables"],
              {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1})
    except Exception as e:
        print(f, "no result", e)
PY
