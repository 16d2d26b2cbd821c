#!/bin/bash
# round 2, run 2: fused record scatter, spill unification, u128 geometries, C3 bin balance
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "superkmer or heavy or paged or fused or host_batcher or two_rank" > gpurun_out/r02_t2.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_t2.log
B="timeout 300 python bench.py --steps 5 --warmup 3"
Q="--no-cpu --no-probe --no-consumer --no-check --no-e2e"
run() { name=$1; shift; $B $Q "$@" > gpurun_out/r02_b2_$name.json 2>> gpurun_out/r02_b2.err; echo "$name rc=$?"; }
run c3k63_default --workload c3k63
run c3k63_l2v5 --workload c3k63 --opt l2s_variant=5
run c3k63_per8 --workload c3k63 --opt l1_per=8
run c3k63_pt640 --workload c3k63 --opt l2s_variant=5 --opt page_threads=640
run c3_default --workload c3
run c3_sub25 --workload c3 --sub-log2 25
run c3_sub24 --workload c3 --sub-log2 24
run c3_l2v4 --workload c3 --opt l2s_variant=4
run c3_l2v0 --workload c3 --opt l2s_variant=0
run c2_pt640 --opt page_threads=640
run c2_per4 --opt l1_per=4
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_b2_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r02_b2_")[1], round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms", d["table"]["sub_tables"],
              {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1})
    except Exception as e:
        print(f, "no result", e)
PY
