#!/bin/bash
# round 2, run 1: parity of the SoA / bulk-copy refactor, then the page kernel variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t1.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_t1.log
B="timeout 300 python bench.py --steps 5 --warmup 3"
$B > gpurun_out/r02_b1_c2.json 2> gpurun_out/r02_b1_c2.err; echo "c2 rc=$?"
Q="--no-cpu --no-probe --no-consumer --no-check --no-e2e"
$B $Q --opt page_nbuf=2 --opt page_threads=1024 > gpurun_out/r02_b1_c2_nbuf2_1024.json 2>> gpurun_out/r02_b1.err; echo "rc=$?"
$B $Q --opt page_nbuf=2 --opt page_threads=768 > gpurun_out/r02_b1_c2_nbuf2_768.json 2>> gpurun_out/r02_b1.err; echo "rc=$?"
$B $Q --opt page_threads=640 > gpurun_out/r02_b1_c2_640.json 2>> gpurun_out/r02_b1.err; echo "rc=$?"
$B --workload c3 --no-cpu --no-probe --no-consumer > gpurun_out/r02_b1_c3.json 2> gpurun_out/r02_b1_c3.err; echo "c3 rc=$?"
$B --workload c3k63 --no-cpu --no-probe --no-consumer > gpurun_out/r02_b1_c3k63.json 2> gpurun_out/r02_b1_c3k63.err; echo "c3k63 rc=$?"
$B --workload c3k63 $Q --opt page_nbuf=2 --opt page_threads=1024 > gpurun_out/r02_b1_c3k63_nbuf2.json 2>> gpurun_out/r02_b1.err; echo "rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_b1_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"] / 1e9, 2), "G/s", round(d["ms_per_step"], 3), "ms", "e2e", d["e2e"] and round(d["e2e"]["value"] / 1e9, 2),
              "check", d["digest_check"] and (d["digest_check"]["equal"], d["digest_check"]["golden_equal"]),
              {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
    except Exception as e:
        print(f, "no result", e)
PY
