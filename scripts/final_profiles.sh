#!/bin/bash
# One gpurun call: the round's bench line, the reference arm, the ncu launch list of the bench
# command and a `--set full` capture of its big kernels (B200_PROFILING.md recipe).  usage: final_profiles.sh <tag>
cd "$(dirname "$0")/.."
T=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-probe --no-e2e"
$CMD > gpurun_out/plain_$T.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv $CMD > gpurun_out/ncu_l_$T.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_$T.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:pack_flat|check_reads|scatter_reads|scatter_buckets|update_pages' \
    -s 15 -c 5 -o gpurun_out/prof_$T -f $CMD > gpurun_out/ncu_f_$T.log 2>&1
echo "full rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$T.json").read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["kernel"])
print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
r = json.loads(open("gpurun_out/bench_ref_$T.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_optimistic"]["value"], r["cpu_optimistic"]["cores"])
PY
