#!/bin/bash
# One gpurun call: the ncu launch list of the bench command and `--set full` captures of its big kernels
# (B200_PROFILING.md recipe), for C2 (u64) and the first reads of C3 k=63 (u128).  usage: r02_profiles.sh <tag>
cd "$(dirname "$0")/.."
T=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-probe --no-e2e --no-check --no-consumer"
$CMD > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c2.csv $CMD > gpurun_out/${T}_ncu_l.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:pack_flat|scatter_reads|scatter_buckets|update_pages' \
    -s 8 -c 4 -o gpurun_out/${T}_prof_c2 -f $CMD > gpurun_out/${T}_ncu_f.log 2>&1
echo "full c2 rc=$?"
CMD2="python bench.py --workload c3k63 --steps 1 --warmup 1 --no-cpu --no-probe --no-e2e --no-check --no-consumer"
ncu --set full --clock-control none --import-source on -k 'regex:scatter_reads|scatter_buckets|update_pages' \
    -s 3 -c 3 -o gpurun_out/${T}_prof_c3k63 -f $CMD2 > gpurun_out/${T}_ncu_f2.log 2>&1
echo "full c3k63 rc=$?"
ls -la gpurun_out/${T}_prof_*.ncu-rep
