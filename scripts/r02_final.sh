#!/bin/bash
# The round's final measurements on one GPU: the whole -m gpu suite, smoke(), the bench line with all its legs,
# the reference arm (bounded sample, and ALL of C2 once on one core, in the background), C3 at k = 31 / 63.
# usage: scripts/r02_final.sh [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02}
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/${T}_gpu_box.txt 2>&1
nproc >> gpurun_out/${T}_gpu_box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/${T}_gpu_box.txt
python bench.py --impl reference --full > gpurun_out/${T}_reference_full_c2.json 2> gpurun_out/${T}_reference_full_c2.err &
REF=$!
python -m pytest tests -m gpu -q > gpurun_out/${T}_tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${T}_tests_gpu.log
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${T}_smoke.log
wait $REF; echo "reference full rc=$?"
python bench.py > gpurun_out/${T}_bench_c2_final.json 2> gpurun_out/${T}_bench_c2_final.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/${T}_bench_c2_reference_arm.json 2> gpurun_out/${T}_bench_c2_reference_arm.err; echo "reference arm rc=$?"
python bench.py --workload c3 --no-cpu --no-probe --no-consumer > gpurun_out/${T}_bench_c3_final.json 2> gpurun_out/${T}_bench_c3_final.err; echo "c3 rc=$?"
python bench.py --workload c3k63 --no-cpu --no-probe --no-consumer > gpurun_out/${T}_bench_c3k63_final.json 2> gpurun_out/${T}_bench_c3k63_final.err; echo "c3k63 rc=$?"
