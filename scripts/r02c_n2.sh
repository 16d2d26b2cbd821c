#!/bin/bash
# N = 2: BASELINE config 3 (strong scaling) through the key exchange and the direct exchange (big-tile sender),
# k = 31 and k = 63, and the 2-GPU tests.  usage: scripts/r02c_n2.sh [tag] [quick]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02c}
P=29711
tr() { name=$1; shift; P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus 2 --steps 4 --warmup 2 --no-cpu "$@" > gpurun_out/${TAG}_n2_$name.json 2> gpurun_out/${TAG}_n2_$name.err; echo "$name rc=$?"; }
tr direct --exchange direct --no-e2e
tr direct_k63 --exchange direct --workload c3k63 --no-e2e
[ "$2" = quick ] && exit 0
tr keys --exchange keys --no-e2e
tr keys_k63 --exchange keys --workload c3k63 --no-e2e
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_tmulti_n2.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/${TAG}_tmulti_n2.log
