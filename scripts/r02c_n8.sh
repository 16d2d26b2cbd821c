#!/bin/bash
# N GPUs (default 8): BASELINE config 3 (strong scaling) through the default exchange and the direct exchange
# (big-tile sender).  usage: scripts/r02c_n8.sh [tag] [N]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02c}; N=${2:-8}
P=29811
tr() { name=$1; shift; P=$((P+1))
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --steps 4 --warmup 2 --no-cpu "$@" > gpurun_out/${TAG}_n${N}_$name.json 2> gpurun_out/${TAG}_n${N}_$name.err; echo "$name rc=$?"; }
tr direct --exchange direct
tr default --no-e2e
tr direct_k63 --exchange direct --workload c3k63 --no-e2e
