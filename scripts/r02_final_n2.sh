#!/bin/bash
# Final N = 2 checks: the 2-GPU tests, BASELINE config 3 through bench.py's default exchange (direct below 4 ranks)
# with the end-to-end leg, and ONE handle over two devices with and without the direct exchange.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r02}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${T}_tmulti_n2.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/${T}_tmulti_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_scale_n2.json 2> gpurun_out/${T}_scale_n2.err; echo "bench n2 rc=$?"
timeout 300 python scripts/bench_multi_handle.py --devices 0,1 --workload c3 > gpurun_out/${T}_mh_c3_n2.json 2> gpurun_out/${T}_mh_c3_n2.err; echo "mh rc=$?"
timeout 300 python scripts/bench_multi_handle.py --devices 0,1 --workload c3 --opt mg_direct=1 > gpurun_out/${T}_mh_c3_n2_direct.json 2> gpurun_out/${T}_mh_c3_n2_direct.err; echo "mh direct rc=$?"
