#!/bin/bash
# Multi-GPU measurements of one gpurun call: both fused exchanges at every N this box has.
# usage: scripts/run_multi.sh "<list of N>" [c5]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=29511
for N in $1; do
  for X in keys skm; do
    P=$((P+1))
    KTG_EXCHANGE=$X timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port $P bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/mg_n${N}_$X.json 2> gpurun_out/mg_n${N}_$X.err
    echo "N=$N $X rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/mg_n${N}_$X.json").read().strip().splitlines()[-1])
    print("  ", d["config"].get("exchange"), round(d["value"] / 1e9, 1), "G/s", round(d["ms_per_step"], 2), "ms  e2e",
          d["e2e"] and round(d["e2e"]["value"] / 1e9, 1), {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items()})
except Exception as e:
    print("   no result", e)
PY
  done
done
if [ "$2" = "c5" ]; then
  for X in ${C5_MODES:-fused}; do
    P=$((P+1))
    KTG_EXCHANGE=$X timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port $P bench.py --gpus 8 --workload c5 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/mg_c5_$X.json 2> gpurun_out/mg_c5_$X.err
    echo "C5 $X rc=$?"; tail -c 1500 gpurun_out/mg_c5_$X.json
  done
fi
