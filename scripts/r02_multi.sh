#!/bin/bash
# Multi-GPU measurements of one gpurun call (N = number of GPUs of the box):
#   the 2-GPU tests, the torchrun bench at every power of two up to N (C3 strong scaling, both exchanges at
#   N = 2 and 4), and ONE handle over 2..N devices in one process (scripts/bench_multi_handle.py).
# usage: scripts/r02_multi.sh <N> [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; TAG=${2:-r02}
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/${TAG}_tmulti_n$N.log 2>&1; echo "multi tests rc=$?"; tail -5 gpurun_out/${TAG}_tmulti_n$N.log
P=29611
tr() { n=$1; name=$2; shift 2; P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/${TAG}_mg_$name.json 2> gpurun_out/${TAG}_mg_$name.err; echo "$name rc=$?"; }
for n in 2 4 8; do
  [ $n -le $N ] || continue
  tr $n n${n}_default
  if [ $n -le 4 ]; then tr $n n${n}_skm --exchange skm --no-e2e --no-check; tr $n n${n}_keys --exchange keys --no-e2e --no-check; fi
  tr $n n${n}_k63 --workload c3k63 --no-e2e
  ids=$(seq -s, 0 $((n-1)))
  timeout 600 python scripts/bench_multi_handle.py --devices $ids --workload c3 --export > gpurun_out/${TAG}_mh_c3_n$n.json 2> gpurun_out/${TAG}_mh_c3_n$n.err; echo "mh n=$n rc=$?"
done
timeout 300 python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --no-probe --no-consumer > gpurun_out/${TAG}_mg_n1_c3.json 2> gpurun_out/${TAG}_mg_n1_c3.err
timeout 300 python scripts/bench_multi_handle.py --devices 0 --workload c3 --export > gpurun_out/${TAG}_mh_c3_n1.json 2> gpurun_out/${TAG}_mh_c3_n1.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_mg_*.json")) + sorted(glob.glob("gpurun_out/${TAG}_mh_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "kernels" in d:
            print(f.split("/")[-1], d["config"].get("exchange"), round(d["value"] / 1e9, 1), "G/s", round(d["ms_per_step"], 2), "ms  e2e",
                  d["e2e"] and round(d["e2e"]["value"] / 1e9, 1), "check", d["digest_check"] and (d["digest_check"]["equal"], d["digest_check"]["golden_equal"]),
                  {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if v["ms_per_step"] > 0.1})
        else:
            print(f.split("/")[-1], d["devices"], round(d["value"] / 1e9, 1), "G/s", round(d["ms_per_step"], 2), "ms golden", d["golden_equal"], "export ms", d.get("export_ms"))
    except Exception as e:
        print(f, "no result", e)
PY
