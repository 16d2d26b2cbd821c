#!/bin/bash
# One 8-GPU call: the strong-scaling curve of BASELINE config 3 at N = 8, 4, 2 (per-rank processes, torchrun), the
# u128 variant and config 5 at N = 8, ONE handle over 8 / 4 devices in one process, and the 8-device parity test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
nvidia-smi topo -m > gpurun_out/${TAG}_topo_n8.txt 2>&1; (nproc; free -g | head -2; numactl -H 2>/dev/null | head -6) >> gpurun_out/${TAG}_topo_n8.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k one_handle > gpurun_out/${TAG}_tmulti_n8.log 2>&1; echo "8-device handle test rc=$?"; tail -3 gpurun_out/${TAG}_tmulti_n8.log
P=29711
tr() { n=$1; name=$2; shift 2; P=$((P+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $n --steps 5 --warmup 3 "$@" > gpurun_out/${TAG}_mg_$name.json 2> gpurun_out/${TAG}_mg_$name.err; echo "$name rc=$?"; }
tr 8 n8_default
tr 8 n8_keys --exchange keys --no-e2e --no-check
tr 8 n8_k63 --workload c3k63 --no-e2e
tr 8 n8_c5 --workload c5 --steps 2 --warmup 1 --no-e2e --no-check
tr 4 n4_default
tr 4 n4_keys --exchange keys --no-e2e --no-check
tr 2 n2_default
timeout 600 python scripts/bench_multi_handle.py --devices 0,1,2,3,4,5,6,7 --workload c3 --export > gpurun_out/${TAG}_mh_c3_n8.json 2> gpurun_out/${TAG}_mh_c3_n8.err; echo "mh n=8 rc=$?"
timeout 600 python scripts/bench_multi_handle.py --devices 0,1 --workload c3 --trace > gpurun_out/${TAG}_mh_c3_n2.json 2> gpurun_out/${TAG}_mh_c3_n2_trace.err; echo "mh n=2 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_mg_n[248]_*.json")) + sorted(glob.glob("gpurun_out/${TAG}_mh_c3_n[28].json")):
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
