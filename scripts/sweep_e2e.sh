#!/bin/bash
# e2e sweep of the host batcher's knobs on C2, one GPU: "<eager 0|1> <flush pcts> [chunk MB] [staging buffers]" per run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {
  local tag="e${1}_f${2}_c${3:-64}_b${4:-4}"
  local extra=""
  [ "$1" = "0" ] && extra="KTG_NO_EAGER=1"
  env $extra KTG_FLUSH_PCT=$2 KTG_CHUNK_MB=${3:-64} KTG_STAGE_BUFS=${4:-4} timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-probe \
    > gpurun_out/e2e_$tag.json 2> gpurun_out/e2e_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/e2e_$tag.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print("$tag: e2e", round(e["ms_per_step"], 2), "ms", round(e["value"] / 1e9, 2), "G/s  copy-only",
          round(e["h2d_copy_only_ms"], 2), " kernel-only", round(d["ms_per_step"], 2), "ms ",
          {k: round(v["ms_per_step"], 2) for k, v in e["kernels"].items()})
except Exception as ex:
    print("$tag: no result", ex)
PY
}
for cfg in "$@"; do run $cfg; done
