#!/usr/bin/env python3
"""Host timeline (option `trace`) of one end-to-end build from pinned host memory: trace_e2e.py [c2|c3|c3k63] [--no-trace] [--no-profile]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from katome_b200 import GpuGIR, synth_reads_device
from katome_b200.workloads import BY_NAME

wl = BY_NAME[next((a for a in sys.argv[1:] if not a.startswith("-")), "c2")]
n, L = wl.n_reads, wl.read_len
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
d = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n, stream=stream)
h = torch.empty(n * L, dtype=torch.uint8).pin_memory()
h.copy_(d[: n * L])
offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64).pin_memory()
PROFILE = "--no-profile" not in sys.argv
g = GpuGIR(wl.k, True, device=0, stream=stream, profile=PROFILE, edges_count=wl.expected_distinct_edges(),
           options={"trace": 0 if "--no-trace" in sys.argv else 1})
for i in range(6 if "--no-trace" in sys.argv else 3):
    torch.cuda.synchronize()
    print(f"=== step {i}", file=sys.stderr, flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g.reset_profile()
    e0.record()
    g.reset()
    g.add_reads_host_ptr(h.data_ptr(), offs.data_ptr(), n)
    dig = g.digest()
    e1.record()
    torch.cuda.synchronize()
    print(f"=== step {i} took {e0.elapsed_time(e1):.2f} ms", {k: round(v['ms'], 2) for k, v in g.profile().items() if v['launches']},
          file=sys.stderr, flush=True)
