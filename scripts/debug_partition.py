import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from katome_b200 import GpuGIR, synth_reads_device
n, L = 200000, 100
dev = torch.device("cuda", 0)
d_bases = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
synth_reads_device(d_bases, 7, 1000000, L, 5000, 0, n, stream=torch.cuda.current_stream().cuda_stream)
d_offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
g = GpuGIR(31, True, world_size=2, rank=0, stream=torch.cuda.current_stream().cuda_stream, profile=True)
print("timed-style", g.partition_reads_device(d_bases, d_offs, n, n * L)[1]); g.finalize()
h_bases = torch.empty(n * L, dtype=torch.uint8).pin_memory()
h_bases.copy_(d_bases[: n * L])
h_offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64).pin_memory()
stage = torch.empty_like(d_bases)
for it in range(2):
    g.reset()
    stage[: n * L].copy_(h_bases, non_blocking=True)
    so = h_offs.to(dev, non_blocking=True)
    print("e2e-style", it, g.partition_reads_device(stage, so, n, n * L)[1]); g.finalize()
    print(g.digest())
print("ok")
