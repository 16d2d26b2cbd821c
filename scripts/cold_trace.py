import sys, time, torch
sys.path.insert(0, "/root/repo")
from katome_b200 import GpuGIR, synth_reads_device
from katome_b200.workloads import C2 as wl
L, n = wl.read_len, wl.n_reads
d = torch.empty(n * L + 64, dtype=torch.uint8, device="cuda")
synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n, stream=torch.cuda.current_stream().cuda_stream)
h = torch.empty(n * L, dtype=torch.uint8).pin_memory(); h.copy_(d[: n * L])
offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64).pin_memory()
torch.cuda.synchronize()
for it in range(3):
    t0 = time.perf_counter(); g = GpuGIR(wl.k, True, device=0, profile=(it == 2)); torch.cuda.synchronize()
    if it == 2: g.set_option("trace", 1)
    t1 = time.perf_counter(); g.add_reads_host_ptr(h.data_ptr(), offs.data_ptr(), n); torch.cuda.synchronize()
    t2 = time.perf_counter(); dg = g.digest(); t3 = time.perf_counter()
    inf = g.info(); prof = g.profile() if it == 2 else None
    t4 = time.perf_counter(); g.close(); torch.cuda.synchronize(); t5 = time.perf_counter()
    print(it, "create %.1f add %.1f digest %.1f close %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2), 1e3*(t5-t4)), inf["grow_events"], file=sys.stderr)
    if prof: print({k: (v["launches"], round(v["ms"], 2)) for k, v in prof.items()}, file=sys.stderr)

# the same from a FASTQ file in /dev/shm
import numpy as np, os
rec = np.empty((n, 2 * L + 7), dtype=np.uint8)
rec[:, 0:3] = np.frombuffer(b"@r\n", np.uint8); rec[:, 3:3 + L] = h.numpy().reshape(n, L)
rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", np.uint8); rec[:, 6 + L:6 + 2 * L] = ord("I"); rec[:, 6 + 2 * L] = ord("\n")
path = "/dev/shm/ktg_cold.fastq"; rec.tofile(path); del rec
for it in range(3):
    t0 = time.perf_counter()
    g, nb = GpuGIR.create([path], "fastq", True, 0, k=wl.k, device=0, options={"trace": 1} if it == 2 else None)
    t1 = time.perf_counter(); dg = g.digest(); t2 = time.perf_counter(); g.close(); t3 = time.perf_counter()
    print("file", it, "create %.1f digest %.1f close %.1f ms" % (1e3*(t1-t0), 1e3*(t2-t1), 1e3*(t3-t2)), file=sys.stderr)
os.unlink(path)
