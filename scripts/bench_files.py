#!/usr/bin/env python3
"""Build::create from a FASTQ file: device-side record parsing vs the host reader (same result).
Usage: python scripts/bench_files.py [n_reads]   (writes a synthetic FASTQ under /tmp)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def write_fastq(path, reads, L):
    n = reads.size // L
    rec = np.empty((n, 11 + L + 3 + L + 1), dtype=np.uint8)  # "@r%08d\n" seq "\n+\n" qual "\n"
    ids = np.char.zfill(np.arange(n).astype("U8"), 8)
    rec[:, 0:2] = np.frombuffer(b"@r", np.uint8)
    rec[:, 2:10] = np.frombuffer("".join(ids).encode(), np.uint8).reshape(n, 8)
    rec[:, 10] = 10
    rec[:, 11:11 + L] = reads.reshape(n, L)
    rec[:, 11 + L:14 + L] = np.frombuffer(b"\n+\n", np.uint8)
    rec[:, 14 + L:14 + 2 * L] = ord("I")
    rec[:, 14 + 2 * L] = 10
    rec.tofile(path)
    return rec.size


def main():
    from katome_b200 import GpuGIR
    from oracle import oracle as O
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    L, k = 100, 31
    reads = O.synth_reads(0x6B61746F6D65 + 1, 4_600_000, L, 5000, 0, n)
    path = "/tmp/ktg_bench.fastq"
    size = write_fastq(path, reads, L)
    res = {}
    for name, opts in (("device_parse", {}), ("host_parse", {"host_parse": 1})):
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            g, nbytes = GpuGIR.create([path], "fastq", True, 0, k=k, edges_count=2 * 4_600_000 * 12, options=opts)
            dig = g.digest()
            dt = time.perf_counter() - t0
            g.close()
            best = dt if best is None else min(best, dt)
        res[name] = (best, dig, nbytes)
        print(f"{name:13s} {best * 1e3:8.1f} ms  {size / best / 1e9:6.2f} GB/s of FASTQ  "
              f"{n * (L - k + 1) / best / 1e9:6.2f} G k-mers/s  digest {dig[0]:#x}")
    assert res["device_parse"][1:] == res["host_parse"][1:]
    os.remove(path)


if __name__ == "__main__":
    main()
