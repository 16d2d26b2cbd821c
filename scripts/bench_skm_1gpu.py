#!/usr/bin/env python3
"""Kernel times of the super-k-mer exchange on ONE GPU (no NVLink): rank 0 of an emulated world
sends the C2 read set, every owner unrolls what it received.  Prints one JSON line per world."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import katome_b200 as K  # noqa: E402
from katome_b200.workloads import BY_NAME  # noqa: E402


def main():
    wl = BY_NAME[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    L, k, n = wl.read_len, wl.k, wl.n_reads
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    d = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    K.synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, 0, n, stream=s)
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device=dev)
    for W in (1, 8):
        gs = [K.GpuGIR(k, True, world_size=W, rank=r, force_partition=True, stream=s, profile=True,
                       edges_count=wl.expected_distinct_edges() // W) for r in range(W)]
        windows = n * (L - k + 1)
        for it in range(3):
            for g in gs:
                g.reset()
                g.reset_profile()
            prep = [g.mg_skm_prepare(windows) for g in gs]
            peers, cap = [p[0] for p in prep], prep[0][2]
            cur_ptr, kc_ptr = gs[0].mg_skm_scatter_reads_device(d, offs, n, n * L, peers)
            cur = torch.as_tensor(K.DeviceArray(cur_ptr, W), device=dev)
            kc = torch.as_tensor(K.DeviceArray(kc_ptr, W), device=dev)
            torch.cuda.synchronize()
            n_rec = int((cur - torch.arange(W, device=dev) * cap).sum())
            for r, g in enumerate(gs):
                ends = torch.arange(W, dtype=torch.int64, device=dev) * cap
                ends[0] = min(int(cur[r]) - r * cap, cap)  # only rank 0 sent
                g.mg_skm_insert_buckets(ends, int(kc[r]))
                g.finalize()
            torch.cuda.synchronize()
        prof0 = gs[0].profile()
        owner_ms = [sum(g.profile().get(n, {"ms": 0})["ms"] for n in ("unroll_records", "scatter_received", "scatter_records")) for g in gs]
        digs = [g.digest() for g in gs]
        assert sum(x[2] for x in digs) == 2 * windows, (digs, windows)
        print(json.dumps({"world": W, "workload": wl.name, "windows": windows, "records": n_rec,
                          "records_per_window": n_rec / windows, "spilled": gs[0].mg_skm_spill()[1],
                          "sender_ms": prof0["scatter_superkmers_p2p"]["ms"], "owner_ms_sum": sum(owner_ms),
                          "owner_ms": owner_ms, "key_share": [int(x) / windows for x in kc.tolist()],
                          "rank0": {a: round(b["ms"], 3) for a, b in prof0.items() if b["launches"]}}))
        for g in gs:
            g.close()


if __name__ == "__main__":
    main()
