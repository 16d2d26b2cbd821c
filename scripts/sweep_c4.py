#!/usr/bin/env python3
"""BASELINE config 4 on one GPU: k sweep (21..63) and substitution-rate sweep (0..2 %) on the 46 Mbp /
150 bp / 50x read set, with the Standardize/Prune threshold filter on the GPU.  One JSON line per
(k, error rate): build throughput, edge counts before / after remove_weak_edges(t) for t = 2, 3, 5,
the result of standardize_edges(G, k, 3), the time of every call, and the size-independent checks
(weight conservation, monotone and idempotent filter)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import katome_b200 as K  # noqa: E402
from katome_b200.workloads import C3_K31, Workload  # noqa: E402

KS = (21, 25, 31, 32, 33, 41, 47, 55, 63)
ERRS = (0, 5000, 10000, 20000)  # ppm


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), out


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    base = C3_K31
    G, L, n = (base.genome_len, base.read_len, base.n_reads) if not quick else (2_000_000, 150, 400_000)
    dev = torch.device("cuda", 0)
    s = torch.cuda.current_stream().cuda_stream
    d = torch.empty(n * L + 64, dtype=torch.uint8, device=dev)
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device=dev)
    configs = [(k, 5000) for k in KS] + [(k, e) for k in (31, 63) for e in ERRS if e != 5000]
    cur_err = None
    for k, err in sorted(configs, key=lambda c: (c[1], c[0])):
        if err != cur_err:
            K.synth_reads_device(d, base.seed, G, L, err, 0, n, stream=s)
            cur_err = err
        wl = Workload("c4", base.config_index, G, L, base.coverage, err, k)
        windows = n * (L - k + 1)
        g = K.GpuGIR(k, True, device=0, stream=s, edges_count=wl.expected_distinct_edges())

        def build():
            g.reset()
            g.add_reads_device(d, offs, n, n * L)
            g.finalize()
        build()  # warm-up (allocations)
        ms = min(timed(build)[0] for _ in range(2))
        D0, E0, S0, M0 = g.digest()
        assert S0 == 2 * windows, (S0, windows)  # every window adds 1 to each strand's edge
        rec = {"k": k, "err_ppm": err, "genome_len": G, "read_len": L, "reads": n, "windows": windows, "build_ms": ms,
               "kmers_per_sec": windows / ms * 1e3, "edges": E0, "max_weight": M0, "key": "u64" if k <= 32 else "u128",
               "filter": []}
        prev = E0
        for t in (2, 3, 5):
            fms, _ = timed(lambda: g.remove_weak_edges(t))
            _, E, S, _ = g.digest()
            assert E <= prev
            fms2, _ = timed(lambda: g.remove_weak_edges(t))  # idempotent
            assert g.digest()[1:3] == (E, S)
            rec["filter"].append({"t": t, "edges": E, "sum_w": S, "ms": fms, "ms_again": fms2})
            prev = E
        build()
        sms, _ = timed(lambda: g.standardize_edges(G, k, 3))
        _, E, S, M = g.digest()
        rec["standardize"] = {"G": G, "k": k, "t": 3, "edges": E, "sum_w": S, "max_weight": M, "ms": sms}
        info = g.info()
        rec["table_bytes"] = info["table_bytes"]
        print(json.dumps(rec), flush=True)
        g.close()
        del g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
