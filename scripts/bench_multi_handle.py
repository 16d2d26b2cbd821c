#!/usr/bin/env python3
"""End to end through ONE handle over N GPUs of this process (ktg_config.n_devices; csrc/multi.cuh): what a
katome host gets from Build::create on a multi-GPU box without torch, NCCL or several processes.

    python scripts/bench_multi_handle.py --devices 0,1,2,3 [--workload c3] [--steps 3]

A step = ktg_reset + ktg_add_reads (pinned host reads of the whole workload) + ktg_digest, wall clock.  The
digest is compared with the committed golden digest of the workload (tests/golden/baseline_digests.json).
Prints one JSON line.  `--devices 0,0` shards one GPU (no NVLink involved: a functional check only)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0")
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--export", action="store_true", help="also time remove_weak_edges(3) + export_graph")
    ap.add_argument("--trace", action="store_true", help="host timeline of the last step on stderr (option trace)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE")
    args = ap.parse_args()
    import numpy as np
    import torch
    from katome_b200 import GpuGIR, synth_reads_device
    from katome_b200.workloads import BY_NAME
    wl = BY_NAME[args.workload]
    ids = [int(x) for x in args.devices.split(",")]
    L, n = wl.read_len, wl.n_reads
    # the reads, generated on the first device in pieces and kept in pinned host memory
    h = torch.empty(n * L, dtype=torch.uint8).pin_memory()
    torch.cuda.set_device(ids[0])
    step = (256 << 20) // L
    d = torch.empty(step * L + 64, dtype=torch.uint8, device="cuda")
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        synth_reads_device(d, wl.seed, wl.genome_len, L, wl.err_ppm, r0, r1, stream=torch.cuda.current_stream().cuda_stream)
        h[r0 * L: r1 * L].copy_(d[: (r1 - r0) * L])
    torch.cuda.synchronize()
    del d
    offs = torch.arange(0, (n + 1) * L, L, dtype=torch.int64).pin_memory()
    kw = {"device_ids": ids} if len(ids) > 1 else {"device": ids[0]}
    kw["options"] = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}
    g = GpuGIR(wl.k, True, edges_count=wl.expected_distinct_edges(), **kw)
    times, dig = [], None
    for i in range(args.warmup + args.steps):
        if args.trace and i == args.warmup + args.steps - 1:
            g.set_option("trace", 1)
        t0 = time.perf_counter()
        g.reset()
        g.add_reads_host_ptr(h.data_ptr(), offs.data_ptr(), n)
        dig = g.digest()
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_digests.json")))["workloads"].get(args.workload)
    line = {"what": "one handle over several GPUs, end to end from pinned host memory", "devices": ids,
            "workload": wl.name, "windows": wl.n_windows, "ms_per_step": 1e3 * min(times),
            "value": wl.n_windows / min(times), "unit": "k-mers/s", "h2d_bytes_per_step": n * L, "digest": list(dig),
            "golden_equal": None if gold is None else list(dig) == gold["built"]}
    if args.export:
        t0 = time.perf_counter()
        g.remove_weak_edges(3)
        graph = g.export_graph()
        line["export_ms"] = 1e3 * (time.perf_counter() - t0)
        line["export_nodes"], line["export_edges"] = int(len(graph["node_lo"])), int(len(graph["weight"]))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
