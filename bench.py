#!/usr/bin/env python3
"""bench.py -- k-mers inserted/s of the GIR build stage (BASELINE.json metric).

One step = one whole build of the workload's read set (reset table -> pack -> extract ->
partition -> insert -> finalize).  N=1: BASELINE config 2 (4.6 Mbp genome, 100 bp reads,
100x, 0.5 % substitutions, k=31, reverse_complement=true).  N>1: the same per-GPU work on
an N-times larger genome (weak scaling), table hash-sharded with an all-to-all of keys.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = sys.stdout
METRIC = "kmers_inserted_per_sec"
# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full`
# capture of this command at N=1 on workload c2 (profiles/r01_ncu_full_c2_final3.txt)
NCU_SOURCE = "profiles/r01_ncu_full_c2_final3.txt"
NCU_TRAFFIC_C2 = {"scatter_reads": 0.122055e9 + 2.519546e9, "scatter_pages": 2.577942e9 + 2.520448e9,
                  "update_pages": 2.599678e9 + 1.687045e9, "pack_reads": 0.460009e9 + 0.117671e9}
UNIT = "k-mers/s"


def workload_for(name, world):
    from katome_b200.workloads import BY_NAME, Workload
    wl = BY_NAME[name]
    if name == "c5":  # BASELINE config 5 is one 1 Gbp job sharded over the GPUs (strong scaling)
        return wl
    if world > 1:  # weak scaling: per-GPU reads fixed, genome grows with the world
        wl = Workload(f"{wl.name} x{world} (weak)", wl.config_index, wl.genome_len * world, wl.read_len,
                      wl.coverage, wl.err_ppm, wl.k)
    return wl


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(wl, n_reads, rc=True):
    """The oracle port (single thread: the reference is single-threaded by construction,
    prelude.rs:32-34) on the first n_reads reads of the workload."""
    import numpy as np
    from oracle import oracle as O
    reads = O.synth_reads(wl.seed, wl.genome_len, wl.read_len, wl.err_ppm, 0, n_reads)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * wl.read_len
    t0 = time.perf_counter()
    g = O.OracleGIR(wl.k)
    g.add_reads(reads, offsets, rc)
    g.counts()
    dt = time.perf_counter() - t0
    return n_reads * wl.windows_per_read / dt, dt


def cpu_optimistic(wl, n_reads, rc=True):
    """SURVEY 8(d)'s "optimistic CPU" line: the same edge multiset counted with rolling extraction,
    canonical keys and an edge-keyed table, hash-sharded over every host thread
    (oracle/katome_oracle_mt.c).  NOT the reference's work shape; reported beside the faithful port."""
    import numpy as np
    from oracle import oracle as O
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    reads = O.synth_reads(wl.seed, wl.genome_len, wl.read_len, wl.err_ppm, 0, n_reads)
    offsets = np.arange(n_reads + 1, dtype=np.uint64) * wl.read_len
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        dig, _, _ = O.mt_build_digest(wl.k, reads, offsets, rc, threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    windows = n_reads * wl.windows_per_read
    assert dig[2] == (2 if rc else 1) * windows, dig
    return {"value": windows / best, "unit": UNIT, "cores": threads, "kind": "port-optimistic",
            "sample": f"first {n_reads} reads of the workload ({windows} windows, {best:.2f} s), rolling canonical "
                      f"edge-keyed counter hash-sharded over {threads} threads (oracle/katome_oracle_mt.c); "
                      "not the reference's work shape"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload_for(args.workload, 1)
    n = args.cpu_sample_reads
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_sample(wl, n, args.rc)
        if i >= args.warmup:
            vals.append((v, dt))
    v = sum(x for x, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    sample = f"first {n} reads of {wl.name} ({n * wl.windows_per_read} windows), oracle port, 1 thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": wl.name, "k": wl.k, "reverse_complement": args.rc,
                   "note": "the Rust reference cannot be built here (no rustc/cargo); this is the C port"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reads_per_sec": v / wl.windows_per_read,
        "cpu_optimistic": cpu_optimistic(wl, args.cpu_opt_reads, args.rc),
    }
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(pw) if pw else None}


# --------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import numpy as np
    from katome_b200 import GpuGIR, synth_reads_device
    from katome_b200.dist import ShardedGIR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = workload_for(args.workload, world)
    L, k = wl.read_len, wl.k
    n_total = wl.n_reads
    n_local = n_total // world
    r0 = rank * n_local
    stream = torch.cuda.current_stream().cuda_stream

    # synthetic reads of this rank, resident in HBM (larger than L2: 460 MB per rank)
    d_bases = torch.empty(n_local * L + 64, dtype=torch.uint8, device=dev)
    synth_reads_device(d_bases, wl.seed, wl.genome_len, L, wl.err_ppm, r0, r0 + n_local, stream=stream)
    d_offs = torch.arange(0, (n_local + 1) * L, L, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    windows_local = n_local * wl.windows_per_read
    windows_total = windows_local * world
    hint = wl.expected_distinct_edges() if args.hint else None

    # a batch holds at most ~400 M windows (32-bit positions inside the partitioner)
    n_batches = args.batches or max(1, -(-windows_local // 400_000_000))
    per = -(-n_local // n_batches)
    cuts = [min(i * per, n_local) for i in range(n_batches + 1)]

    def feed(add):
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b > a:  # offsets stay absolute: bias the base pointer instead of rebuilding them
                add(d_bases.data_ptr() + a * L, d_offs[: b - a + 1], b - a, (b - a) * L)

    exchange = None
    if world == 1:
        g = GpuGIR(k, args.rc, device=local, stream=stream, profile=True, edges_count=hint,
                   sub_table_log2_bytes=args.sub_log2)
        def step():
            g.reset()
            feed(g.add_reads_device)
            g.finalize()
        digest = g.digest
        builder = g
    else:
        sg = ShardedGIR(k, args.rc, edges_count=hint, profile=True, sub_table_log2_bytes=args.sub_log2)
        exchange = sg.exchange
        def step():
            sg.reset()
            feed(sg.add_reads_device)
            sg.finalize()
        digest = sg.digest
        builder = sg.gir

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    builder.reset_profile()
    launches0 = builder.info()["kernel_launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    prof = builder.profile()
    info = builder.info()
    launches = info["kernel_launches"] - launches0 - 1  # info() itself launches one scan
    dig = digest()
    # weight conservation: every window adds 1 to each strand (2 to a palindrome)
    assert dig[2] == (2 if args.rc else 1) * windows_total, (dig, windows_total)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = windows_total / (ms_step * 1e-3)

    # ---- end to end: pinned host reads -> H2D -> build -> D2H of the digest, every step
    e2e = None
    if not args.no_e2e and n_batches == 1:
        h_bases = torch.empty(n_local * L, dtype=torch.uint8).pin_memory()
        h_bases.copy_(d_bases[: n_local * L])
        h_offs = torch.arange(0, (n_local + 1) * L, L, dtype=torch.int64).pin_memory()
        if world == 1:
            def estep():
                g.reset()
                g.add_reads_host_ptr(h_bases.data_ptr(), h_offs.data_ptr(), n_local)
                return g.digest()
        else:
            def estep():
                sg.reset()
                sg.add_reads_host(h_bases, h_offs, n_local)
                return sg.digest()
        # the call a user makes: per-launch timing off (the events around every kernel cost ~20 us of GPU
        # time per launch, 0.8 ms of a host-fed C2 build that has ~40 launches); the kernel breakdown below
        # comes from one more, untimed, step with the timing back on
        builder.set_profile(False)
        estep()
        barrier()
        # the PCIe floor of this step: the same bytes, copy only
        e0.record()
        for _ in range(3):
            d_bases[: n_local * L].copy_(h_bases, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        h2d_ms = e0.elapsed_time(e1) / 3
        barrier()
        e0.record()
        for _ in range(args.steps):
            edig = estep()
        e1.record()
        barrier()
        ems = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        assert edig == dig, (edig, dig)
        builder.set_profile(True)
        builder.reset_profile()
        e0.record()
        assert estep() == dig
        e1.record()
        barrier()
        profiled_ms = e0.elapsed_time(e1)
        eprof = builder.profile()
        e2e = {"value": windows_total / (ems / args.steps * 1e-3), "unit": UNIT,
               # one GPU: equally long reads need no offsets on the device (generated there); N > 1 copies both
               "h2d_bytes_per_step": (n_local * L + ((n_local + 1) * 8 if world > 1 else 0)) * world,
               "d2h_bytes_per_step": 40 * world,
               "ms_per_step": ems / args.steps, "h2d_copy_only_ms": h2d_ms,
               "h2d_gbs": n_local * L / h2d_ms / 1e6,
               "timing": "per-launch events off in the timed steps; `kernels` from one more step with them on",
               "profiled_step_ms": profiled_ms,
               "kernels": {n: {"launches": p["launches"], "ms_per_step": p["ms"]}
                           for n, p in eprof.items() if p["launches"]}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (CUDA events inside the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kern = {n: p for n, p in prof.items() if p["launches"]}
    top = max(kern, key=lambda n: kern[n]["ms"]) if kern else None
    roof = None
    if top:
        avg_ms = kern[top]["ms"] / kern[top]["launches"]
        # algorithmic bytes: SURVEY 8(d) per-window figure x windows one launch processes
        alg = wl.algorithmic_bytes_per_window() * windows_local
        ach = alg / (avg_ms * 1e-3) / 1e9
        traffic = NCU_TRAFFIC_C2.get(top) if (args.workload == "c2" and world == 1) else None
        roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": NCU_SOURCE if traffic else None,
                "peak_source": peak_src, "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": alg,
                "kernel_share_of_step": kern[top]["ms"] / args.steps / ms_step,
                # the same algorithmic bytes over the WHOLE step (every kernel of the build), for scale
                "step": {"achieved": alg / (ms_step * 1e-3) / 1e9, "frac": alg / (ms_step * 1e-3) / 1e9 / peak,
                         "unit": "GB/s"}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.workload == "c5" else "weak",
        "vs_baseline": None, "dtype": "u64" if k <= 32 else "u128",
        "data": "synthetic",
        "config": {"workload": wl.name, "genome_len": wl.genome_len, "read_len": L, "coverage": wl.coverage,
                   "err_ppm": wl.err_ppm, "k": k, "reverse_complement": args.rc, "reads": n_total,
                   "windows": windows_total, "parallelism": f"hash-shard x{world}", "exchange": exchange,
                   "batches_per_step": n_batches,
                   "l2": "inputs (460 MB of reads per GPU) and table exceed the 126 MB L2; no explicit flush",
                   "capacity_hint": bool(args.hint)},
        "reads_per_sec": n_total / (ms_step * 1e-3),
        "edge_inserts_per_sec_reference_equivalent": (2 if args.rc else 1) * value,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
        "kernels": {n: {"launches": p["launches"], "ms_per_step": p["ms"] / args.steps} for n, p in kern.items()},
        "table": {"bytes": info["table_bytes"], "sub_tables": info["n_sub_tables"], "slot_bytes": info["slot_bytes"],
                  "load": info["occupied_slots"] / max(1, info["capacity_slots"]), "partitioned": info["partitioned"]},
        "digest": {"D": dig[0], "edges": dig[1], "sum_w": dig[2], "max_w": dig[3]},
    }
    if world == 1 and not args.no_cpu:
        v, dt = cpu_sample(wl, args.cpu_sample_reads, args.rc)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"first {args.cpu_sample_reads} reads of the workload "
                                          f"({args.cpu_sample_reads * wl.windows_per_read} windows, {dt:.1f} s), "
                                          "oracle port of hm_gir.rs:39-153, 1 thread"}
        line["cpu_optimistic"] = cpu_optimistic(wl, args.cpu_opt_reads, args.rc)
    if not args.no_probe and world == 1:
        # the random-access roofline of SURVEY 8(d): uniformly random "load key + atomicAdd weight" over an
        # array as large as this build's table (and over one that fits in L2), measured on this GPU now
        from katome_b200 import random_access_probe
        n_upd = 1 << 28
        sb = info["slot_bytes"]
        tb = max(int(info["table_bytes"]) // sb * sb, 1 << 24)
        at_table = n_upd / (random_access_probe(tb, n_upd, sb) * 1e-3)
        in_l2 = n_upd / (random_access_probe(16 << 20, n_upd, sb) * 1e-3)
        line["random_access"] = {"unit": "updates/s", "table_bytes": tb, "at_table_footprint": at_table,
                                 "at_16MiB_l2_resident": in_l2,
                                 # SURVEY 8(d) random_access_fraction = windows/s / random updates/s
                                 "value_over_at_table_footprint": value / at_table,
                                 "value_over_l2_resident": value / in_l2}
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if world > 1:
        dist.destroy_process_group()


def main():
    # Libraries (NCCL prints its version) write to the C-level stdout; the contract is ONE JSON
    # line there, so everything else goes to stderr and the line is written to the saved fd.
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000)
    ap.add_argument("--no-hint", dest="hint", action="store_false")
    ap.add_argument("--no-rc", dest="rc", action="store_false",
                    help="reverse_complement=false (SURVEY 8d asks for one such run of C2); the default is the settings default, true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-opt-reads", type=int, default=1_000_000,
                    help="sample of the multi-threaded optimistic CPU counter")
    ap.add_argument("--no-probe", action="store_true", help="skip the random-access roofline probe")
    ap.add_argument("--sub-log2", type=int, default=0)
    ap.add_argument("--batches", type=int, default=0, help="add_reads calls per step (0: as few as fit)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
