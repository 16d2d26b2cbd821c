#!/usr/bin/env python3
"""bench.py -- k-mers inserted/s of the GIR build stage (BASELINE.json metric).

One step = one whole build of the workload's read set (reset table -> pack -> extract ->
partition -> insert -> finalize).
  N=1   BASELINE config 2 (4.6 Mbp genome, 100 bp reads, 100x, 0.5 % substitutions, k=31,
        reverse_complement=true), the configuration the metric is quoted on.
  N>1   BASELINE config 3 (46 Mbp, 150 bp, 50x, k=31): ONE job whose reads are split over the
        ranks and whose table is hash-sharded (strong scaling), so the result -- and its digest --
        is the same at every N.  `--workload c3k63` is the u128 variant, `--workload c5` config 5,
        `--weak` the round-1 curve (config 2 per GPU on an N-times larger genome).
Every line carries `digest_check`: the GPU result (after the build, after remove_weak_edges and after
standardize_edges) compared with the CPU oracle run on the same reads (oracle/katome_oracle_mt.c,
outside every timed region) and with the committed golden digests (tests/golden/baseline_digests.json).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = sys.stdout
METRIC = "kmers_inserted_per_sec"
UNIT = "k-mers/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of every big kernel, from the committed
# `ncu --set full` capture of this command at N=1 (written by scripts/ncu_summary.py)
NCU_TRAFFIC = "profiles/ncu_traffic.json"
FILTER_T, STD_T, STD_G_FACTOR = 2, 3, 64  # the stages of tests/golden/make_baseline_digests.py


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def workload_for(name, world, weak=False):
    from katome_b200.workloads import BY_NAME, Workload
    wl = BY_NAME[name or ("c2" if world == 1 or weak else "c3")]
    if weak and world > 1:  # per-GPU reads fixed, genome grows with the world
        wl = Workload(f"{wl.name} x{world} (weak)", wl.config_index, wl.genome_len * world, wl.read_len,
                      wl.coverage, wl.err_ppm, wl.k)
    return wl


def host_threads():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


# --------------------------------------------------------------------------- CPU arm
def cpu_sample(wl, n_reads, rc=True):
    """The oracle port (single thread: the reference is single-threaded by construction,
    prelude.rs:32-34) on the first n_reads reads of the workload."""
    import numpy as np
    from oracle import oracle as O
    reads = O.synth_reads_mt(wl.seed, wl.genome_len, wl.read_len, wl.err_ppm, 0, n_reads)
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
    threads = host_threads()
    reads = O.synth_reads_mt(wl.seed, wl.genome_len, wl.read_len, wl.err_ppm, 0, n_reads)
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
    wl = workload_for(args.workload, args.gpus, args.weak)
    n = wl.n_reads if args.full else min(args.cpu_sample_reads, wl.n_reads)
    vals = []
    steps, warmup = (1, 0) if args.full else (args.steps, args.warmup)
    for i in range(warmup + steps):
        v, dt = cpu_sample(wl, n, args.rc)
        if i >= warmup:
            vals.append((v, dt))
    v = sum(x for x, _ in vals) / len(vals)
    ms = 1e3 * sum(d for _, d in vals) / len(vals)
    what = "ALL" if n == wl.n_reads else "first"
    sample = f"{what} {n} reads of {wl.name} ({n * wl.windows_per_read} windows), oracle port, 1 thread"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak" if (args.weak or args.gpus == 1) else "strong", "vs_baseline": None,
        "dtype": "u64" if wl.k <= 32 else "u128", "data": "synthetic",
        "config": {"workload": wl.name, "k": wl.k, "reverse_complement": args.rc, "reads_in_sample": n,
                   "note": "the Rust reference cannot be built here (no rustc/cargo); this is the C port of "
                           "hm_gir.rs:39-153, single-threaded like the reference (prelude.rs:32-34)"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reads_per_sec": v / wl.windows_per_read,
        "cpu_optimistic": cpu_optimistic(wl, min(args.cpu_opt_reads, wl.n_reads), args.rc),
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


# --------------------------------------------------------------------------- the checker
def oracle_stages(wl, rc, dev, stream):
    """[D, |E|, sum w, max w] after the build, after remove_weak_edges(FILTER_T) and after
    standardize_edges(STD_G_FACTOR * G, k, STD_T), computed by the CPU oracle (katome_oracle_mt.c) on the
    reads of the device generator -- the very bytes the GPU build consumes -- copied back chunk by chunk.
    Outside every timed region; cached per workload in the temp directory (the driver runs N = 1, 2, 4, 8
    of one workload back to back)."""
    import numpy as np
    import torch
    from katome_b200 import synth_reads_device
    from oracle import oracle as O
    src = open(os.path.join(ROOT, "oracle", "katome_oracle_mt.c"), "rb").read()
    key = hashlib.sha1(repr((wl.seed, wl.genome_len, wl.read_len, wl.coverage, wl.err_ppm, wl.k, rc, FILTER_T, STD_T,
                             STD_G_FACTOR)).encode() + src).hexdigest()[:20]
    cache = os.path.join(tempfile.gettempdir(), f"ktg_oracle_stages_{key}.json")
    try:
        doc = json.load(open(cache))
        doc["cached"] = True
        return doc
    except (OSError, ValueError):
        pass
    t0 = time.perf_counter()
    threads = host_threads()
    m = O.MtCounter(wl.k, rc, threads)
    L, n = wl.read_len, wl.n_reads
    step = max(1, (256 << 20) // L)
    d_buf = torch.empty(step * L + 64, dtype=torch.uint8, device=dev)
    h_buf = torch.empty(step * L, dtype=torch.uint8).pin_memory()
    offs = np.arange(step + 1, dtype=np.uint64) * L
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        synth_reads_device(d_buf, wl.seed, wl.genome_len, L, wl.err_ppm, r0, r1, stream=stream)
        h_buf[: (r1 - r0) * L].copy_(d_buf[: (r1 - r0) * L])
        torch.cuda.synchronize()
        m.add_reads_ptr(h_buf.data_ptr(), offs.ctypes.data, r1 - r0)
    t_build = time.perf_counter() - t0
    doc = {"built": list(m.digest())}
    m.remove_weak_edges(FILTER_T)
    doc["filtered"] = list(m.digest())
    m.standardize_edges(STD_G_FACTOR * wl.genome_len, wl.k, STD_T)
    doc["standardized"] = list(m.digest())
    nr, nb = m.counters()
    doc.update(accepted_reads=nr, accepted_bytes=nb, threads=threads, build_s=t_build,
               total_s=time.perf_counter() - t0, windows_per_s=wl.n_windows / t_build, cached=False)
    del m
    try:
        json.dump(doc, open(cache, "w"))
    except OSError:
        pass
    return doc


def golden_stages(args, wl):
    if not args.rc or args.weak:
        return None
    try:
        doc = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_digests.json")))
    except OSError:
        return None
    for g in doc["workloads"].values():
        if all(g[f] == getattr(wl, f) for f in ("seed", "genome_len", "read_len", "coverage", "err_ppm", "k")):
            return g
    return None


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
    numa = None
    if world > 1:
        from katome_b200.dist import bind_to_gpu_numa
        numa = bind_to_gpu_numa(local)  # before any pinned allocation
        dist.init_process_group("nccl", device_id=dev)
    wl = workload_for(args.workload, world, args.weak)
    L, k = wl.read_len, wl.k
    n_total = wl.n_reads
    # reads [r0, r0 + n_local) of the ONE job belong to this rank
    r0 = n_total * rank // world
    n_local = n_total * (rank + 1) // world - r0
    stream = torch.cuda.current_stream().cuda_stream

    # synthetic reads of this rank, resident in HBM (larger than L2)
    d_bases = torch.empty(n_local * L + 64, dtype=torch.uint8, device=dev)
    synth_reads_device(d_bases, wl.seed, wl.genome_len, L, wl.err_ppm, r0, r0 + n_local, stream=stream)
    d_offs = torch.arange(0, (n_local + 1) * L, L, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    windows_local = n_local * wl.windows_per_read
    windows_total = n_total * wl.windows_per_read
    hint = wl.expected_distinct_edges() if args.hint else None

    # the pinned host copy of this rank's reads (end-to-end legs)
    h_bases = h_offs = None
    fresh = None
    if not args.no_e2e:
        h_bases = torch.empty(n_local * L, dtype=torch.uint8).pin_memory()
        h_bases.copy_(d_bases[: n_local * L])
        h_offs = torch.arange(0, (n_local + 1) * L, L, dtype=torch.int64).pin_memory()
        if world == 1 and not args.no_consumer:
            fresh = fresh_handle_legs(args, wl, h_bases, h_offs, n_local, dev)

    n_batches = args.batches or 1
    per = -(-n_local // n_batches)
    cuts = [min(i * per, n_local) for i in range(n_batches + 1)]

    def feed(add):
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b > a:  # offsets stay absolute: bias the base pointer instead of rebuilding them
                add(d_bases.data_ptr() + a * L, d_offs[: b - a + 1], b - a, (b - a) * L)

    exchange = None
    if world == 1:
        g = GpuGIR(k, args.rc, device=local, stream=stream, profile=True, edges_count=hint,
                   sub_table_log2_bytes=args.sub_log2, options=args.options)
        sg = None
        def step():
            g.reset()
            feed(g.add_reads_device)
            g.finalize()
        top, builder = g, g
    else:
        sg = ShardedGIR(k, args.rc, edges_count=hint, profile=True, sub_table_log2_bytes=args.sub_log2,
                        exchange=args.exchange, options=args.options)
        def step():
            sg.reset()
            feed(sg.add_reads_device)
            sg.finalize()
        top, builder = sg, sg.gir

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

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
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if sampler else None
    prof = builder.profile()
    info = builder.info()
    launches = info["kernel_launches"] - launches0 - 1  # info() itself launches one scan
    dig = top.digest()
    # weight conservation: every window adds 1 to each strand (2 to a palindrome)
    assert dig[2] == (2 if args.rc else 1) * windows_total, (dig, windows_total)
    ms_step = ms / args.steps
    value = windows_total / (ms_step * 1e-3)
    if sg is not None:
        exchange = sg.last_exchange

    # ---- end to end: pinned host reads -> H2D -> build -> D2H of the digest, every step
    e2e = None
    if not args.no_e2e:
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
        ems = max_over_ranks(e0.elapsed_time(e1))
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
               "h2d_bytes_per_step": n_total * L + ((n_total + world) * 8 if world > 1 else 0),
               "d2h_bytes_per_step": 40 * world,
               "ms_per_step": ems / args.steps, "h2d_copy_only_ms": h2d_ms,
               "h2d_gbs": n_local * L / h2d_ms / 1e6,
               "timing": "per-launch events off in the timed steps; `kernels` from one more step with them on",
               "profiled_step_ms": profiled_ms,
               "kernels": {n: {"launches": p["launches"], "ms_per_step": p["ms"]}
                           for n, p in eprof.items() if p["launches"]}}

    # ---- what katome consumes (one GPU): build from the host, filter, the graph Convert::create_from
    # takes (hm_gir.rs:156-226) back on the host; a build from a FASTQ file; a cold build without a hint
    consumer = None
    if world == 1 and not args.no_e2e and not args.no_consumer:
        consumer = consumer_legs(args, wl, g, h_bases, h_offs, n_local, dig, dev, stream)
        for leg in ("cold", "file"):  # measured before the bench's own builder existed; same result?
            if fresh and "digest" in fresh.get(leg, {}):
                fresh[leg]["digest_equal"] = tuple(fresh[leg].pop("digest")) == tuple(dig)
        consumer.update(fresh or {})

    # ---- parity at full size: the oracle on the same reads, stage by stage
    check = None
    if not args.no_check:
        builder.set_profile(False)
        step()  # the table of the resident-input build again
        stages = {"built": list(top.digest())}
        top.remove_weak_edges(FILTER_T)
        stages["filtered"] = list(top.digest())
        top.standardize_edges(STD_G_FACTOR * wl.genome_len, k, STD_T)
        stages["standardized"] = list(top.digest())
        barrier()
        if rank == 0:
            t0 = time.perf_counter()
            ora = oracle_stages(wl, args.rc, dev, stream)
            gold = golden_stages(args, wl)
            names = ("built", "filtered", "standardized")
            eq = all(stages[s] == ora[s] for s in names)
            check = {"oracle": "mt", "equal": bool(eq), "stages": list(names),
                     "what": f"[D, |E|, sum w, max w] after the build, after remove_weak_edges({FILTER_T}) and after "
                             f"standardize_edges({STD_G_FACTOR} G, k, {STD_T}); oracle/katome_oracle_mt.c on the reads "
                             "of the device generator, copied back",
                     "gpu": stages, "cpu": {s: ora[s] for s in names},
                     "golden_equal": None if gold is None else bool(all(stages[s] == gold[s] for s in names)),
                     "oracle_threads": ora["threads"], "oracle_build_s": ora["build_s"],
                     "oracle_kmers_per_s": ora["windows_per_s"], "oracle_cached": ora["cached"],
                     "check_s": time.perf_counter() - t0}
            if not eq or check["golden_equal"] is False:
                log("DIGEST MISMATCH", json.dumps(check))
        builder.set_profile(True)
        barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline (CUDA events around every launch, inside the timed region)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kern = {n: p for n, p in prof.items() if p["launches"]}
    roof = roofline(args, wl, world, kern, info, windows_local, ms_step, peak, peak_src)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if (args.weak or world == 1) else "strong",
        "vs_baseline": None, "dtype": "u64" if k <= 32 else "u128",
        "data": "synthetic",
        "config": {"workload": wl.name, "genome_len": wl.genome_len, "read_len": L, "coverage": wl.coverage,
                   "err_ppm": wl.err_ppm, "k": k, "reverse_complement": args.rc, "reads": n_total,
                   "windows": windows_total, "parallelism": f"hash-shard x{world}", "exchange": exchange,
                   "batches_per_step": n_batches,
                   "l2": f"inputs ({n_local * L / 1e6:.0f} MB of reads per GPU) and table exceed the 126 MB L2; "
                         "no explicit flush",
                   "capacity_hint": bool(args.hint), "options": args.options or None, "numa_node_rank0": numa},
        "reads_per_sec": n_total / (ms_step * 1e-3),
        "edge_inserts_per_sec_reference_equivalent": (2 if args.rc else 1) * value,
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
        "digest_check": check,
        "kernels": {n: {"launches": p["launches"], "ms_per_step": p["ms"] / args.steps} for n, p in kern.items()},
        "table": {"bytes": info["table_bytes"], "sub_tables": info["n_sub_tables"], "slot_bytes": info["slot_bytes"],
                  "load": info["occupied_slots"] / max(1, info["capacity_slots"]), "partitioned": info["partitioned"]},
        "digest": {"D": dig[0], "edges": dig[1], "sum_w": dig[2], "max_w": dig[3]},
    }
    if consumer:
        line["consumer"] = consumer
    if world == 1 and not args.no_cpu:
        n_cpu = min(args.cpu_sample_reads, wl.n_reads)
        v, dt = cpu_sample(wl, n_cpu, args.rc)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": f"first {n_cpu} reads of the workload "
                                          f"({n_cpu * wl.windows_per_read} windows, {dt:.1f} s), "
                                          "oracle port of hm_gir.rs:39-153, 1 thread"}
        line["cpu_optimistic"] = cpu_optimistic(wl, min(args.cpu_opt_reads, wl.n_reads), args.rc)
    if not args.no_probe and world == 1:
        # the random-access roofline of SURVEY 8(d): uniformly random "load key + atomicAdd weight" over an
        # array as large as this build's table (and over one that fits in L2), measured on this GPU now
        from katome_b200 import random_access_probe
        n_upd = 1 << 28
        sb = 16 if k <= 32 else 32
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


def roofline(args, wl, world, kern, info, windows_local, ms_step, peak, peak_src):
    """`frac` is the WHOLE step: SURVEY 8(d)'s algorithmic bytes of the build (per window: ASCII in + key +
    weight read + weight write) over the step time -- no single kernel performs that unit of work, so no
    kernel's time is its denominator.  `kernels` lists every kernel with its share of the step, the bytes IT
    must move and, where a committed ncu capture of this command exists, its DRAM traffic."""
    if not kern:
        return None
    alg = wl.algorithmic_bytes_per_window() * windows_local
    key = 8 if wl.k <= 32 else 16
    slot = key + 4
    n_windows, bases = windows_local, windows_local / wl.windows_per_read * wl.read_len
    table_slots = info["capacity_slots"]
    own = {  # bytes a kernel has to move for its role, per step
        "pack_reads": bases * (1 + 0.25 + 1 / 8),                   # ASCII in, 2-bit stream + validity bits out
        "check_reads": bases / 8,
        "scatter_reads": bases * 0.25 + n_windows * key,             # 2-bit stream in, one key per window out
        "scatter_pages": n_windows * 2 * key,                        # keys in, keys out (grouped by page)
        "update_pages": n_windows * key + table_slots * slot,    # keys in, table out (a fresh table is not read)
        "scatter_received": n_windows * 2 * key,
    }
    traffic = {}
    try:
        doc = json.load(open(os.path.join(ROOT, NCU_TRAFFIC)))
        name = args.workload or ("c2" if world == 1 else "c3")
        if world == 1 and name in doc:
            traffic = doc[name]
    except (OSError, ValueError):
        pass
    rows = {}
    for n, p in kern.items():
        ms_k = p["ms"] / max(1, args.steps)
        rows[n] = {"ms_per_step": ms_k, "launches_per_step": p["launches"] / max(1, args.steps),
                   "share_of_step": ms_k / ms_step,
                   "own_bytes_per_step": own.get(n), "own_gbs": own[n] / (ms_k * 1e-3) / 1e9 if n in own and ms_k else None,
                   "own_frac_of_peak": own[n] / (ms_k * 1e-3) / 1e9 / peak if n in own and ms_k else None,
                   "dram_bytes_per_step": traffic.get("kernels", {}).get(n)}
    total_traffic = sum(v for v in (r["dram_bytes_per_step"] for r in rows.values()) if v) or None
    top = max(kern, key=lambda n: kern[n]["ms"])
    avg_ms = kern[top]["ms"] / kern[top]["launches"]
    ach = alg / (ms_step * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": total_traffic, "traffic_source": traffic.get("source"),
            "traffic_over_algorithmic": total_traffic / alg if total_traffic else None,
            "scope": "whole step: algorithmic bytes of the build / ms_per_step (no single kernel does the unit of work)",
            "peak_source": peak_src, "algorithmic_bytes_per_step": alg,
            "algorithmic_bytes_per_window": wl.algorithmic_bytes_per_window(),
            "dominant_kernel": {"kernel": top, "avg_launch_ms": avg_ms, "share_of_step": rows[top]["share_of_step"],
                                # the recipe's per-kernel figure: the build's algorithmic bytes over this kernel's time
                                "achieved_recipe": alg / (rows[top]["ms_per_step"] * 1e-3) / 1e9,
                                "frac_recipe": alg / (rows[top]["ms_per_step"] * 1e-3) / 1e9 / peak,
                                "own_gbs": rows[top]["own_gbs"], "dram_bytes_per_step": rows[top]["dram_bytes_per_step"]},
            "kernels": rows}


def fresh_handle_legs(args, wl, h_bases, h_offs, n_local, dev):
    """file and cold (see consumer_legs), measured BEFORE the bench's own builder exists: a fresh handle in a
    process that holds nothing else on the GPU, which is what a first-time caller has."""
    import numpy as np
    import torch
    from katome_b200 import GpuGIR
    L, k = wl.read_len, wl.k
    windows = n_local * wl.windows_per_read
    out = {}

    def best_of(f, n=3):
        best, res = None, None
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = f()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, res

    # -- cold
    try:
        def cold_step():
            gg = GpuGIR(k, args.rc, device=dev.index)
            gg.add_reads_host_ptr(h_bases.data_ptr(), h_offs.data_ptr(), n_local)
            d_ = gg.digest()
            inf = gg.info()
            gg.close()
            return d_, inf
        dt, (cd, inf) = best_of(cold_step, 3)  # (the first of the three also loads the kernels)
        out["cold"] = {"value": windows / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "grow_events": inf["grow_events"],
                       "digest": list(cd),
                       "what": "fresh handle, no capacity hint, no warm-up: ktg_create + ktg_add_reads(host) + digest + "
                               "ktg_destroy (cudaMalloc of every buffer and growth from the sketch inside the timed region)"}
    except Exception as e:  # noqa: BLE001
        out["cold"] = {"error": repr(e)}
    # -- file
    path = None
    try:
        d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
        path = os.path.join(d, f"ktg_bench_{os.getpid()}.fastq")
        rec = np.empty((n_local, 2 * L + 7), dtype=np.uint8)
        rec[:, 0:3] = np.frombuffer(b"@r\n", np.uint8)
        rec[:, 3:3 + L] = h_bases.numpy().reshape(n_local, L)
        rec[:, 3 + L:6 + L] = np.frombuffer(b"\n+\n", np.uint8)
        rec[:, 6 + L:6 + 2 * L] = ord("I")
        rec[:, 6 + 2 * L] = ord("\n")
        rec.tofile(path)
        fbytes = rec.nbytes
        del rec
        def file_step():
            gg, nbytes = GpuGIR.create([path], "fastq", args.rc, 0, k=k, device=dev.index, edges_count=None)
            d_ = gg.digest()
            gg.close()
            return d_, nbytes
        dt, (fd, nbytes) = best_of(file_step, 2)
        assert nbytes == n_local * L, nbytes
        out["file"] = {"value": windows / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "file_bytes": fbytes,
                       "file_gbs": fbytes / dt / 1e9, "d2h_bytes_per_step": 40, "digest": list(fd),
                       "what": "GpuGIR.create([fastq]) = ktg_create_from_files on a fresh handle without a hint "
                               "(reader threads pread blocks into page-locked memory, records cut on the device) + digest; "
                               "file in " + d}
    except Exception as e:  # noqa: BLE001
        out["file"] = {"error": repr(e)}
    finally:
        if path and os.path.exists(path):
            os.unlink(path)
    return out


def consumer_legs(args, wl, g, h_bases, h_offs, n_local, dig, dev, stream):
    """Three more end-to-end figures on one GPU (everything inside the timed region, wall clock around a
    synchronous call sequence, best of 3):
      export  reset -> ktg_add_reads(host) -> remove_weak_edges(3) -> ktg_export_graph into host arrays:
              the hand-off Convert::create_from consumes (hm_gir.rs:156-226); d2h = the real export size
      file    Build::create from a FASTQ file of the workload (builder.rs:42-54), page cache warm
      cold    a fresh handle without a capacity hint: cudaMalloc, growth from the sketch, no warm-up"""
    import numpy as np
    import torch
    from katome_b200 import GpuGIR
    L, k = wl.read_len, wl.k
    windows = n_local * wl.windows_per_read
    out = {}

    def best_of(f, n=3):
        best, res = None, None
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = f()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, res

    g.set_profile(False)
    # -- export (the page-locked arrays are allocated by the first, untimed, call and reused)
    store = {}
    def export_step():
        g.reset()
        g.add_reads_host_ptr(h_bases.data_ptr(), h_offs.data_ptr(), n_local)
        g.remove_weak_edges(3)
        return g.export_graph(pinned=True, out=store)
    try:
        export_step()
        dt, graph = best_of(export_step)
        d2h = sum(int(v.nbytes) for v in graph.values())
        t0 = time.perf_counter()
        g.reset()
        g.add_reads_host_ptr(h_bases.data_ptr(), h_offs.data_ptr(), n_local)
        g.remove_weak_edges(3)
        pageable = g.export_graph()
        torch.cuda.synchronize()
        dt_pageable = time.perf_counter() - t0
        del pageable
        out["export"] = {"value": windows / dt, "unit": UNIT, "ms_per_step": dt * 1e3,
                         "h2d_bytes_per_step": n_local * L, "d2h_bytes_per_step": d2h,
                         "nodes": int(len(graph["node_lo"])), "edges": int(len(graph["weight"])),
                         "ms_per_step_pageable_arrays": dt_pageable * 1e3,
                         "what": "reset -> add_reads(host) -> remove_weak_edges(3) -> export_graph (sorted nodes, "
                                 "src/dst/weight, compress_edge bytes) into page-locked host arrays from ktg_host_alloc "
                                 "(allocated once, reused); sorts are cub (library)"}
        del graph
    except Exception as e:  # noqa: BLE001 -- a leg that fails is reported, the headline stands
        out["export"] = {"error": repr(e)}
    g.set_profile(True)
    return out


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
    ap.add_argument("--workload", default=None, help="c2 | c3 | c3k63 | c5 (default: c2 on one GPU, c3 on several)")
    ap.add_argument("--weak", action="store_true", help="N > 1: the workload per GPU on an N-times larger genome")
    ap.add_argument("--cpu-sample-reads", type=int, default=200_000)
    ap.add_argument("--full", action="store_true", help="--impl reference: the whole workload, once")
    ap.add_argument("--no-hint", dest="hint", action="store_false")
    ap.add_argument("--no-rc", dest="rc", action="store_false",
                    help="reverse_complement=false (SURVEY 8d asks for one such run of C2); the default is the settings default, true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle digest check")
    ap.add_argument("--no-consumer", action="store_true", help="skip the export / file / cold legs")
    ap.add_argument("--cpu-opt-reads", type=int, default=1_000_000,
                    help="sample of the multi-threaded optimistic CPU counter")
    ap.add_argument("--no-probe", action="store_true", help="skip the random-access roofline probe")
    ap.add_argument("--sub-log2", type=int, default=0)
    ap.add_argument("--exchange", default=None, choices=["fused", "direct", "skm", "keys", "nccl"],
                    help="N > 1: which exchange (default: the measured winner for this world size and k)")
    ap.add_argument("--batches", type=int, default=0, help="add_reads calls per step (0: one)")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="ktg_set_option on the builder (measurements; include/katome_gpu.h lists the names)")
    args = ap.parse_args()
    args.options = {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
