#!/usr/bin/env python3
"""Writes tests/golden/baseline_digests.json: the digests of the BASELINE-size workloads (bench.py's C2, C3 at
k=31 and k=63) computed by the CPU oracle (oracle/katome_oracle_mt.c, digest-equal to the faithful port:
tests/test_oracle_golden.py) at three stages:

    built                          reads -> table           (hm_gir.rs:39-87)
    filtered   remove_weak_edges(2)                         (pruner.rs:95-119, edges.rs:51-58)
    standardized   standardize_edges(64 G, k, 3) on the filtered graph   (standardizer.rs:42-70)

Each entry is [D, |E|, sum w, max w] over the both-strand expanded edge set.
CPU only; run from the repo root:  python tests/golden/make_baseline_digests.py [c2 c3 c3k63]
bench.py and tests/test_gpu_parity.py compare the GPU builds with these numbers AND with the oracle run live.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "baseline_digests.json")
FILTER_T, STD_T = 2, 3
# standardize_edges' genome length: 64 x the real one, so that the ratio is ~0.5 and the rounding of
# standardizer.rs:56-66 is exercised on every weight (with the real length every weight rounds to 0 or 1)
STD_G_FACTOR = 64


def stages(wl, rc=True, n_reads=None, log=None):
    """digests of the three stages for the first n_reads reads of a workload (all of them by default)"""
    import numpy as np
    from oracle import oracle as O
    n = wl.n_reads if n_reads is None else n_reads
    m = O.MtCounter(wl.k, rc)
    step = max(1, (256 << 20) // wl.read_len)
    t0 = time.perf_counter()
    for r0 in range(0, n, step):
        r1 = min(n, r0 + step)
        reads = O.synth_reads_mt(wl.seed, wl.genome_len, wl.read_len, wl.err_ppm, r0, r1)
        m.add_reads(reads, np.arange(r1 - r0 + 1, dtype=np.uint64) * wl.read_len)
        if log:
            log(f"  {wl.name}: {r1}/{n} reads, {time.perf_counter() - t0:.1f} s")
    out = {"built": list(m.digest())}
    m.remove_weak_edges(FILTER_T)
    out["filtered"] = list(m.digest())
    m.standardize_edges(STD_G_FACTOR * wl.genome_len, wl.k, STD_T)
    out["standardized"] = list(m.digest())
    return out


def main():
    from katome_b200.workloads import BY_NAME
    names = sys.argv[1:] or ["c2", "c3", "c3k63"]
    try:
        doc = json.load(open(OUT))
    except OSError:
        doc = {"filter_threshold": FILTER_T, "standardize_threshold": STD_T, "standardize_genome_len_factor": STD_G_FACTOR, "reverse_complement": True,
               "how": "tests/golden/make_baseline_digests.py (oracle/katome_oracle_mt.c)", "workloads": {}}
    for name in names:
        wl = BY_NAME[name]
        t0 = time.perf_counter()
        d = stages(wl, log=lambda s: print(s, file=sys.stderr))
        d.update(name=wl.name, seed=wl.seed, genome_len=wl.genome_len, read_len=wl.read_len, coverage=wl.coverage,
                 err_ppm=wl.err_ppm, k=wl.k, reads=wl.n_reads, windows=wl.n_windows)
        doc["workloads"][name] = d
        print(name, d, f"{time.perf_counter() - t0:.1f} s", file=sys.stderr)
        json.dump(doc, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
