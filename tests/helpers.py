"""Shared test helpers: fixtures on disk, a pure-Python set-level twin of the oracle."""
import os
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
COMP = str.maketrans("ACGT", "TGCA")


def load_seqs(name):
    with open(os.path.join(GOLDEN, f"c1_{name}.seqs")) as f:
        return [line.rstrip("\n") for line in f if line.strip() != "" or True][: None]


def golden_seqs(name):
    with open(os.path.join(GOLDEN, f"c1_{name}.seqs")) as f:
        return f.read().split("\n")[:-1]


def write_fastq(path, seqs, qual_len=None):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            q = "I" * (len(s) if qual_len is None else qual_len)
            f.write(f"@r{i} synthetic header\n{s}\n+\n{q}\n")
    return str(path)


def write_fasta(path, seqs, width=60):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f">r{i}\n")
            for j in range(0, len(s), width):
                f.write(s[j:j + width] + "  \n")  # trailing blanks are trimmed by the reader
    return str(path)


def batch_of(seqs):
    """list of str/bytes -> (bases u8, offsets u64)"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    offsets = np.zeros(len(bs) + 1, np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs])
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy()
    return bases, offsets


def revcomp(s: str) -> str:
    return s.translate(COMP)[::-1]


def kmer_int(s: str) -> int:
    v = 0
    for c in s:
        v = (v << 2) | CODE[c]
    return v


def py_build(seqs, k, rc):
    """SURVEY Appendix A, rules 2-7, in plain Python: -> (Counter edges, accepted_reads, bytes)."""
    w = Counter()
    reads = nbytes = 0
    for s in seqs:
        if any(c not in "ACGT" for c in s):
            continue
        reads += 1
        nbytes += len(s)
        if len(s) < k:
            raise ValueError("Read is too short!")
        for i in range(len(s) - k + 1):
            x = s[i:i + k]
            w[x] += 1
            if rc:
                w[revcomp(x)] += 1
    return w, reads, nbytes


def py_nodes(edges):
    n = set()
    for e in edges:
        n.add(e[:-1])
        n.add(e[1:])
    return n


def py_sorted_arrays(edges: Counter):
    items = sorted((kmer_int(e), wt & 0xFFFFFFFF) for e, wt in edges.items() if wt > 0)
    hi = np.array([v >> 64 for v, _ in items], np.uint64)
    lo = np.array([v & 0xFFFFFFFFFFFFFFFF for v, _ in items], np.uint64)
    w = np.array([x for _, x in items], np.uint32)
    return hi, lo, w


def random_reads(rng, n, lo_len, hi_len, genome=None, n_rate=0.0):
    """reads sampled from a small random genome (so k-mers repeat), some with an 'N'"""
    if genome is None:
        genome = "".join(rng.choice(list("ACGT"), size=2000))
    out = []
    for _ in range(n):
        ln = int(rng.integers(lo_len, hi_len + 1))
        st = int(rng.integers(0, len(genome) - ln + 1))
        s = genome[st:st + ln]
        if rng.random() < 0.5:
            s = revcomp(s)
        if rng.random() < n_rate:
            p = int(rng.integers(0, ln))
            s = s[:p] + rng.choice(list("Nacgt")) + s[p + 1:]
        out.append(s)
    return out
