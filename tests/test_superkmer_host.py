"""The record cutter of the super-k-mer exchange (katome_b200/csrc/superkmer.cuh), through its
host-only twin: no GPU needed.  The device kernels run the same __host__ __device__ functions
per work item; tests/test_gpu_parity.py covers the kernels themselves."""
import ctypes as C
from collections import Counter

import numpy as np
import pytest

from katome_b200 import _lib
from tests.helpers import CODE, kmer_int, revcomp

MASK64 = (1 << 64) - 1


def pack_stream(seq: str, pad_words: int = 4) -> np.ndarray:
    """flat 2-bit stream, first base of word j in bits 63:62 (pack_flat_kernel's layout)"""
    n_words = (len(seq) + 31) // 32 + pad_words
    out = np.zeros(n_words, np.uint64)
    for j in range(0, len(seq), 32):
        v = 0
        chunk = seq[j:j + 32]
        for c in chunk:
            v = (v << 2) | CODE[c]
        v <<= 2 * (32 - len(chunk))
        out[j // 32] = v
    return out


def cut(reads, k, world):
    """-> list of (record int, read index, first window) via ktg_skm_items_host"""
    L = _lib.lib()
    flat = "".join(reads)
    packed = pack_stream(flat)
    pos, valid, origin = [], [], []
    base = 0
    for ri, r in enumerate(reads):
        nwin = len(r) - k + 1
        for u in range(0, max(nwin, 0), 16):
            pos.append(base + u)
            valid.append((1 << min(16, nwin - u)) - 1)
            origin.append((ri, u))
        base += len(r)
    pos = np.array(pos, np.uint64)
    valid = np.array(valid, np.uint32)
    cap = int(sum(bin(v).count("1") for v in valid.tolist())) + 1
    out = np.zeros(2 * cap, np.uint64)
    n = C.c_uint64(0)
    rc = L.ktg_skm_items_host(packed.ctypes.data, pos.ctypes.data, valid.ctypes.data, len(pos), k, world,
                              out.ctypes.data, cap, C.byref(n))
    assert rc == 0, L.ktg_last_error()
    recs = [(int(out[2 * i + 1]) << 64) | int(out[2 * i]) for i in range(n.value)]
    return recs


def kmers_of(rec, k):
    n = (rec & 63) + 1
    return [(rec >> (128 - 2 * (k + j))) & ((1 << (2 * k)) - 1) for j in range(n)]


def canon_int(x, k):
    s = "".join("ACGT"[(x >> (2 * (k - 1 - i))) & 3] for i in range(k))
    return min(x, kmer_int(revcomp(s)))


@pytest.mark.parametrize("k", [23, 24, 27, 31])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_records_hold_exactly_the_windows_and_one_owner_each(k, world):
    rng = np.random.default_rng(1000 * k + world)
    genome = "".join(rng.choice(list("ACGT"), size=3000))
    reads = []
    for _ in range(120):
        ln = int(rng.integers(k, 160))
        st = int(rng.integers(0, len(genome) - ln))
        s = genome[st:st + ln]
        reads.append(revcomp(s) if rng.random() < 0.5 else s)
    reads += ["A" * 100, "ACGT" * 30, "AC" * 40, "T" * k, "G" * (k + 17)]  # low complexity, exactly-k
    recs = cut(reads, k, world)
    L = _lib.lib()
    want = Counter()
    for r in reads:
        for i in range(len(r) - k + 1):
            want[kmer_int(r[i:i + k])] += 1
    got = Counter()
    n_windows = 0
    for rec in recs:
        n = (rec & 63) + 1
        assert 1 <= n <= 16
        span = n + k - 1
        assert rec & ((1 << (128 - 2 * span)) - 1) & ~0xFFF == 0  # nothing below the bases but the fields
        owner = (rec >> 8) & 15
        assert owner < world
        ks = kmers_of(rec, k)
        n_windows += n
        for x in ks:
            got[x] += 1
            # the owner is a function of the k-mer alone, and the same for both strands
            assert L.ktg_skm_owner_of_kmer(x, k, world) == owner
            assert L.ktg_skm_owner_of_kmer(canon_int(x, k), k, world) == owner
    assert got == want
    # random sequence: about 2.8 records per 16 windows; anything near 1 record per window
    # would mean the minimizers are not shared
    assert len(recs) < 0.45 * n_windows


def test_owners_are_balanced_on_random_sequence():
    rng = np.random.default_rng(5)
    k, world = 31, 8
    genome = "".join(rng.choice(list("ACGT"), size=60000))
    reads = [genome[i:i + 150] for i in range(0, len(genome) - 150, 75)]
    recs = cut(reads, k, world)
    per_owner = np.zeros(world)
    for rec in recs:
        per_owner[(rec >> 8) & 15] += (rec & 63) + 1
    share = per_owner / per_owner.sum()
    assert share.min() > 0.09 and share.max() < 0.16, share


def test_unsupported_k_is_rejected():
    L = _lib.lib()
    assert L.ktg_mg_skm_supported(31) == 1 and L.ktg_mg_skm_supported(23) == 1
    assert L.ktg_mg_skm_supported(22) == 0 and L.ktg_mg_skm_supported(32) == 0 and L.ktg_mg_skm_supported(63) == 0
    out = np.zeros(4, np.uint64)
    n = C.c_uint64(0)
    z = np.zeros(8, np.uint64)
    v = np.ones(1, np.uint32)
    assert L.ktg_skm_items_host(z.ctypes.data, z.ctypes.data, v.ctypes.data, 1, 40, 2, out.ctypes.data, 2, C.byref(n)) != 0
