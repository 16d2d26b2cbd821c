"""The placement hash (common.cuh mix64, mirrored in katome_b200/hashing.py) only decides where a
key lives, but a bad one overloads sub-tables and pages.  Uniformity on structured key sets."""
import numpy as np

from katome_b200 import hashing


def _spread(h, n_sub=245, sub_log2=14, page_log2=10):
    hi, lo = h >> np.uint64(32), h & np.uint64(0xFFFFFFFF)
    part = ((hi * np.uint64(n_sub)) >> np.uint64(32)).astype(np.int64)
    slot = (lo & np.uint64((1 << sub_log2) - 1)).astype(np.int64)
    page = part * (1 << (sub_log2 - page_log2)) + (slot >> page_log2)
    out = []
    for v, m in ((part, n_sub), (slot, 1 << sub_log2), (page, n_sub << (sub_log2 - page_log2))):
        c = np.bincount(v, minlength=m)
        out.append(c.std() / np.sqrt(c.mean()))  # 1.0 for a uniformly random placement
    return out


def test_mix64_is_uniform_on_structured_keys():
    n = 1_000_000
    i = np.arange(n, dtype=np.uint64)
    rng = np.random.default_rng(2)
    pos, val = rng.integers(0, 31, (n, 3)), rng.integers(1, 4, (n, 3))
    low = np.zeros(n, dtype=np.uint64)
    for j in range(3):  # k-mers that are poly-A except for three bases
        low |= val[:, j].astype(np.uint64) << (np.uint64(2) * pos[:, j].astype(np.uint64))
    sets = {"sequential": i, "shifted 32": i << np.uint64(32), "shifted 40": i << np.uint64(40),
            "stride 4^8": i * np.uint64(65536), "stride 3 << 20": (i * np.uint64(3)) << np.uint64(20),
            "low complexity": np.unique(low)}
    for name, keys in sets.items():
        for r in _spread(hashing.mix64(keys)):
            assert r < 1.25, (name, r)


def test_owner_balance_u128():
    rng = np.random.default_rng(5)
    hi = rng.integers(0, 2**62, 200_000, dtype=np.uint64)
    lo = rng.integers(0, 2**63, 200_000, dtype=np.uint64)
    own = hashing.owner_of(hi, lo, 63, 8, True)
    c = np.bincount(own, minlength=8)
    assert c.min() > 0.95 * c.mean() and c.max() < 1.05 * c.mean()
