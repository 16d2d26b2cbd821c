"""The placement hashes (common.cuh place_hash / slot_hash, mirrored in katome_b200/hashing.py) only
decide where a key lives, but bad ones overload sub-tables and pages.  Uniformity on structured key sets."""
import numpy as np

from katome_b200 import hashing


def _spread(keys, n_sub=245, sub_log2=14, page_log2=10, k=31):
    hi = np.zeros_like(keys)
    ph, sh = hashing.place_hash(hi, keys, k), hashing.slot_hash(hi, keys, k)
    part = ((ph * np.uint64(n_sub)) >> np.uint64(32)).astype(np.int64)
    slot = (sh & np.uint64((1 << sub_log2) - 1)).astype(np.int64)
    page = part * (1 << (sub_log2 - page_log2)) + (slot >> page_log2)
    out = []
    for v, m in ((part, n_sub), (slot, 1 << sub_log2), (page, n_sub << (sub_log2 - page_log2))):
        c = np.bincount(v, minlength=m)
        out.append(c.std() / np.sqrt(c.mean()))  # 1.0 for a uniformly random placement
    sampled = (ph & np.uint64(511)) == 0  # the cardinality sketch samples 1 key in 512
    return out, float(sampled.mean() * 512)


def test_placement_is_uniform_on_structured_keys():
    n = 1_000_000
    i = np.arange(n, dtype=np.uint64)
    rng = np.random.default_rng(2)
    pos, val = rng.integers(0, 31, (n, 3)), rng.integers(1, 4, (n, 3))
    low = np.zeros(n, dtype=np.uint64)
    for j in range(3):  # k-mers that are poly-A except for three bases
        low |= val[:, j].astype(np.uint64) << (np.uint64(2) * pos[:, j].astype(np.uint64))
    unit = rng.integers(0, 4, 37)
    tandem = np.array([sum(int(unit[(s + j) % 37]) << (2 * (30 - j)) for j in range(31)) for s in range(37)], dtype=np.uint64)
    sets = {"sequential": i, "shifted 32": i << np.uint64(32), "shifted 40": i << np.uint64(40),
            "stride 4^8": i * np.uint64(65536), "stride 3 << 20": (i * np.uint64(3)) << np.uint64(20),
            "low complexity": np.unique(low),
            "tandem repeat + point changes": np.unique((tandem[:, None] ^ (np.uint64(1) << (i[:20000] % np.uint64(62)))[None, :]).ravel())}
    for name, keys in sets.items():
        spread, sample = _spread(keys)
        for r in spread:
            assert r < 1.25, (name, spread)
        assert 0.85 < sample < 1.15, (name, sample)


def test_owner_balance_u128():
    rng = np.random.default_rng(5)
    hi = rng.integers(0, 2**62, 200_000, dtype=np.uint64)
    lo = rng.integers(0, 2**63, 200_000, dtype=np.uint64)
    own = hashing.owner_of(hi, lo, 63, 8, True)
    c = np.bincount(own, minlength=8)
    assert c.min() > 0.95 * c.mean() and c.max() < 1.05 * c.mean()
    # place and slot of u128 keys are independent: pages of one sub-table fill evenly
    ph, sh = hashing.place_hash(hi, lo, 63), hashing.slot_hash(hi, lo, 63)
    part = ((ph * np.uint64(16)) >> np.uint64(32)).astype(np.int64)
    page = part * 64 + (sh & np.uint64(63)).astype(np.int64)
    c = np.bincount(page, minlength=1024)
    assert c.std() / np.sqrt(c.mean()) < 1.25
