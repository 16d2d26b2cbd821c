"""Host logic of the batcher behind ktg_add_reads (katome_b200/csrc/host_plan.h), on the CPU through the
test hook ktg_plan_chunks: where a batch is cut into chunks, the short chunks a large call ends in, and
where the stage is flushed on the way."""
import ctypes as C

import numpy as np
import pytest

from katome_b200 import _lib

MB = 1 << 20


def plan(offsets, chunk_bytes, pcts=(55,)):
    L = _lib.lib()
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n_reads = len(offsets) - 1
    cap = n_reads + 8
    cuts = np.zeros(cap + 1, np.uint64)
    flush = np.zeros(cap, np.uint8)
    p = np.asarray(pcts, dtype=np.uint32)
    n, tail = C.c_uint32(0), C.c_int64(0)
    rc = L.ktg_plan_chunks(offsets.ctypes.data, n_reads, chunk_bytes, p.ctypes.data if len(p) else None, len(p),
                           cuts.ctypes.data, flush.ctypes.data, cap, C.byref(n), C.byref(tail))
    assert rc == 0
    return cuts[: n.value + 1].astype(np.int64), flush[: n.value].astype(bool), tail.value


def check_cover(offsets, cuts, limit_of):
    assert cuts[0] == 0 and cuts[-1] == len(offsets) - 1 and np.all(np.diff(cuts) >= 1)
    for c in range(len(cuts) - 1):
        nbytes = int(offsets[cuts[c + 1]] - offsets[cuts[c]])
        one_read = cuts[c + 1] - cuts[c] == 1
        assert nbytes <= limit_of(c) or one_read, (c, nbytes)
        if cuts[c + 1] < len(offsets) - 1:  # maximal: the next read would not have fitted
            assert int(offsets[cuts[c + 1] + 1] - offsets[cuts[c]]) > limit_of(c), c


def test_small_call_is_cut_at_the_chunk_size_and_never_flushed():
    offsets = np.arange(0, 25001, dtype=np.uint64) * 100  # 2.5 MB of 100 bp reads
    cuts, flush, tail = plan(offsets, MB)
    assert tail == -1 and not flush.any() and len(cuts) - 1 == 3
    check_cover(offsets, cuts, lambda c: MB)
    assert cuts[1] == MB // 100


@pytest.mark.parametrize("n_reads,read_len", [(460_000, 100), (100_000, 150), (31_459, 100), (40_004, 100)])
def test_large_uniform_call_tapers_and_flushes_once(n_reads, read_len):
    offsets = np.arange(0, n_reads + 1, dtype=np.uint64) * read_len + 12345  # absolute offsets
    total = n_reads * read_len
    cuts, flush, tail = plan(offsets, MB)
    n = len(cuts) - 1
    assert tail >= 2 and n - tail <= 4
    # the full chunks hold as many whole reads as fit in 1 MiB
    for c in range(tail):
        assert cuts[c + 1] - cuts[c] == MB // read_len
    # what is left for the tail is the last 0.5 .. 1.5 chunks, in (up to) four roughly equal pieces
    tail_bytes = int(offsets[cuts[-1]] - offsets[cuts[tail]])
    assert MB // 2 - read_len <= tail_bytes <= MB + MB // 2
    sizes = [int(offsets[cuts[c + 1]] - offsets[cuts[c]]) for c in range(tail, n)]
    if tail_bytes >= 4 * MB // 4 * 1:  # pieces of at least 1 MiB / 4 are not floored to the 1 MiB minimum
        assert max(sizes) <= max(tail_bytes // 4 + read_len, MB)
    # one flush, after the first chunk that ends at or beyond 55 % of the bytes, never inside the tail
    assert flush.sum() == 1
    c = int(np.flatnonzero(flush)[0])
    assert c < tail
    done = int(offsets[cuts[c + 1]] - offsets[0])
    before = int(offsets[cuts[c]] - offsets[0])
    assert before < total // 100 * 55 and (done >= total // 100 * 55 or c == tail - 1)


def test_flush_schedule_knob_and_its_clamp():
    offsets = np.arange(0, 100_001, dtype=np.uint64) * 100  # 10 MB -> 9 full chunks of 1 MiB + tail
    cuts, flush, tail = plan(offsets, MB, pcts=(29, 58))
    assert flush.sum() == 2
    ends = [int(offsets[cuts[c + 1]]) for c in np.flatnonzero(flush)]
    assert ends[0] >= 0.29 * 10_000_000 > ends[0] - MB and ends[1] >= 0.58 * 10_000_000 > ends[1] - MB
    # a flush asked for at 100 % happens after the last FULL chunk: the tail chunks are already queued then
    cuts, flush, tail = plan(offsets, MB, pcts=(100,))
    assert flush.sum() == 1 and int(np.flatnonzero(flush)[0]) == tail - 1
    # no percentages: no flush on the way
    cuts, flush, tail = plan(offsets, MB, pcts=())
    assert not flush.any() and tail >= 2


def test_ragged_reads_and_a_read_longer_than_a_chunk():
    rng = np.random.default_rng(5)
    lens = rng.integers(40, 400, size=30_000).astype(np.uint64)
    lens[1234] = 3 * MB  # one read of three chunks: it gets a chunk to itself
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    cuts, flush, tail = plan(offsets, MB)
    total = int(offsets[-1])
    n = len(cuts) - 1

    def limit_of(c):
        if tail < 0 or c < tail:
            return MB
        left = int(offsets[-1] - offsets[cuts[c]])
        return max(left // (4 - min(3, c - tail)), MB)
    check_cover(offsets, cuts, limit_of)
    big = int(np.searchsorted(cuts, 1234, side="right")) - 1
    assert cuts[big] == 1234 and cuts[big + 1] == 1235
    assert total >= 3 * MB and flush.sum() == 1 and n - tail <= 4


def test_item_to_read_without_a_division():
    """The extraction kernels turn an item index into (read, granule) with a multiplication by
    floor(2^64 / items_per_read) + 1 (kernels.cuh: ReadView::set_ipr, div_magic): exact for every 32-bit item."""
    from katome_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(7)
    edge = np.array([0, 1, 2, 3, 2**16 - 1, 2**16, 2**31 - 1, 2**31, 2**32 - 2, 2**32 - 1], dtype=np.uint64)
    for d in [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 63, 64, 65, 100, 255, 256, 1023, 1024, 1250, 4097, 65535, 65536,
              99991, 2**20, 2**20 + 1, 2**31 - 1, 2**31, 2**32 - 1]:
        top = (2**32 - 1) // d * d  # the largest multiple of d, and the ones just below it
        mult = np.array([m for m in (top, top - d, top - 2 * d, d, 2 * d) if m > 0], dtype=np.uint64)
        near = np.concatenate([mult - 1, mult, (mult + 1) % 2**32, edge * d % 2**32, np.arange(0, min(4 * d + 2, 4000), dtype=np.uint64)])
        items = np.concatenate([rng.integers(0, 2**32, size=20000, dtype=np.uint64), edge, near]).astype(np.uint32)
        out = np.empty(len(items), dtype=np.uint32)
        assert L.ktg_item_reads(items.ctypes.data, len(items), d, out.ctypes.data) == 0
        assert np.array_equal(out, (items.astype(np.uint64) // d).astype(np.uint32)), d
