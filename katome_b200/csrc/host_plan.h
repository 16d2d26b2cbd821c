// host_plan.h -- the host batcher's plan for one ktg_add_reads call: where the batch is cut into
// chunks and after which chunks the stage is flushed.  Pure host logic (no CUDA), so that it is
// unit-tested on the CPU through ktg_plan_chunks (tests/test_abi.py).
//
// The batch form of the create_fastq loop (algorithms/builder.rs:152-160) hands over many reads at
// once; the batcher copies them to the device in chunks and overlaps the copies with the kernels.
// Only what cannot start before the LAST copy has landed is exposed: the last chunk's kernels, the
// flush of whatever is staged then, the caller's first query.  So a large call (>= 3 chunks)
//   * ends in (up to) four short chunks, together the last 0.5-1.5 chunks of bytes, each with its own
//     staging buffer, all queued on the copy engine as soon as the last full chunk is;
//   * is flushed on the way after the given percentages of its bytes (default: once, after ~55 %; the
//     builder's own cadence, a flush per 0.75 x capacity keys, is held back: every extra sweep of the
//     table is GPU time that queues up behind the copies), never later than after the last full chunk.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ktg {

struct ChunkPlan {
    std::vector<uint64_t> cut;     // chunk c holds reads [cut[c], cut[c+1])
    std::vector<char> flush_here;  // flush the stage after chunk c
    size_t tail_first = (size_t)-1; // first of the short chunks at the end ((size_t)-1: none)
    bool large = false;
    size_t n_chunks() const { return cut.size() - 1; }
};

// offsets[0..n_reads]: absolute, non-decreasing.  A chunk holds the largest run of whole reads that is
// at most `limit` bytes long (at least one read, however long).
inline ChunkPlan plan_chunks(const uint64_t *offsets, uint64_t n_reads, uint64_t chunk_bytes,
                             const std::vector<uint64_t> &flush_pcts, bool taper) {
    ChunkPlan p;
    p.cut.push_back(0);
    const uint64_t total = offsets[n_reads] - offsets[0];
    p.large = taper && total >= 3 * chunk_bytes;
    while (p.cut.back() < n_reads) {
        const uint64_t r = p.cut.back();
        uint64_t limit = chunk_bytes;
        if (p.large) {
            const uint64_t left = offsets[n_reads] - offsets[r];
            if (left <= chunk_bytes + chunk_bytes / 2) {
                if (p.tail_first == (size_t)-1) p.tail_first = p.cut.size() - 1;
                limit = std::max<uint64_t>(left / (4 - std::min<size_t>(3, p.cut.size() - 1 - p.tail_first)), 1u << 20);
            }
        }
        uint64_t lo = r + 1, hi = n_reads;
        while (lo < hi) {
            const uint64_t mid = (lo + hi + 1) / 2;
            if (offsets[mid] - offsets[r] <= limit) lo = mid;
            else hi = mid - 1;
        }
        p.cut.push_back(lo);
    }
    p.flush_here.assign(p.n_chunks(), 0);
    if (p.large) {
        const size_t n = p.n_chunks();
        const size_t last_full = (p.tail_first != (size_t)-1 && p.tail_first > 0) ? p.tail_first - 1 : n - 1;
        for (uint64_t pct : flush_pcts) {
            size_t c = 0;
            while (c + 1 < n && offsets[p.cut[c + 1]] - offsets[0] < total / 100 * pct) ++c;
            p.flush_here[std::min(c, last_full)] = 1;
        }
    }
    return p;
}

} // namespace ktg
