// superkmer.cuh -- the multi-GPU exchange in super-k-mers ("minimizer-based
// ownership (route super-k-mers) cuts NVLink bytes").
//
// The key exchange of builder.cuh (mg_scatter_reads) ships one 8-byte canonical k-mer per
// read window to hash(k-mer)'s owner and is NVLink bound from 4 GPUs on.  Here the owner of a
// k-mer is a function of its MINIMIZER (the smallest hashed canonical m-mer among the
// k-m+1 = 16 m-mers inside the k-mer, m = k-15): consecutive windows of a read mostly share
// their minimizer, so a run of n <= 16 consecutive windows travels as ONE 16-byte record
// holding its n+k-1 bases (2 bits each) -- about 2.8 bytes per window instead of 8.  The
// owner unrolls the records back into canonical k-mers while it partitions them by sub-table
// (what it did anyway with the received keys), so nothing downstream changes: the table is
// still keyed and placed by the k-mer itself; only WHICH shard holds a k-mer differs, and
// shards are disjoint either way (a k-mer and its reverse complement contain the same
// canonical m-mers, hence the same minimizer and the same owner).
//
// It replaces, like the key exchange, the window loop of add_read_fastaq
// (/root/reference/src/katome/collections/girs/hm_gir.rs:39-87) on the sending side and
// compress_kmer_with_rev_compl (compress.rs:34-48) on the owner's side.
//
// Record (u128, first base in bits 127:126 like the flat read stream):
//   bits 127 .. 128-2*(n+k-1)   the bases of the run, MSB first; zero below
//   bits 11..8                  owner rank (only used by the sender's own scatter)
//   bits 5..0                   n - 1
// All-ones is never a record (n would be 64): it is the filler of padded NVLink runs.
//
// Supported for 23 <= k <= 31 (m = 8..16: the m-mer fits 32 bits and there are enough distinct
// minimizers to balance 8 owners); other k keep the key exchange.
#pragma once
#include "common.cuh"

namespace ktg {

constexpr uint32_t SKM_W = 16;            // m-mers per k-mer == windows per work item
constexpr uint32_t SKM_K_MIN = 23, SKM_K_MAX = 31;
constexpr uint32_t SKM_OWNER_SHIFT = 8;

__host__ __device__ __forceinline__ bool skm_supported(uint32_t k) { return k >= SKM_K_MIN && k <= SKM_K_MAX; }
__host__ __device__ __forceinline__ uint32_t skm_m(uint32_t k) { return k - (SKM_W - 1); }

// order of the m-mers: any fixed function of the canonical m-mer will do; two rounds so that
// overlapping m-mers (x and 4x+b) do not get correlated ranks, and poly-A is not the minimum
__host__ __device__ __forceinline__ uint32_t skm_hash(uint32_t c) {
    uint32_t h = (c ^ 0x5BD1E995u) * 0x9E3779B1u;
    h ^= h >> 15;
    return h * 0x85EBCA6Bu;
}
__host__ __device__ __forceinline__ uint32_t skm_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
// owner of a k-mer from its minimizer's hash (minimizers are the SMALL hashes: re-mix)
__host__ __device__ __forceinline__ uint32_t skm_owner(uint32_t mz, uint32_t world) {
    uint32_t x = mz * 0xC2B2AE35u;
    x ^= x >> 16;
    x *= 0x9E3779B1u;
    return (uint32_t)(((uint64_t)x * world) >> 32);
}

// y0:y1 = 64 bases starting at the item's first window (first base in bit 63 of y0).
// mz[j] = minimizer hash of window j (bases j .. j+k-1), j = 0..15: the minimum over the
// m-mers starting at j .. j+15.  31 hashes per 16 windows: suffix minima over m-mers 0..15,
// running prefix minimum over m-mers 16..30.
__host__ __device__ __forceinline__ void skm_minimizers(uint64_t y0, uint64_t y1, uint32_t m, uint32_t (&mz)[SKM_W]) {
    const uint32_t mask = m == 16 ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
    const uint32_t rsh = 2 * m - 2;
    uint32_t f = (uint32_t)(y0 >> (64 - 2 * m));
    uint32_t r = (uint32_t)revcomp((uint64_t)f, m);
    const uint64_t z = (y0 << (2 * m)) | (y1 >> (64 - 2 * m)); // bases m .. m+31
    const uint32_t zh = (uint32_t)(z >> 32), zl = (uint32_t)z;
    uint32_t H[SKM_W];
    H[0] = skm_hash(skm_min(f, r));
#pragma unroll
    for (int t = 0; t < 15; ++t) {
        const uint32_t b = (zh >> (30 - 2 * t)) & 3u;
        f = ((f << 2) | b) & mask;
        r = (r >> 2) | ((b ^ 3u) << rsh);
        H[t + 1] = skm_hash(skm_min(f, r));
    }
#pragma unroll
    for (int i = 14; i >= 0; --i) H[i] = skm_min(H[i], H[i + 1]);
    mz[0] = H[0];
    uint32_t p = 0xFFFFFFFFu;
#pragma unroll
    for (int t = 15; t < 30; ++t) {
        const uint32_t b = t < 16 ? (zh >> (30 - 2 * t)) & 3u : (zl >> (30 - 2 * (t - 16))) & 3u;
        f = ((f << 2) | b) & mask;
        r = (r >> 2) | ((b ^ 3u) << rsh);
        p = skm_min(p, skm_hash(skm_min(f, r)));
        mz[t - 14] = skm_min(H[t - 14], p);
    }
}

// minimizer hash of ONE k-mer given as an integer (the key-level twin of the above: the owner
// of a k-mer must not depend on the read it was seen in)
__host__ __device__ __forceinline__ uint32_t skm_minimizer_of_kmer(uint64_t kmer, uint32_t k) {
    const uint32_t m = skm_m(k);
    const uint64_t mask = (1ull << (2 * m)) - 1ull;
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < SKM_W; ++i) {
        const uint32_t f = (uint32_t)((kmer >> (2 * (k - m - i))) & mask);
        const uint32_t r = (uint32_t)revcomp((uint64_t)f, m);
        best = skm_min(best, skm_hash(skm_min(f, r)));
    }
    return best;
}

// bit j of the result: a run (super-k-mer) starts at window j.  valid: bit j = window j exists.
// Runs never continue across work items, so every item is self-contained.
__host__ __device__ __forceinline__ uint32_t skm_run_starts(const uint32_t (&mz)[SKM_W], uint32_t valid) {
    uint32_t diff = 1u;
#pragma unroll
    for (int j = 1; j < (int)SKM_W; ++j) diff |= (uint32_t)(mz[j] != mz[j - 1]) << j;
    return valid & (diff | ~(valid << 1));
}
// owners of the 16 windows, 4 bits each
__host__ __device__ __forceinline__ uint64_t skm_owner_pack(const uint32_t (&mz)[SKM_W], uint32_t world) {
    uint64_t pack = 0;
#pragma unroll
    for (int j = 0; j < (int)SKM_W; ++j) pack |= (uint64_t)skm_owner(mz[j], world) << (4 * j);
    return pack;
}
// length of the run that starts at window j
__host__ __device__ __forceinline__ uint32_t skm_run_length(uint32_t starts, uint32_t valid, uint32_t j) {
    const uint32_t cont = (valid & ~starts) >> (j + 1); // following windows that continue a run
    const uint32_t stop = ~cont;                        // bit 31-j.. are ones: never zero
#ifdef __CUDA_ARCH__
    return 1u + (uint32_t)(__ffs((int)stop) - 1);
#else
    return 1u + (uint32_t)__builtin_ctz(stop);
#endif
}

__host__ __device__ __forceinline__ u128 skm_make_record(uint64_t y0, uint64_t y1, uint32_t j, uint32_t n,
                                                         uint32_t k, uint32_t owner) {
    u128 v = (((u128)y0 << 64) | y1) << (2 * j);
    const uint32_t span = n + k - 1; // <= 46 bases
    v &= ~((((u128)1) << (128 - 2 * span)) - 1);
    return v | ((u128)owner << SKM_OWNER_SHIFT) | (u128)(n - 1);
}
__host__ __device__ __forceinline__ uint32_t skm_record_count(u128 rec) { return ((uint32_t)rec & 63u) + 1u; }
__host__ __device__ __forceinline__ uint32_t skm_record_owner_field(u128 rec) {
    return ((uint32_t)rec >> SKM_OWNER_SHIFT) & 15u;
}
struct SkmRecordOwner { // the bin of a record in the sender's tile scatter
    __device__ __forceinline__ uint32_t operator()(u128 rec) const { return skm_record_owner_field(rec); }
};
// the j-th k-mer of a record, forward strand
__host__ __device__ __forceinline__ uint64_t skm_record_kmer(u128 rec, uint32_t k, uint32_t j) {
    return (uint64_t)(rec >> (128 - 2 * (k + j))) & ((1ull << (2 * k)) - 1ull);
}
// owner of a record, recomputed from its first k-mer (spill route)
__host__ __device__ __forceinline__ uint32_t skm_record_owner(u128 rec, uint32_t k, uint32_t world) {
    return skm_owner(skm_minimizer_of_kmer(skm_record_kmer(rec, k, 0), k), world);
}

// ---- the per-item logic of the sending kernel, restated for the host so that the bit
// twiddling is testable without a GPU (tests/test_superkmer_host.py through
// ktg_debug_skm_host).  packed: flat 2-bit stream (first base of word j in bits 63:62),
// pos: flat index of the item's first window, valid: its window mask.
// Returns the number of records written to out[<= 16].
inline uint32_t skm_item_host(const uint64_t *packed, uint64_t pos, uint32_t valid, uint32_t k, uint32_t world,
                              u128 *out) {
    const uint64_t w = pos >> 5;
    const uint32_t s = (uint32_t)pos & 31u;
    uint64_t x0 = packed[w], x1 = packed[w + 1], x2 = packed[w + 2];
    uint64_t y0 = x0, y1 = x1;
    if (s) {
        y0 = (x0 << (2 * s)) | (x1 >> (64 - 2 * s));
        y1 = (x1 << (2 * s)) | (x2 >> (64 - 2 * s));
    }
    uint32_t mz[SKM_W];
    skm_minimizers(y0, y1, skm_m(k), mz);
    uint32_t starts = skm_run_starts(mz, valid);
    const uint64_t owners = skm_owner_pack(mz, world);
    uint32_t n_out = 0;
    const uint32_t all = starts;
    while (starts) {
        const uint32_t j = (uint32_t)__builtin_ctz(starts);
        starts &= starts - 1;
        const uint32_t n = skm_run_length(all, valid, j);
        out[n_out++] = skm_make_record(y0, y1, j, n, k, (uint32_t)(owners >> (4 * j)) & 15u);
    }
    return n_out;
}

#ifdef __CUDACC__

// block-wide exclusive scan of one u32 per thread; all threads call it
template <int THREADS> __device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t &total) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_t;
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += x;
    }
    if (lane == 31) s_w[wid] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t w = threadIdx.x < THREADS / 32 ? s_w[threadIdx.x] : 0u;
        uint32_t iw = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, iw, d);
            if (threadIdx.x >= (uint32_t)d) iw += x;
        }
        s_w[threadIdx.x] = iw - w;
        if (threadIdx.x == 31) s_t = iw;
    }
    __syncthreads();
    const uint32_t res = s_w[wid] + incl - v;
    total = s_t;
    __syncthreads(); // s_w / s_t may be rewritten by the next call
    return res;
}

// ---- sender: packed reads -> super-k-mer records, grouped by owner, written straight into
// the owners' receive buckets (PeerOut, NVLink) through the same tile scatter as the key
// exchange (records are its 16-byte "keys", bins are owner ranks).
// A work item is 16 consecutive window starts of one read; a lane takes SKM_IPL items per tile.
// The records of a tile (2.8 per item on random sequence, 16 at worst) are compacted into
// shared memory in rounds of SCATTER_TILE and scattered from there.
struct SkmView {
    ReadView v;        // packed / valid / wstart / n_words / shift0 / ulen of the batch
    uint64_t n_items;  // uniform: n_reads * ipr; ragged: 2 * n_words
    uint32_t ipr;      // uniform: ceil(windows per read / 16)
};

constexpr int SKM_IPL = 2; // work items per lane and tile: 8192 windows per tile, ~1500 records per scatter
__global__ void __launch_bounds__(SCATTER_THREADS, 3)
scatter_superkmers_kernel(SkmView sv, uint32_t k, uint32_t world, ScatterOut o, PeerOut po,
                          unsigned long long *__restrict__ key_counts) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ uint32_t s_kc[MAX_P2P_WORLD];
    ScatterSmem<u128, SCATTER_TILE> sm;
    sm.carve(smem, world, false);
    u128 *stage = sm.keys; // dead between two tile_scatter calls
    for (uint32_t i = threadIdx.x; i < 2 * world; i += SCATTER_THREADS) sm.cnt[i] = 0;
    if (threadIdx.x < MAX_P2P_WORLD) s_kc[threadIdx.x] = 0;
    __syncthreads();
    const ReadView &v = sv.v;
    constexpr uint64_t TILE_ITEMS = (uint64_t)SCATTER_THREADS * SKM_IPL;
    const uint64_t n_tiles = (sv.n_items + TILE_ITEMS - 1) / TILE_ITEMS;
    uint32_t parity = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t y0[SKM_IPL], y1[SKM_IPL], owners[SKM_IPL];
        uint32_t starts[SKM_IPL], valid[SKM_IPL];
        uint32_t nrec = 0;
#pragma unroll
        for (int a = 0; a < SKM_IPL; ++a) {
            const uint64_t item = tile * TILE_ITEMS + (uint64_t)a * SCATTER_THREADS + threadIdx.x;
            uint64_t pos = 0;
            valid[a] = 0;
            if (item < sv.n_items) {
                if (v.ulen) {
                    const uint64_t rd = item / sv.ipr;
                    const uint32_t u = (uint32_t)(item - rd * sv.ipr);
                    const uint32_t total = v.ulen - k + 1, done = SKM_W * u;
                    const uint32_t nwin = total - done < SKM_W ? total - done : SKM_W;
                    pos = rd * v.ulen + done + v.shift0;
                    valid[a] = v.valid[rd] ? (1u << nwin) - 1u : 0u;
                }
                else {
                    const uint64_t w = item >> 1;
                    const uint32_t half = (uint32_t)item & 1u;
                    pos = w * 32 + SKM_W * half;
                    valid[a] = (v.wstart[w] >> (SKM_W * half)) & 0xFFFFu;
                }
            }
            y0[a] = y1[a] = owners[a] = 0;
            starts[a] = 0;
            if (valid[a]) {
                const uint64_t w = pos >> 5;
                const uint32_t s = (uint32_t)pos & 31u;
                const uint64_t x0 = v.packed[w], x1 = v.packed[w + 1], x2 = v.packed[w + 2]; // padded by 4 words
                y0[a] = x0;
                y1[a] = x1;
                if (s) {
                    y0[a] = (x0 << (2 * s)) | (x1 >> (64 - 2 * s));
                    y1[a] = (x1 << (2 * s)) | (x2 >> (64 - 2 * s));
                }
                uint32_t mz[SKM_W];
                skm_minimizers(y0[a], y1[a], skm_m(k), mz);
                starts[a] = skm_run_starts(mz, valid[a]);
                owners[a] = skm_owner_pack(mz, world);
            }
            nrec += __popc(starts[a]);
        }
        uint32_t total = 0;
        const uint32_t roff = block_exscan<SCATTER_THREADS>(nrec, total);
        for (uint32_t lo = 0; lo < total; lo += SCATTER_TILE) {
            {   // this round's records of the lane go to the stage, in lane order
                uint32_t idx = roff;
#pragma unroll
                for (int a = 0; a < SKM_IPL; ++a) {
                    uint32_t rest = starts[a];
                    while (rest) {
                        const uint32_t j = (uint32_t)__ffs((int)rest) - 1u;
                        rest &= rest - 1u;
                        if (idx >= lo && idx < lo + SCATTER_TILE) {
                            const uint32_t n = skm_run_length(starts[a], valid[a], j);
                            const uint32_t own = (uint32_t)(owners[a] >> (4 * j)) & 15u;
                            stage[idx - lo] = skm_make_record(y0[a], y1[a], j, n, k, own);
                            atomicAdd(&s_kc[own], n);
                        }
                        ++idx;
                    }
                }
            }
            __syncthreads();
            const uint32_t cnt = total - lo < (uint32_t)SCATTER_TILE ? total - lo : (uint32_t)SCATTER_TILE;
            u128 rec[SCATTER_PER];
            uint32_t bin[SCATTER_PER];
            uint32_t vmask = 0;
#pragma unroll
            for (int q = 0; q < SCATTER_PER; ++q) {
                const uint32_t i = q * SCATTER_THREADS + threadIdx.x;
                const bool in = i < cnt;
                rec[q] = in ? stage[i] : (u128)0;
                bin[q] = skm_record_owner_field(rec[q]);
                if (in) vmask |= 1u << q;
            }
            // (tile_scatter's first barrier orders these reads before it rewrites sm.keys)
            tile_scatter<u128, SCATTER_THREADS, SCATTER_PER>(rec, bin, vmask, sm, world, o.cursors, 0, o, parity, SkmRecordOwner{}, &po);
            parity ^= 1u;
        }
    }
    __syncthreads();
    if (threadIdx.x < world && s_kc[threadIdx.x]) atomicAdd(&key_counts[threadIdx.x], (unsigned long long)s_kc[threadIdx.x]);
}

// ---- owner: received records -> canonical k-mers -> level-1 buckets, in ONE kernel.  Bucket q of
// the receive slot holds what source rank q wrote: records [q*cap, min(ends[q], (q+1)*cap)).
// (Round 1 unrolled the records into a flat array in HBM and partitioned that array with a second
// kernel: 8 bytes per key written and read again.)
// Two stages inside the CTA, both in shared memory:
//   A  a lane unrolls ITS record with the rolling fw / rc of the read extraction (a dozen
//      instructions per k-mer) into a dense staging array -- the warp reserves the room of its 32
//      records with one shared atomicAdd.  Records are 1..16 windows long (5.5 on average), so the
//      lanes of a warp finish at different times; that is cheap here because the loop body is, and
//      the k-mers land in shared memory, not in HBM (the balanced scheme -- lane l makes k-mer
//      32 it + l, finds its record with two popcounts and four shuffles -- cost 3x the instructions
//      per k-mer and left the partitioner's tiles 69 % full);
//   B  whenever 2048 k-mers are staged they go through the tile scatter (bins = sub-tables) as one
//      full tile, which also feeds the cardinality sketch.
constexpr int SREC_THREADS = 256, SREC_PER = 8, SREC_TILE = SREC_THREADS * SREC_PER;
constexpr int SREC_STAGE = SREC_TILE + SREC_THREADS * (int)SKM_W; // what is left over + one round of records at their longest
template <bool RC>
__global__ void __launch_bounds__(SREC_THREADS, 3)
scatter_records_kernel(const u128 *__restrict__ rx, const unsigned long long *__restrict__ ends, uint64_t cap,
                       uint32_t n_buckets, uint32_t k, Table<uint64_t> t, ScatterOut o, uint32_t *__restrict__ g_regs) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ unsigned long long s_end[MAX_P2P_WORLD];
    __shared__ uint32_t s_tbase[MAX_P2P_WORLD + 1]; // data tiles before bucket q (tiles past a bucket's fill are not visited)
    __shared__ uint32_t s_count;                    // k-mers in the staging array
    const uint32_t n_bins = t.n_sub;
    ScatterSmem<uint64_t, SREC_TILE> sm;
    sm.carve(smem, n_bins, false);
    uint64_t *stage = (uint64_t *)(smem + ((ScatterSmem<uint64_t, SREC_TILE>::bytes(n_bins, false) + 15) & ~(size_t)15));
    for (uint32_t i = threadIdx.x; i < 2 * n_bins; i += SREC_THREADS) sm.cnt[i] = 0;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t kmask = (1ull << (2 * k)) - 1ull;
    const uint32_t rc_shift = 2 * (k - 1);
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t q = 0; q < n_buckets; ++q) {
            const unsigned long long beg = (unsigned long long)q * cap, lim = beg + cap, fill = ends[q];
            const unsigned long long e = fill < lim ? fill : lim;
            s_end[q] = e;
            s_tbase[q] = run;
            run += e > beg ? (uint32_t)((e - beg + SREC_THREADS - 1) / SREC_THREADS) : 0u;
        }
        s_tbase[n_buckets] = run;
        s_count = 0;
    }
    __syncthreads();
    const uint32_t n_tiles = s_tbase[n_buckets];
    // the records of the NEXT tile are loaded before this one is processed
    auto fetch = [&](uint32_t tile, uint64_t &hi, uint64_t &lo) {
        hi = lo = ~0ull;
        if (tile >= n_tiles) return;
        uint32_t q = 0;
        while (q + 1 < n_buckets && tile >= s_tbase[q + 1]) ++q;
        const uint64_t i = (uint64_t)q * cap + (uint64_t)(tile - s_tbase[q]) * SREC_THREADS + threadIdx.x;
        if (i < s_end[q]) {
            const ulonglong2 raw = __ldcs((const ulonglong2 *)(rx + i));
            lo = raw.x;
            hi = raw.y;
        }
    };
    uint32_t parity = 0;
    // stage B: the last `take` staged k-mers as one tile of the partitioner
    auto drain = [&](uint32_t count) {
        const uint32_t take = count < (uint32_t)SREC_TILE ? count : (uint32_t)SREC_TILE, from = count - take;
        uint64_t key[SREC_PER];
        uint32_t bin[SREC_PER];
        uint32_t vmask = 0, sampled = 0;
#pragma unroll
        for (int q = 0; q < SREC_PER; ++q) {
            const uint32_t i = q * SREC_THREADS + threadIdx.x;
            const bool in = i < take;
            key[q] = in ? stage[from + i] : 0ull;
            const uint32_t ph = KeyTraits<uint64_t>::place_hash(key[q]);
            bin[q] = place_of(ph, t.world, t.n_sub).part;
            if (in) {
                vmask |= 1u << q;
                if (hll_sampled(ph)) sampled |= 1u << q;
            }
        }
        hll_update_tile<uint64_t, SREC_PER>(g_regs, key, sampled);
        tile_scatter<uint64_t, SREC_THREADS, SREC_PER>(key, bin, vmask, sm, n_bins, o.cursors, 0, o, parity,
                                                       BinByPlace<uint64_t, BIN_PART>{t.world, t.n_sub});
        parity ^= 1u;
        if (threadIdx.x == 0) s_count = from;
        __syncthreads();
    };
    uint64_t nhi, nlo;
    fetch(blockIdx.x, nhi, nlo);
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t rhi = nhi, rlo = nlo;
        fetch(tile + gridDim.x, nhi, nlo);
        // stage A
        const uint32_t n = (rhi & rlo) == ~0ull ? 0u : ((uint32_t)rlo & 63u) + 1u; // all-ones: filler of a padded run
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= (uint32_t)d) incl += x;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t base = 0;
        if (lane == 0 && total) base = atomicAdd(&s_count, total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0) + incl - n;
        if (n) {
            // the record's bases, first base in bits 127:126: the first k-mer is its top 2k bits (k <= 31),
            // the up to 15 bases that follow are the next bits
            uint64_t fw = rhi >> (64 - 2 * k);
            uint64_t rest = (rhi << (2 * k)) | (rlo >> (64 - 2 * k));
            uint64_t rc = RC ? revcomp(fw, k) : 0ull;
            for (uint32_t j = 0; j < n; ++j) {
                stage[base + j] = (RC && rc < fw) ? rc : fw;
                const uint64_t b = rest >> 62;
                rest <<= 2;
                fw = ((fw << 2) | b) & kmask;
                if (RC) rc = (rc >> 2) | ((3ull - b) << rc_shift);
            }
        }
        __syncthreads();
        uint32_t staged = s_count; // every thread reads the same value: nobody adds to it before the next barrier
        __syncthreads();
        for (; staged >= (uint32_t)SREC_TILE; staged -= SREC_TILE) drain(staged);
    }
    __syncthreads();
    if (s_count) drain(s_count);
}

// ---- spill route (records that did not fit a receive bucket): group by owner with plain
// atomics -- it is rare and small.
__global__ void count_record_owners_kernel(const u128 *__restrict__ recs, uint64_t n, uint32_t k, uint32_t world,
                                           unsigned long long *__restrict__ counts) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(&counts[skm_record_owner(recs[i], k, world)], 1ull);
}
__global__ void scatter_record_owners_kernel(const u128 *__restrict__ recs, uint64_t n, uint32_t k, uint32_t world,
                                             unsigned long long *__restrict__ cursors, u128 *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u128 r = recs[i];
        out[atomicAdd(&cursors[skm_record_owner(r, k, world)], 1ull)] = r;
    }
}

#endif // __CUDACC__

} // namespace ktg
