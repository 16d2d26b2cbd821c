// kernels.cuh -- the hand-written sm_100a kernels of the build stage.
//
//   K1  pack_reads_kernel        ASCII -> 2-bit words + per-word window counts
//   K2  extract (rolling k-mer / reverse complement / canonical) fused into
//   K3  extract_insert_kernel    direct insert into the open-addressing table
//       hist_*/scatter_* kernels partition keys by sub-table (or owner rank)
//       insert_keys_kernel       insert an array of keys (partitioned / received)
//   K4  table scans: stats+digest, filter, standardize, compaction (export)
//
// Reference code each one replaces is cited at the kernel.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace ktg {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint64_t warp_sum(uint64_t v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}
__device__ __forceinline__ uint64_t warp_max(uint64_t v) {
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t x = __shfl_xor_sync(0xFFFFFFFFu, v, o);
        v = x > v ? x : v;
    }
    return v;
}

// block-wide sum of N counters, one atomicAdd per counter per block
template <int N>
__device__ __forceinline__ void block_accumulate(const uint64_t (&v)[N], unsigned long long *out) {
    __shared__ unsigned long long sh[N];
    if (threadIdx.x < N) sh[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        uint64_t s = warp_sum(v[i]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&sh[i], (unsigned long long)s);
    }
    __syncthreads();
    if (threadIdx.x < N && sh[threadIdx.x]) atomicAdd(&out[threadIdx.x], sh[threadIdx.x]);
}

// ======================================================================= K1
// Replaces encode_fasta_symbol (compress.rs:347-378), compress_node
// (compress.rs:55-73) and the whole-read ACGT filter + byte total of
// create_fastq (algorithms/builder.rs:155-158), plus the length assert of
// add_read_fastaq (collections/girs/hm_gir.rs:40).
//
// Layout: the batch is ONE flat 2-bit stream.  Word w holds the bases at byte
// positions [32w, 32w+32) counted from the 32-byte aligned address at or below the
// first read, first base in bits 63:62 (the MSB-first order of compress_node).  A
// second stream has one bit per byte: "not one of A C G T".  Packing is therefore a
// pure streaming kernel over aligned 32-byte chunks, with no per-read control flow;
// read boundaries only matter to check_reads_kernel (validity, counters) and to the
// extraction kernels, which address the stream at 2-bit granularity.
struct PackCounters {
    unsigned long long accepted_reads, accepted_bytes, windows, short_reads;
    unsigned long long min_len, max_len; // over ALL reads of the batch (valid or not)
    unsigned long long shift0;           // flat position of the batch's first base (0..31)
};

// four ASCII bases (little endian: first base in the low byte) -> 8 bits, MSB first;
// bad: bit q set iff byte q is not one of "ACGT"
__device__ __forceinline__ uint32_t pack4(uint32_t x, uint32_t &bad) {
    uint32_t c = ((x >> 1) & 0x03030303u) ^ ((x >> 2) & 0x01010101u);
    uint32_t sel = (c & 0xFu) | ((c >> 4) & 0xF0u) | ((c >> 8) & 0xF00u) | ((c >> 12) & 0xF000u);
    uint32_t d = __byte_perm(0x54474341u, 0u, sel) ^ x; // "ACGT" looked up by code, against the input
    uint32_t nz = (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u; // 0x80 in every non-zero byte
    bad = (((nz >> 7) * 0x00204081u) >> 21) & 0xFu; // bits 0, 8, 16, 24 gathered into a nibble
    return (c * 0x40100401u) >> 24;
}

__global__ void __launch_bounds__(256)
pack_flat_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ offsets,
                 uint64_t total_bases, uint64_t *__restrict__ packed, uint32_t *__restrict__ bad,
                 PackCounters *ctr) {
    const uintptr_t first = (uintptr_t)(bases + offsets[0]);
    const uint4 *src = (const uint4 *)(first & ~(uintptr_t)31);
    const uint32_t shift0 = (uint32_t)(first & 31);
    const uint64_t n_words = (total_bases + shift0 + 31) / 32;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctr->shift0 = shift0;
    for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
        const uint4 v0 = __ldcs(src + 2 * w), v1 = __ldcs(src + 2 * w + 1);
        const uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t hi = 0, lo = 0, bm = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t b0, b1;
            hi = (hi << 8) | pack4(x[q], b0);
            lo = (lo << 8) | pack4(x[q + 4], b1);
            bm |= (b0 << (4 * q)) | (b1 << (4 * q + 16));
        }
        packed[w] = ((uint64_t)hi << 32) | lo;
        bad[w] = bm;
    }
}

// One thread per read: accept / reject (any bad byte rejects the whole read), counters, and
// valid[r] = 1 iff the read contributes windows (accepted and len >= k).
__global__ void __launch_bounds__(1024)
check_reads_kernel(const uint64_t *__restrict__ offsets, uint64_t n_reads,
                   const uint8_t *__restrict__ bases, uint32_t k, const uint32_t *__restrict__ bad,
                   uint8_t *__restrict__ valid, PackCounters *ctr) {
    const uint64_t off0 = offsets[0];
    const uint32_t shift0 = (uint32_t)((uintptr_t)(bases + off0) & 31);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t acc[4] = {0, 0, 0, 0};
    uint64_t lmin = ~0ull, lmax = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        const uint64_t o0 = offsets[r] - off0 + shift0, o1 = offsets[r + 1] - off0 + shift0;
        const uint64_t len = o1 - o0;
        lmin = len < lmin ? len : lmin;
        lmax = len > lmax ? len : lmax;
        uint32_t any = 0;
        if (len) {
            const uint64_t wf = o0 >> 5, wl = (o1 - 1) >> 5;
            for (uint64_t w = wf; w <= wl; ++w) {
                uint32_t m = bad[w];
                if (w == wf) m &= 0xFFFFFFFFu << (o0 & 31);
                if (w == wl) m &= 0xFFFFFFFFu >> (31 - ((o1 - 1) & 31));
                any |= m;
            }
        }
        const bool ok = any == 0;
        const uint64_t nwin = (ok && len >= k) ? len - k + 1 : 0;
        valid[r] = nwin != 0;
        if (ok) {
            acc[0] += 1;
            acc[1] += len;
            acc[2] += nwin;
            acc[3] += len < k;
        }
    }
    block_accumulate<4>(acc, (unsigned long long *)ctr);
    // one atomic per block (one per warp on a single address was a fixed 50 us per launch)
    __shared__ unsigned long long s_min, s_max;
    if (threadIdx.x == 0) {
        s_min = ~0ull;
        s_max = 0;
    }
    __syncthreads();
    lmax = warp_max(lmax);
    lmin = ~warp_max(~lmin);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&s_max, (unsigned long long)lmax);
        atomicMin(&s_min, (unsigned long long)lmin);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_max) atomicMax(&ctr->max_len, s_max);
        atomicMin(&ctr->min_len, s_min);
    }
}

// Ragged batches only: wstart bit i of word w <=> a window starts at flat base 32w + i,
// i.e. the base belongs to an accepted read and the window ends inside it.  (Batches whose
// reads all have one length enumerate their windows in closed form and skip this.)
__global__ void __launch_bounds__(256)
mark_starts_kernel(const uint64_t *__restrict__ offsets, uint64_t n_reads, const uint8_t *__restrict__ bases,
                   uint32_t k, const uint8_t *__restrict__ valid, uint32_t *__restrict__ wstart) {
    const uint64_t off0 = offsets[0];
    const uint32_t shift0 = (uint32_t)((uintptr_t)(bases + off0) & 31);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += stride) {
        if (!valid[r]) continue;
        const uint64_t o0 = offsets[r] - off0 + shift0, o1 = offsets[r + 1] - off0 + shift0;
        const uint64_t last = o1 - k; // last window start (valid[r] implies len >= k)
        const uint64_t wf = o0 >> 5, wl = last >> 5;
        for (uint64_t w = wf; w <= wl; ++w) {
            uint32_t m = 0xFFFFFFFFu;
            if (w == wf) m &= 0xFFFFFFFFu << (o0 & 31);
            if (w == wl) m &= 0xFFFFFFFFu >> (31 - (last & 31));
            if (w == wf || w == wl) atomicOr(&wstart[w], m); // boundary words are shared with neighbours
            else wstart[w] = m;
        }
    }
}

// =================================================================== K2 core
// Rolling (k)-mer extraction over the windows that start in one packed word.
// Replaces the per-window re-packing of compress_kmer[_with_rev_compl]
// (compress.rs:18-48) and reverse_compressed_node (compress.rs:153-169):
//   fw' = ((fw << 2) | b) & mask        rc' = (rc >> 2) | ((3 - b) << 2(k-1))
// A lane owns word w and gets the k-1 overlap from its neighbours' words with
// warp shuffles (only the last lanes of a warp load them).
template <class K> struct Roller {
    K fw, rc, mask;
    uint32_t nxt; // the 16 bases after the first window, left aligned (a work item takes 7 steps)
    uint32_t k, rc_shift;

    __device__ __forceinline__ void init(uint64_t w0, uint64_t w1, uint64_t w2, uint32_t k_) {
        k = k_;
        rc_shift = 2 * (k - 1);
        uint64_t t;
        if (sizeof(K) == 8) {
            mask = k == 32 ? (K)~0ull : (K)((1ull << (2 * k)) - 1);
            fw = (K)(w0 >> (64 - 2 * k));
            t = k == 32 ? w1 : ((w0 << (2 * k)) | (w1 >> (64 - 2 * k)));
        }
        else {
            mask = k == 64 ? ~(K)0 : (K)((((u128)1) << (2 * k)) - 1);
            fw = (K)((((u128)w0 << 64) | w1) >> (128 - 2 * k));
            t = k == 64 ? w2 : ((w1 << (2 * k - 64)) | (w2 >> (128 - 2 * k)));
        }
        nxt = (uint32_t)(t >> 32);
        rc = revcomp(fw, k);
    }
    __device__ __forceinline__ void step() {
        const uint32_t b = nxt >> 30;
        nxt <<= 2;
        fw = ((fw << 2) | (K)b) & mask;
        rc = (rc >> 2) | ((K)(3u - b) << rc_shift);
    }
};

// ---- mbarrier / bulk-copy PTX (sm_90+; the 1-D flavour of TMA needs no tensor map)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// orders this thread's generic-proxy writes to shared memory before later async-proxy (bulk copy) reads
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ============================================================ cardinality
// HyperLogLog sketch (2^12 registers) of the keys offered to the table, merged
// into a persistent sketch with atomicMax.  The host sizes / grows the table
// from the estimate, so no capacity hint is needed (create_fastq starts from
// T::default(), builder.rs:145) and no pessimistic "every window is new" bound
// is used.  Only keys whose hash falls in a fixed 1/512 of the hash space are
// sketched (consistent for duplicates, so distinct(sample) * 512 estimates
// distinct(all); at the smallest table that ever needs an estimate, 1 Mi slots,
// that is still ~700 samples, and an underestimate only means a replayed overflow);
// the sampled keys are re-mixed so that register index and rank are independent of
// the bits that place the key in the table.  The warp vote turns the update into a
// real branch that most warps skip (as straight-line predicated code it was 20 % of
// the scatter kernel's instructions).
constexpr uint32_t HLL_P = 12, HLL_M = 1u << HLL_P, HLL_SAMPLE = 512;

// g: 64 well mixed bits of the key (KeyTraits::hash64), independent of the bits that place it
__device__ __forceinline__ void hll_insert(uint32_t *regs, uint64_t g) {
    uint32_t idx = (uint32_t)(g >> (64 - HLL_P));
    uint64_t rest = g << HLL_P;
    uint32_t rank = rest ? (uint32_t)__clzll((long long)rest) + 1u : 64u - HLL_P + 1u;
    if (regs[idx] < rank) atomicMax(&regs[idx], rank);
}
__global__ void hll_merge_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < HLL_M && src[i] > dst[i]) atomicMax(&dst[i], src[i]);
}
// ph: the key's place_hash; owner and sub-table come from its high bits, the sample from its low 9
__device__ __forceinline__ bool hll_sampled(uint32_t ph) { return (ph & (HLL_SAMPLE - 1u)) == 0; }
template <class K> __device__ __forceinline__ void hll_update(uint32_t *regs, K key, uint32_t ph, bool valid = true) {
    const bool s = valid && hll_sampled(ph);
    if (__any_sync(__activemask(), s)) {
        if (s) hll_insert(regs, KeyTraits<K>::hash64(key));
    }
}
// one vote for the PER keys a lane handles in a tile; mask bit j = key j is sampled
template <class K, int PER>
__device__ __forceinline__ void hll_update_tile(uint32_t *regs, const K (&key)[PER], uint32_t mask) {
    if (__any_sync(0xFFFFFFFFu, mask != 0)) {
#pragma unroll
        for (int j = 0; j < PER; ++j)
            if (mask & (1u << j)) hll_insert(regs, KeyTraits<K>::hash64(key[j]));
    }
}

// ================================================================ work items
// The extraction kernels hand every lane a work item of 8 consecutive window start
// positions of the flat stream (instead of one lane per 32-base word: a 100 bp read has
// 70 windows = 32+32+6+0 per word, i.e. 55 % lane utilisation per word but 97 % per
// 8-window item).
//   uniform batch (all reads have length ulen; the usual case for short-read sequencers):
//       items are enumerated in closed form, read r = item / ipr, granule u = item % ipr
//       with ipr = ceil((ulen-k+1)/8), so no lane is handed an empty item;
//   ragged batch: item = (flat word, granule of 8 positions) and the window starts are the
//       set bits of wstart; a warp covers 8 consecutive words, lanes 0..10 load them once
//       and every lane picks its words with warp shuffles.
constexpr int GRAN = 8, ITEMS_PER_WORD = 32 / GRAN;

struct ReadView {
    const uint64_t *packed; // flat 2-bit stream (+4 words of padding)
    const uint8_t *valid;   // uniform: valid[r]
    const uint32_t *wstart; // ragged: window start bits
    uint64_t n_words;
    uint64_t n_items;
    uint32_t shift0; // flat position of the first base of the batch
    uint32_t ulen;   // 0: ragged
    uint32_t ipr;    // items per read (uniform)
    uint64_t ipr_magic; // floor(2^64 / ipr) + 1, or 0 when ipr == 1: item / ipr without a division
    __host__ void set_ipr(uint32_t v) {
        ipr = v;
        // (2^64 - 1) / v == floor(2^64 / v) unless v is a power of two, where it is one less
        ipr_magic = v > 1 ? ~0ull / v + ((v & (v - 1)) == 0 ? 2 : 1) : 0;
    }
};

// n / d for n < 2^32 and d < 2^32 with m = floor(2^64 / d) + 1: hi64(n * m) (exact because n * d < 2^64);
// a 32-bit division is ~20 instructions, this is two wide multiplies
__host__ __device__ __forceinline__ uint32_t div_magic(uint32_t n, uint64_t m) {
    const uint64_t lo = (uint64_t)n * (uint32_t)m;
    const uint64_t hi = (uint64_t)n * (uint32_t)(m >> 32) + (lo >> 32);
    return (uint32_t)(hi >> 32);
}

template <class K> struct ItemWindows {
    Roller<K> r;
    uint32_t mask; // bit j: the j-th position of the item starts a window
    // all 32 lanes of the warp must call this with consecutive items
    __device__ __forceinline__ void load(const ReadView &v, uint64_t item, uint32_t k) {
        constexpr bool WIDE = sizeof(K) == 16;
        uint64_t w0, w1, w2, w3 = 0;
        uint32_t s;
        if (v.ulen) {
            uint64_t rd, pos;
            uint32_t u;
            if (v.n_items <= 0xFFFFFFFFull) {
                rd = v.ipr_magic ? div_magic((uint32_t)item, v.ipr_magic) : (uint32_t)item;
                u = (uint32_t)item - (uint32_t)rd * v.ipr;
            }
            else {
                rd = item / v.ipr;
                u = (uint32_t)(item - rd * v.ipr);
            }
            const bool in = item < v.n_items;
            pos = rd * v.ulen + GRAN * u + v.shift0;
            const uint64_t w = pos >> 5;
            s = (uint32_t)pos & 31;
            const uint32_t total = v.ulen - k + 1, done = GRAN * u;
            const uint32_t nwin = total - done < GRAN ? total - done : GRAN;
            mask = (in && v.valid[rd]) ? (1u << nwin) - 1u : 0u;
            w0 = in ? v.packed[w] : 0;
            w1 = in ? v.packed[w + 1] : 0; // the packed buffer is padded by 4 words
            w2 = in ? v.packed[w + 2] : 0;
            if (WIDE) w3 = in ? v.packed[w + 3] : 0;
        }
        else {
            const int lane = threadIdx.x & 31;
            const uint64_t warp_word = (item - lane) / ITEMS_PER_WORD; // first word of the warp
            const uint64_t lw = warp_word + lane;
            uint64_t x = (lane < 8 + 3 && lw < v.n_words) ? v.packed[lw] : 0;
            const int src = lane / ITEMS_PER_WORD;
            w0 = __shfl_sync(0xFFFFFFFFu, x, src);
            w1 = __shfl_sync(0xFFFFFFFFu, x, src + 1);
            w2 = __shfl_sync(0xFFFFFFFFu, x, src + 2);
            if (WIDE) w3 = __shfl_sync(0xFFFFFFFFu, x, src + 3);
            const uint64_t w = item / ITEMS_PER_WORD;
            const uint32_t g = (uint32_t)(item % ITEMS_PER_WORD);
            s = GRAN * g;
            mask = w < v.n_words ? (v.wstart[w] >> s) & 0xFFu : 0u;
        }
        if (s) { // left-align the span at the item's first base
            const uint32_t sh = 2 * s;
            w0 = (w0 << sh) | (w1 >> (64 - sh));
            w1 = (w1 << sh) | (w2 >> (64 - sh));
            w2 = (w2 << sh) | (WIDE ? (w3 >> (64 - sh)) : 0);
        }
        r.init(w0, w1, w2, k);
    }
    template <bool RC> __device__ __forceinline__ K key() const {
        return (RC && r.rc < r.fw) ? r.rc : r.fw;
    }
    // asks L2 for the packed words a later load(v, item, k) of a uniform batch will read.  (Loading them
    // into registers one tile ahead instead was slower: 1.85 against 1.73 ms on C2, the kernel sits at
    // its 64-register cap.)
    static __device__ __forceinline__ void prefetch(const ReadView &v, uint64_t item) {
        if (!v.ulen || item >= v.n_items || v.n_items > 0xFFFFFFFFull) return;
        const uint32_t rd = v.ipr_magic ? div_magic((uint32_t)item, v.ipr_magic) : (uint32_t)item;
        const uint32_t u = (uint32_t)item - rd * v.ipr;
        const uint64_t pos = (uint64_t)rd * v.ulen + GRAN * u + v.shift0;
        const uint64_t *p = v.packed + (pos >> 5);
        prefetch_l2(p);
        prefetch_l2(p + (sizeof(K) == 16 ? 3 : 2)); // the span may cross into the next 32-byte sector
    }
};

// ======================================================================= K3 (direct)
// Fused extract + insert with L2 atomics (the single-GPU path while the whole table is
// L2 sized): the loop of add_read_fastaq (hm_gir.rs:55-85) with both-strand insertion
// folded into one canonical update: key = min(fw, rc), +2 if fw == rc (a palindrome is
// inserted twice by hm_gir.rs:55-74).
template <class K, bool RC>
__global__ void __launch_bounds__(256)
extract_insert_kernel(ReadView v, uint32_t k, Table<K> t) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_items = (v.n_items + 31) & ~(uint64_t)31;
    for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
        ItemWindows<K> iw;
        iw.load(v, it, k);
#pragma unroll
        for (int j = 0; j < GRAN; ++j) {
            if (iw.mask & (1u << j)) {
                K key = iw.r.fw;
                uint32_t inc = 1;
                if (RC) {
                    if (iw.r.rc < key) key = iw.r.rc;
                    if (iw.r.rc == iw.r.fw) inc = 2;
                }
                table_add(t, key, inc);
            }
            iw.r.step();
        }
    }
}

// stand-alone sketch pass for the direct (unpartitioned) path, only run when
// the trivial bound cannot prove that the batch fits
template <class K, bool RC>
__global__ void __launch_bounds__(256)
hll_reads_kernel(ReadView v, uint32_t k, uint32_t *__restrict__ g_regs) {
    __shared__ uint32_t regs[HLL_M];
    for (uint32_t i = threadIdx.x; i < HLL_M; i += blockDim.x) regs[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_items = (v.n_items + 31) & ~(uint64_t)31;
    for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
        ItemWindows<K> iw;
        iw.load(v, it, k);
#pragma unroll
        for (int j = 0; j < GRAN; ++j) {
            {
                const K key = iw.template key<RC>();
                hll_update(regs, key, KeyTraits<K>::place_hash(key), (iw.mask >> j) & 1u);
            }
            iw.r.step();
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < HLL_M; i += blockDim.x)
        if (regs[i]) atomicMax(&g_regs[i], regs[i]);
}

template <class K>
__global__ void __launch_bounds__(256)
hll_keys_kernel(const K *__restrict__ keys, uint64_t n, uint32_t *__restrict__ g_regs) {
    __shared__ uint32_t regs[HLL_M];
    for (uint32_t i = threadIdx.x; i < HLL_M; i += blockDim.x) regs[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        hll_update(regs, keys[i], KeyTraits<K>::place_hash(keys[i]));
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < HLL_M; i += blockDim.x)
        if (regs[i]) atomicMax(&g_regs[i], regs[i]);
}

// ================================================================ partition
// bin of a key: its owner rank (multi-GPU exchange) or its sub-table.
// histogram of bins over all windows of a packed batch (exact two-pass mode);
// with HLL the same pass also folds the keys into the cardinality sketch.
template <class K, bool RC, bool BY_OWNER, bool HLL>
__global__ void __launch_bounds__(256)
hist_reads_kernel(ReadView v, uint32_t k, Table<K> t, uint32_t n_bins,
                  unsigned long long *__restrict__ g_hist, uint32_t *__restrict__ g_regs) {
    extern __shared__ uint32_t sh_hist[]; // n_bins counters (+ HLL_M registers)
    uint32_t *regs = sh_hist + n_bins;
    for (uint32_t i = threadIdx.x; i < n_bins + (HLL ? HLL_M : 0); i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_items = (v.n_items + 31) & ~(uint64_t)31;
    for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < n_items; it += stride) {
        ItemWindows<K> iw;
        iw.load(v, it, k);
#pragma unroll
        for (int j = 0; j < GRAN; ++j) {
            if (iw.mask & (1u << j)) {
                const K key = iw.template key<RC>();
                const uint32_t h = KeyTraits<K>::place_hash(key);
                Place p = place_of(h, t.world, t.n_sub);
                atomicAdd(&sh_hist[BY_OWNER ? p.owner : p.part], 1u);
                if (HLL) hll_update(regs, key, h);
            }
            iw.r.step();
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&g_hist[i], (unsigned long long)sh_hist[i]);
    if (HLL)
        for (uint32_t i = threadIdx.x; i < HLL_M; i += blockDim.x)
            if (regs[i]) atomicMax(&g_regs[i], regs[i]);
}

template <class K, bool BY_OWNER, bool HLL>
__global__ void __launch_bounds__(256)
hist_keys_kernel(const K *__restrict__ keys, uint64_t n, Table<K> t, uint32_t n_bins,
                 unsigned long long *__restrict__ g_hist, uint32_t *__restrict__ g_regs) {
    extern __shared__ uint32_t sh_hist[];
    uint32_t *regs = sh_hist + n_bins;
    for (uint32_t i = threadIdx.x; i < n_bins + (HLL ? HLL_M : 0); i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t h = KeyTraits<K>::place_hash(keys[i]);
        Place p = place_of(h, t.world, t.n_sub);
        atomicAdd(&sh_hist[BY_OWNER ? p.owner : p.part], 1u);
        if (HLL) hll_update(regs, keys[i], h);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_bins; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&g_hist[i], (unsigned long long)sh_hist[i]);
    if (HLL)
        for (uint32_t i = threadIdx.x; i < HLL_M; i += blockDim.x)
            if (regs[i]) atomicMax(&g_regs[i], regs[i]);
}

// exclusive scan of the bin histogram (n_bins <= a few thousand): one block.
// bucket_cap == 0: cursors = exclusive prefix of hist (exact mode);
// bucket_cap  > 0: cursors = bin * bucket_cap (one-pass mode, hist unused).
__global__ void scan_bins_kernel(const unsigned long long *__restrict__ hist, uint32_t n_bins,
                                 unsigned long long bucket_cap,
                                 unsigned long long *__restrict__ offsets,
                                 unsigned long long *__restrict__ cursors) {
    __shared__ unsigned long long part[1024];
    if (bucket_cap) {
        for (uint32_t i = threadIdx.x; i <= n_bins; i += blockDim.x) {
            offsets[i] = (unsigned long long)i * bucket_cap;
            if (i < n_bins) cursors[i] = (unsigned long long)i * bucket_cap;
        }
        return;
    }
    const uint32_t per = (n_bins + blockDim.x - 1) / blockDim.x;
    const uint32_t b0 = threadIdx.x * per;
    unsigned long long s = 0;
    for (uint32_t i = b0; i < b0 + per && i < n_bins; ++i) s += hist[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (uint32_t i = 0; i < blockDim.x; ++i) {
            unsigned long long v = part[i];
            part[i] = run;
            run += v;
        }
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = b0; i < b0 + per && i < n_bins; ++i) {
        offsets[i] = run;
        cursors[i] = run;
        run += hist[i];
    }
    if (threadIdx.x == blockDim.x - 1) offsets[n_bins] = run;
}

// ---- tile-local counting sort in shared memory, then coalesced runs to HBM.
// A CTA bins one tile of keys in shared memory and reserves one contiguous
// range per (tile, bin) with a single atomicAdd on the bin's cursor, so HBM
// sees runs instead of scattered 8-byte stores.  Cursors are absolute positions
// in the output array.  Two modes share the code:
//   exact    (bucket_cap == 0): cursors start at the exclusive prefix of a
//            histogram pass; bins are dense and contiguous.
//   one-pass (bucket_cap  > 0): bin b owns [(bin_off+b)*cap, (bin_off+b+1)*cap);
//            no histogram pass.  Keys that do not fit (hash skew / heavy
//            hitters) spill to an overflow array that the host inserts
//            separately; nothing is ever dropped silently (the host checks the
//            spill counter).
// The same routine serves both partition levels: level 1 bins by sub-table (or
// owner rank), level 2 bins the keys of one sub-table by page.
// Output positions must fit 32 bits (the host splits larger batches).
// Per-phase cycle counters of the tile scatter (debug builds with -DKTG_PHASE_TIMERS only).
#ifdef KTG_PHASE_TIMERS
__device__ unsigned long long g_phase_cycles[8];
#define KTG_PHASE(i)                                                                  \
    do {                                                                              \
        if (threadIdx.x == 0) {                                                       \
            long long now_ = clock64();                                               \
            atomicAdd(&g_phase_cycles[i], (unsigned long long)(now_ - phase_t0_));    \
            phase_t0_ = now_;                                                         \
        }                                                                             \
    } while (0)
#define KTG_PHASE_BEGIN() long long phase_t0_ = clock64()
#define KTG_PHASE_RESET()                                                             \
    do {                                                                              \
        if (threadIdx.x == 0) phase_t0_ = clock64();                                  \
    } while (0)
#else
#define KTG_PHASE(i)
#define KTG_PHASE_BEGIN()
#define KTG_PHASE_RESET()
#endif

// dynamic shared memory of a tile scatter: the sorted tile | per bin: gaddr 8, spill 8, start 4, cnt 2 x 4 | sketch
__host__ __device__ constexpr size_t scatter_smem_bytes(size_t tile, size_t key_bytes, size_t n_bins, bool hll) {
    return tile * key_bytes + n_bins * 28 + (hll ? (size_t)4096 * 4 : 0);
}
template <class K, int TILE> struct ScatterSmem {
    K *keys;                    // TILE: the tile sorted by bin
    unsigned long long *gaddr;  // n_bins: byte address of the bin's destination minus its start in the sorted tile
    unsigned long long *spill;  // n_bins: reserved start in the overflow array
    uint32_t *start;            // n_bins: start of the bin's run in the sorted tile
    uint32_t *cnt;              // 2 x n_bins (double buffered)
    uint32_t *regs;             // HLL_M (only with HLL)
    __device__ __forceinline__ void carve(unsigned char *base, uint32_t n_bins, bool hll) {
        keys = (K *)base;
        gaddr = (unsigned long long *)(keys + TILE);
        spill = gaddr + n_bins;
        start = (uint32_t *)(spill + n_bins);
        cnt = start + n_bins;
        regs = cnt + 2 * n_bins;
    }
    __host__ __device__ static size_t bytes(uint32_t n_bins, bool hll) { return scatter_smem_bytes(TILE, sizeof(K), n_bins, hll); }
};

// The bin of a key as a function object: the copy-out of the tile scatter recomputes it from the key
// (a dozen ALU instructions) instead of keeping a 16-bit bin per position in shared memory.  Both
// partition levels are bound by shared-memory wavefronts, not by issue slots (level 2: LSU data pipe at
// 78 %, 22 wavefronts per 32 keys of which the randomly placed 16-bit store and its load were 4.5).
template <class K, int BINS> struct BinByPlace { // BIN_PART / BIN_OWNER / BIN_OWNER_PART of place_hash
    uint32_t world, n_sub;
    __device__ __forceinline__ uint32_t operator()(K key) const {
        const Place p = place_of(KeyTraits<K>::place_hash(key), world, n_sub);
        return BINS == 1 ? p.owner : BINS == 0 ? p.part : p.owner * n_sub + p.part;
    }
};
template <class K> struct BinByPage { // level 2: the page inside the sub-table
    uint32_t sub_mask, page_log2;
    __device__ __forceinline__ uint32_t operator()(K key) const { return (KeyTraits<K>::slot_hash(key) & sub_mask) >> page_log2; }
};

// Multi-GPU fused scatter: bins are owner ranks and the bucket of owner o lives in the
// HBM of rank o (mapped peer memory, NVLink), in the region that rank reserves for this
// sender.  rxb[o] is rank o's receive base biased so that the sender's virtual position
// v = o * cap + fill is also the index into it.
constexpr int MAX_P2P_WORLD = 8;
struct PeerOut {
    void *rxb[MAX_P2P_WORLD];
    uint32_t world, bins_per_owner; // world == 0: single destination (ScatterOut::out)
    // Every (tile, owner) run is reserved in multiples of `pad` keys (128 bytes) and the gap is
    // filled with all-ones keys, which the owner skips: runs then start on 128-byte lines and
    // NVLink carries full packets (unaligned 2 KB runs reached 530 of the 712 GB/s that an
    // aligned store kernel gets).  pad == 1: off (all-ones is a real key: no canonicalisation at
    // full key width).
    uint32_t pad;
};

struct ScatterOut {
    unsigned long long *cursors;   // per bin
    unsigned long long bucket_cap; // 0 = exact mode
    void *out;                     // binned keys
    void *spill_out;               // overflow keys (one-pass mode)
    unsigned long long *spill_cursor;
    unsigned long long spill_cap;
};

// a key stored once and read by a later kernel: evict-first, so that it does not push the tables' and the
// buckets' live lines out of L2
__device__ __forceinline__ void store_stream(uint64_t *p, uint64_t v) { __stcs((unsigned long long *)p, (unsigned long long)v); }
__device__ __forceinline__ void store_stream(u128 *p, u128 v) {
    __stcs((ulonglong2 *)p, make_ulonglong2((unsigned long long)v, (unsigned long long)(v >> 64)));
}

// Copy-out of a sorted tile into PEER memory (NVLink): every run is written from its own start, so that a
// warp's store covers whole 128-byte lines of the destination (the runs start on line boundaries there,
// PeerOut::pad) -- lanes that straddle two runs split lines between two stores, and NVLink carries partial
// packets (7.2 against 5.9 ms for the sender of the key exchange on C3 at N = 2).  Few bins (owners): all
// threads walk one run after the other; many bins (owner x sub-table): a warp per run.
template <class K, int THREADS>
__device__ __forceinline__ void copy_out_runs(const K *keys, const uint32_t *start, const unsigned long long *gaddr,
                                              uint32_t n_bins, uint32_t total) {
    constexpr uint32_t NW = THREADS / 32;
    if (n_bins <= 2 * NW) {
        for (uint32_t b = 0; b < n_bins; ++b) {
            const uint32_t beg = start[b], end = b + 1 < n_bins ? start[b + 1] : total;
            const unsigned long long ga = gaddr[b];
            for (uint32_t i = beg + threadIdx.x; i < end; i += THREADS) *(K *)(ga + (unsigned long long)i * sizeof(K)) = keys[i];
        }
    }
    else {
        const uint32_t lane = threadIdx.x & 31;
        for (uint32_t b = threadIdx.x >> 5; b < n_bins; b += NW) {
            const uint32_t beg = start[b], end = b + 1 < n_bins ? start[b + 1] : total;
            if (beg == end) continue;
            const unsigned long long ga = gaddr[b];
            for (uint32_t i = beg + lane; i < end; i += 32) *(K *)(ga + (unsigned long long)i * sizeof(K)) = keys[i];
        }
    }
}

// Preconditions: sm.cnt[parity] is zero, the block is synchronised.
// Leaves sm.cnt[parity ^ 1] zeroed and (SYNC_END) the block synchronised.
// Output positions are 64-bit (a stage may hold any number of keys): every bin of the tile has the
// bin's 64-bit `gaddr` (the byte address its run would have at tile position 0) which the copy-out adds
// to the position; it finds the bin by `bin_of(key)`; neighbouring lanes mostly share a bin, so the
// gaddr load is a broadcast.
// ALL: every key of the tile is valid (vmask is ignored).  Bins must be < n_bins for invalid keys too.
// SYNC_END = false: no barrier after the copy-out.  The next call's first writes to what the copy-out reads
// (the sorted tile, gaddr, start, s_total, s_ovf) all come after its own first barrier, so a caller that
// does not touch the tile's shared memory between two calls may let its warps run ahead into the next
// tile's loads and hashing while the others still copy out (5 barriers per tile instead of 6).  Used by the
// level-2 kernel (1.35 -> 1.32 ms on C2); the level-1 kernel was a little slower without the barrier.
// STREAM: the copy-out's stores are evict-first (st.global.cs).  Measured per kernel: it helps the level-2
// scatter of the large tables (512 pages per sub-table, u128 keys: C3 8.71 -> 8.35 ms, C3 k = 63 10.7 -> 10.2)
// and hurts level 1 and the 256-thread level-2 geometry of C2 (1.14 -> 1.22 ms).
template <class K, int THREADS, int PER, bool ALL = false, bool SYNC_END = true, bool STREAM = false, class BinFn>
__device__ __forceinline__ void tile_scatter(const K (&key)[PER], const uint32_t (&bin)[PER],
                                             uint32_t vmask, ScatterSmem<K, THREADS * PER> &sm,
                                             uint32_t n_bins, unsigned long long *cursors,
                                             uint64_t bin_off, const ScatterOut &o, uint32_t parity,
                                             const BinFn &bin_of, const PeerOut *po = nullptr) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total, s_ovf;
    static_assert(THREADS * PER <= 65536, "ranks are kept in 16 bits");
    uint32_t *cnt = sm.cnt + parity * n_bins;
    // bin (low 16 bits) and rank inside the bin's run (high 16 bits) of every key in ONE register.  As two
    // values the bin did not survive the 64-register cap: the compiler recomputed the key's hash and its
    // range reduction for the placement (14 of the level-1 kernel's 118 instructions per key, 11 of the
    // 85 of level 2; profiles/r02_sass_tile_scatter.md).  The rank comes out of the atomic, so the packed
    // word cannot be rematerialised.
    uint32_t br[PER];
    KTG_PHASE_BEGIN();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        // an invalid key adds 0 to its (in-range) bin: no branch around the atomic
        uint32_t w = bin[j] | (atomicAdd(&cnt[bin[j]], ALL ? 1u : ((vmask >> j) & 1u)) << 16);
        asm volatile("" : "+r"(w)); // opaque: keeps the compiler from splitting the word again and recomputing the bin later
        br[j] = w;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_ovf = 0; // (after the barrier: a warp of the previous call may still have been reading it)
    KTG_PHASE(1);
    // block-wide exclusive scan of the bin counts; reserve HBM ranges.  The global atomicAdd
    // that reserves a bin's range is issued BEFORE the scan so that its round trip to L2
    // (~1 us, the longest single latency of a tile) overlaps the scan's barriers.
    {
        const uint32_t per = (n_bins + THREADS - 1) / THREADS;
        const uint32_t b0 = threadIdx.x * per;
        const bool one = per == 1; // the usual case: at most one bin per thread
        // (a fast path for two bins per thread as well made the one-bin case 4 % slower and did not help at
        // 316 bins: what makes the level-1 scatter twice as slow there is the run length, not the scan)
        uint32_t s = 0, c1 = 0;
        unsigned long long base1 = 0;
        const uint32_t padm = po ? po->pad - 1u : 0u; // reservations are rounded up to pad keys
        if (one) {
            if (b0 < n_bins) c1 = cnt[b0];
            if (c1) base1 = atomicAdd(&cursors[b0], (unsigned long long)((c1 + padm) & ~padm));
            s = c1;
        }
        else {
            for (uint32_t i = b0; i < b0 + per && i < n_bins; ++i) s += cnt[i];
        }
        uint32_t incl = s;
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += x;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t v = threadIdx.x < THREADS / 32 ? s_warp[threadIdx.x] : 0, iv = v;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t x = __shfl_up_sync(0xFFFFFFFFu, iv, d);
                if (threadIdx.x >= d) iv += x;
            }
            s_warp[threadIdx.x] = iv - v;
            if (threadIdx.x == 31) s_total = iv;
        }
        __syncthreads();
        uint32_t run = s_warp[threadIdx.x >> 5] + incl - s;
        for (uint32_t i = b0; i < b0 + per && i < n_bins; ++i) {
            const uint32_t c = one ? c1 : cnt[i];
            unsigned long long base = base1;
            if (c) {
                if (!one) base = atomicAdd(&cursors[i], (unsigned long long)((c + padm) & ~padm));
                const unsigned long long dst0 = (unsigned long long)(po ? po->rxb[i / po->bins_per_owner] : o.out);
                sm.gaddr[i] = dst0 + (base - run) * sizeof(K);
                if (o.bucket_cap) {
                    const unsigned long long lim = (bin_off + i + 1) * o.bucket_cap;
                    if (base + c > lim) {
                        const unsigned long long over = base + c - (base > lim ? base : lim);
                        sm.spill[i] = atomicAdd(o.spill_cursor, over);
                        s_ovf = 1;
                    }
                }
            }
            sm.start[i] = run;
            run += c;
        }
    }
    __syncthreads();
    KTG_PHASE(2);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const uint32_t b = br[j] & 0xFFFFu;
        uint32_t pos = sm.start[b] + (br[j] >> 16);
        // an invalid key goes to the last position of the tile, which no valid key can have then and
        // which the copy-out does not reach: no branch around the stores
        if (!ALL && !((vmask >> j) & 1u)) pos = THREADS * PER - 1;
        sm.keys[pos] = key[j];
    }
    {   // the other counter buffer is free now: zero it for the next tile
        uint32_t *nxt = sm.cnt + (parity ^ 1u) * n_bins;
        for (uint32_t i = threadIdx.x; i < n_bins; i += THREADS) nxt[i] = 0;
    }
    __syncthreads();
    KTG_PHASE(3);
    const uint32_t total = s_total;
    if (!s_ovf && po) copy_out_runs<K, THREADS>(sm.keys, sm.start, sm.gaddr, n_bins, total);
    else if (!s_ovf) { // runs of the sorted tile to their buckets
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const uint32_t i = threadIdx.x + (uint32_t)(j * THREADS);
            if (ALL || i < total) {
                const K kk = sm.keys[i];
                K *dst = (K *)(sm.gaddr[bin_of(kk)] + (unsigned long long)i * sizeof(K));
                if (STREAM) store_stream(dst, kk);
                else *dst = kk;
            }
        }
    }
    else { // some bin of this tile ran past its bucket (rare)
        for (uint32_t i = threadIdx.x; i < total; i += THREADS) {
            const uint32_t b = bin_of(sm.keys[i]);
            K *out = (K *)(po ? po->rxb[b / po->bins_per_owner] : o.out);
            // position of the bin's run in `out` minus its start in the tile (may be negative)
            const long long gdelta = (long long)(sm.gaddr[b] - (unsigned long long)out) / (long long)sizeof(K);
            const unsigned long long lim = (bin_off + b + 1) * o.bucket_cap;
            const unsigned long long dst = (unsigned long long)(gdelta + i), base = (unsigned long long)(gdelta + sm.start[b]);
            if (dst < lim) out[dst] = sm.keys[i];
            else {
                const unsigned long long first = base > lim ? base : lim;
                const unsigned long long so = sm.spill[b] + (dst - first);
                if (so < o.spill_cap) ((K *)o.spill_out)[so] = sm.keys[i];
            }
        }
    }
    if (po && po->pad > 1) { // fill the rounded-up tail of every run (inside the bucket) with all-ones keys
        const uint32_t padm = po->pad - 1u;
        for (uint32_t b = 0; b < n_bins; ++b) {
            const uint32_t c = cnt[b];
            if (c == 0 || (c & padm) == 0) continue;
            K *dst = (K *)po->rxb[b / po->bins_per_owner];
            const long long gdelta = (long long)(sm.gaddr[b] - (unsigned long long)dst) / (long long)sizeof(K);
            const unsigned long long base = (unsigned long long)(gdelta + sm.start[b]), lim = (bin_off + b + 1) * o.bucket_cap;
            for (uint32_t j = c + threadIdx.x; j < ((c + padm) & ~padm); j += THREADS)
                if (base + j < lim) dst[base + j] = KeyTraits<K>::empty();
        }
    }
    if (SYNC_END) __syncthreads();
    KTG_PHASE(4);
}

constexpr int SCATTER_THREADS = 256, SCATTER_PER = 8, SCATTER_TILE = SCATTER_THREADS * SCATTER_PER;

// Level 1 from packed reads: extraction (K2) + partition.  BINS: by sub-table, or by owner
// rank -- into a local array for an NCCL exchange (po.world == 0) or straight into the
// owners' receive buckets over NVLink (po).
// BIN_OWNER_PART: (owner rank, sub-table of that owner), owner-major -- the sender of the direct exchange does
// the owner's level-1 partition as well, so that what arrives is already grouped by sub-table
constexpr int BIN_PART = 0, BIN_OWNER = 1, BIN_OWNER_PART = 2;
// PER: keys per lane and tile.  A work item is GRAN = 8 window starts; with PER = 4 it is partitioned in
// two passes of 4 keys (the rolling state survives the tile scatter in between).  That keeps the u128
// kernel's keys in registers (64 registers, no spill, against 80 and 28 bytes of spill) but halves the
// tile, and the per-tile costs (scan, barriers, reservations) decide: measured on C3 k=63 it takes
// 30.2 ms against 16.3, on C2 3.31 against 1.75 (profiles/r02_sweeps.md), so PER stays 8.
template <class K> struct ScatterGeom { static constexpr int PER = 8; };
template <class K, bool RC, int BINS, bool HLL, int PER = ScatterGeom<K>::PER>
__global__ void __launch_bounds__(SCATTER_THREADS, (BINS != BIN_PART || sizeof(K) == 8 || PER == 4) ? 4 : 3)
scatter_reads_kernel(ReadView v, uint32_t k, Table<K> t, uint32_t n_bins, ScatterOut o,
                     uint32_t *__restrict__ g_regs, PeerOut po) {
    extern __shared__ __align__(16) unsigned char smem[];
    static_assert(GRAN % PER == 0, "a work item is a whole number of passes");
    ScatterSmem<K, SCATTER_THREADS * PER> sm;
    sm.carve(smem, n_bins, HLL);
    if (HLL) {
        for (uint32_t i = threadIdx.x; i < HLL_M; i += SCATTER_THREADS) sm.regs[i] = 0;
    }
    for (uint32_t i = threadIdx.x; i < 2 * n_bins; i += SCATTER_THREADS) sm.cnt[i] = 0;
    __syncthreads();
    const uint64_t n_tiles = (v.n_items + SCATTER_THREADS - 1) / SCATTER_THREADS;
    uint32_t parity = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        ItemWindows<K> iw;
        iw.load(v, tile * SCATTER_THREADS + threadIdx.x, k);
        ItemWindows<K>::prefetch(v, (tile + gridDim.x) * SCATTER_THREADS + threadIdx.x); // this CTA's next tile
#pragma unroll
        for (int pass = 0; pass < GRAN / PER; ++pass, parity ^= 1u) {
            K key[PER];
            uint32_t bin[PER];
            uint32_t sampled = 0;
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                key[j] = iw.template key<RC>();
                const uint32_t h = KeyTraits<K>::place_hash(key[j]);
                Place p = place_of(h, t.world, t.n_sub);
                bin[j] = BINS == BIN_OWNER ? p.owner : BINS == BIN_PART ? p.part : p.owner * t.n_sub + p.part;
                if (HLL && hll_sampled(h)) sampled |= 1u << j;
                iw.r.step();
            }
            const uint32_t vmask = (iw.mask >> (pass * PER)) & ((1u << PER) - 1u);
            if (HLL) hll_update_tile<K, PER>(sm.regs, key, sampled & vmask);
            tile_scatter<K, SCATTER_THREADS, PER>(key, bin, vmask, sm, n_bins, o.cursors, 0, o, parity,
                                                  BinByPlace<K, BINS>{t.world, t.n_sub}, (BINS != BIN_PART && po.world) ? &po : nullptr);
        }
    }
    if (HLL) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < HLL_M; i += SCATTER_THREADS)
            if (sm.regs[i]) atomicMax(&g_regs[i], sm.regs[i]);
    }
}

// ---- Level 1 with a BIG tile, for many bins (more than one per thread of the kernel above).
// The one-pass kernel keeps a tile's keys in registers across its scan, which caps the tile at 2048
// keys; with 450 bins (C3 at k = 63) that is 4.5 keys = 73 bytes per run and 450 cursor reservations
// (64-bit L2 atomics) per 2048 keys, and the kernel takes 12 ps per key where 104 bins cost 5.4.
// This one extracts every window TWICE -- once to count the bins, once more, after the scan, to put the
// key where it belongs in the sorted tile -- so that a thread holds one key at a time and the tile is as
// large as shared memory: 64 KB of keys (8192 u64 / 4096 u128) under 512 threads, two CTAs per SM.
// Extraction is the cheap part (30-40 ALU instructions per window against the shared-memory sort), the
// per-bin work is spread over 2-4 times as many keys and the runs are that much longer.
constexpr int BIG_THREADS = 512;
template <class K> struct BigGeom {
    static constexpr int IPT = sizeof(K) == 8 ? 2 : 1;      // work items (of GRAN windows) per thread and tile
    static constexpr int TILE = BIG_THREADS * IPT * GRAN;   // keys
};
__host__ __device__ constexpr size_t big_scatter_smem(size_t tile, size_t key_bytes, size_t n_bins, bool hll) {
    // per bin: gaddr 8, spill 8, start 4, count / cursor 4; 128-bit keys: the bin of every tile position, 2 bytes
    return tile * key_bytes + n_bins * 24 + (hll ? (size_t)4096 * 4 : 0) + (key_bytes == 16 ? tile * 2 : 0);
}

template <class K, bool RC, int BINS, bool HLL>
__global__ void __launch_bounds__(BIG_THREADS, 2)
scatter_reads_big_kernel(ReadView v, uint32_t k, Table<K> t, uint32_t n_bins, ScatterOut o,
                         uint32_t *__restrict__ g_regs, PeerOut po_) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int IPT = BigGeom<K>::IPT, TILE = BigGeom<K>::TILE, THREADS = BIG_THREADS;
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_total, s_ovf;
    K *keys = (K *)smem;
    unsigned long long *gaddr = (unsigned long long *)(keys + TILE);
    unsigned long long *spill = gaddr + n_bins;
    uint32_t *start = (uint32_t *)(spill + n_bins);
    uint32_t *cnt = start + n_bins; // the bins' counts (pass A), then their running cursors in the sorted tile (pass B)
    uint32_t *regs = cnt + n_bins;
    // The bins found in pass A stay in registers (16 bits each) for pass B, and for 128-bit keys pass B also
    // leaves the bin of every tile position in shared memory for the copy-out: this kernel is bound by the
    // math pipes (its extraction runs twice), and the hash of a u128 key is 14 instructions.  For u64 keys
    // the copy-out recomputes the bin like the one-pass kernel (no room for the array beside 8192 keys).
    constexpr bool BINOF = sizeof(K) == 16;
    uint16_t *binof = (uint16_t *)(regs + (HLL ? HLL_M : 0));
    uint32_t bins2[IPT * GRAN / 2];
    const PeerOut *po = (BINS != BIN_PART && po_.world) ? &po_ : nullptr;
    const BinByPlace<K, BINS> bin_of{t.world, t.n_sub};
    for (uint32_t i = threadIdx.x; i < n_bins + (HLL ? HLL_M : 0); i += THREADS) cnt[i] = 0; // counts and sketch are adjacent
    __syncthreads();
    const uint64_t n_tiles = (v.n_items + THREADS * IPT - 1) / (THREADS * IPT);
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // pass A: count
#pragma unroll
        for (int a = 0; a < IPT; ++a) {
            ItemWindows<K> iw;
            iw.load(v, (tile * IPT + a) * THREADS + threadIdx.x, k);
#pragma unroll
            for (int j = 0; j < GRAN; ++j) {
                const K key = iw.template key<RC>();
                const uint32_t h = KeyTraits<K>::place_hash(key);
                const Place p = place_of(h, t.world, t.n_sub);
                const uint32_t b = BINS == BIN_OWNER ? p.owner : BINS == BIN_PART ? p.part : p.owner * t.n_sub + p.part;
                const uint32_t valid = (iw.mask >> j) & 1u;
                atomicAdd(&cnt[b], valid); // an invalid key adds 0 to its (in-range) bin: no branch
                if (HLL) hll_update(regs, key, h, valid != 0);
                if (j & 1) bins2[(a * GRAN + j) / 2] |= b << 16;
                else bins2[(a * GRAN + j) / 2] = b;
                iw.r.step();
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_ovf = 0;
        {   // scan of the counts, reservation of the output ranges (as in tile_scatter), counts -> cursors
            // (n_bins <= 2 * THREADS: at most two bins per thread, both reservations go out before the scan so
            // that their round trips to L2 overlap its barriers)
            const uint32_t b0 = threadIdx.x * 2;
            const uint32_t padm = po ? po->pad - 1u : 0u;
            uint32_t c2[2] = {0, 0};
            unsigned long long base2[2] = {0, 0};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (b0 + q < n_bins) c2[q] = cnt[b0 + q];
                if (c2[q]) base2[q] = atomicAdd(&o.cursors[b0 + q], (unsigned long long)((c2[q] + padm) & ~padm));
            }
            const uint32_t s = c2[0] + c2[1];
            uint32_t incl = s;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((threadIdx.x & 31) >= d) incl += x;
            }
            if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
            __syncthreads();
            if (threadIdx.x < 32) {
                uint32_t w = threadIdx.x < THREADS / 32 ? s_warp[threadIdx.x] : 0, iv = w;
                for (int d = 1; d < 32; d <<= 1) {
                    uint32_t x = __shfl_up_sync(0xFFFFFFFFu, iv, d);
                    if (threadIdx.x >= d) iv += x;
                }
                s_warp[threadIdx.x] = iv - w;
                if (threadIdx.x == 31) s_total = iv;
            }
            __syncthreads();
            uint32_t run = s_warp[threadIdx.x >> 5] + incl - s;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const uint32_t i = b0 + q, c = c2[q];
                if (i >= n_bins) break;
                if (c) {
                    const unsigned long long base = base2[q];
                    const unsigned long long dst0 = (unsigned long long)(po ? po->rxb[i / po->bins_per_owner] : o.out);
                    gaddr[i] = dst0 + (base - run) * sizeof(K);
                    if (o.bucket_cap) {
                        const unsigned long long lim = ((unsigned long long)i + 1) * o.bucket_cap;
                        if (base + c > lim) {
                            const unsigned long long over = base + c - (base > lim ? base : lim);
                            spill[i] = atomicAdd(o.spill_cursor, over);
                            s_ovf = 1;
                        }
                    }
                }
                start[i] = run;
                cnt[i] = run;
                run += c;
            }
        }
        __syncthreads();
        // pass B: the same windows again, every key to its place in the sorted tile
#pragma unroll
        for (int a = 0; a < IPT; ++a) {
            ItemWindows<K> iw;
            iw.load(v, (tile * IPT + a) * THREADS + threadIdx.x, k);
#pragma unroll
            for (int j = 0; j < GRAN; ++j) {
                const K key = iw.template key<RC>();
                const uint32_t b = (bins2[(a * GRAN + j) / 2] >> (16 * (j & 1))) & 0xFFFFu;
                const uint32_t valid = (iw.mask >> j) & 1u;
                uint32_t pos = atomicAdd(&cnt[b], valid);
                if (!valid) pos = TILE - 1; // the last position is free whenever a key is invalid, and never copied out then
                keys[pos] = key;
                if (BINOF) binof[pos] = (uint16_t)b;
                iw.r.step();
            }
        }
        __syncthreads();
        const uint32_t total = s_total;
        // (Leaving by bulk copies instead -- cp.async.bulk shared -> global, one per run, issued by the thread
        // that owns the bin, the next tile's counting pass running meanwhile -- was implemented and measured:
        // 14.5 against 13.5 ms on C3 k = 63, 7.97 against 8.04 ms for the sender of the direct exchange at
        // N = 2; a few hundred bytes per copy is too little for the copy engine.)
        if (!s_ovf && po) copy_out_runs<K, THREADS>(keys, start, gaddr, n_bins, total);
        else if (!s_ovf) {
            for (uint32_t i = threadIdx.x; i < total; i += THREADS) {
                const K kk = keys[i];
                *(K *)(gaddr[BINOF ? (uint32_t)binof[i] : bin_of(kk)] + (unsigned long long)i * sizeof(K)) = kk;
            }
        }
        else { // some bin of this tile ran past its bucket (rare)
            for (uint32_t i = threadIdx.x; i < total; i += THREADS) {
                const uint32_t b = BINOF ? (uint32_t)binof[i] : bin_of(keys[i]);
                K *out = (K *)(po ? po->rxb[b / po->bins_per_owner] : o.out);
                const long long gdelta = (long long)(gaddr[b] - (unsigned long long)out) / (long long)sizeof(K);
                const unsigned long long lim = ((unsigned long long)b + 1) * o.bucket_cap;
                const unsigned long long dst = (unsigned long long)(gdelta + i), base = (unsigned long long)(gdelta + start[b]);
                if (dst < lim) out[dst] = keys[i];
                else {
                    const unsigned long long first = base > lim ? base : lim;
                    const unsigned long long so = spill[b] + (dst - first);
                    if (so < o.spill_cap) ((K *)o.spill_out)[so] = keys[i];
                }
            }
        }
        if (po && po->pad > 1) { // the rounded-up tail of every run (inside the bucket) is filled with all-ones keys
            const uint32_t padm = po->pad - 1u;
            for (uint32_t b = 0; b < n_bins; ++b) {
                const uint32_t c = (b + 1 < n_bins ? start[b + 1] : total) - start[b];
                if (c == 0 || (c & padm) == 0) continue;
                K *dst = (K *)po->rxb[b / po->bins_per_owner];
                const long long gdelta = (long long)(gaddr[b] - (unsigned long long)dst) / (long long)sizeof(K);
                const unsigned long long base = (unsigned long long)(gdelta + start[b]), lim = ((unsigned long long)b + 1) * o.bucket_cap;
                for (uint32_t j = c + threadIdx.x; j < ((c + padm) & ~padm); j += THREADS)
                    if (base + j < lim) dst[base + j] = KeyTraits<K>::empty();
            }
        }
        __syncthreads(); // the copy-out has read the tile, the cursors are dead: counts back to zero for the next tile
        for (uint32_t i = threadIdx.x; i < n_bins; i += THREADS) cnt[i] = 0;
        __syncthreads();
    }
    if (HLL) {
        for (uint32_t i = threadIdx.x; i < HLL_M; i += THREADS)
            if (regs[i]) atomicMax(&g_regs[i], regs[i]);
    }
}

// Level 1 for an array of keys (receiver side of the multi-GPU exchange)
template <class K, bool BY_OWNER, bool HLL>
__global__ void __launch_bounds__(SCATTER_THREADS)
scatter_keys_kernel(const K *__restrict__ keys, uint64_t n, Table<K> t, uint32_t n_bins,
                    ScatterOut o, uint32_t *__restrict__ g_regs) {
    extern __shared__ __align__(16) unsigned char smem[];
    ScatterSmem<K, SCATTER_TILE> sm;
    sm.carve(smem, n_bins, HLL);
    if (HLL) {
        for (uint32_t i = threadIdx.x; i < HLL_M; i += SCATTER_THREADS) sm.regs[i] = 0;
    }
    for (uint32_t i = threadIdx.x; i < 2 * n_bins; i += SCATTER_THREADS) sm.cnt[i] = 0;
    __syncthreads();
    const uint64_t n_tiles = (n + SCATTER_TILE - 1) / SCATTER_TILE;
    uint32_t parity = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, parity ^= 1u) {
        const uint64_t base = tile * SCATTER_TILE;
        K key[SCATTER_PER];
        uint32_t bin[SCATTER_PER];
        uint32_t sampled = 0, vmask = 0;
#pragma unroll
        for (int j = 0; j < SCATTER_PER; ++j) {
            const uint64_t i = base + (uint64_t)j * SCATTER_THREADS + threadIdx.x;
            const bool in = i < n;
            key[j] = in ? KeyTraits<K>::load_stream(&keys[i]) : (K)0;
            const uint32_t h = KeyTraits<K>::place_hash(key[j]);
            Place p = place_of(h, t.world, t.n_sub);
            bin[j] = BY_OWNER ? p.owner : p.part;
            if (HLL && hll_sampled(h) && in) sampled |= 1u << j;
            if (in) vmask |= 1u << j;
        }
        if (HLL) hll_update_tile<K, SCATTER_PER>(sm.regs, key, sampled);
        tile_scatter<K, SCATTER_THREADS, SCATTER_PER>(key, bin, vmask, sm, n_bins, o.cursors, 0, o, parity,
                                                      BinByPlace<K, BY_OWNER ? BIN_OWNER : BIN_PART>{t.world, t.n_sub});
    }
    if (HLL) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < HLL_M; i += SCATTER_THREADS)
            if (sm.regs[i]) atomicMax(&g_regs[i], sm.regs[i]);
    }
}

// ---- level 2: the keys of sub-table b (level-1 bucket b) grouped by page.
// Tile t covers level-1 bin t / tiles_per_bin; page p of sub-table b owns
// keys2[(b*n2+p)*cap2, +cap2) and cursors2[b*n2+p] starts at (b*n2+p)*cap2.
constexpr int L2S_THREADS = 512, L2S_PER = 8, L2S_TILE = L2S_THREADS * L2S_PER;

// offsets of a batch of equally long reads: out[i] = first + i * len (the host batcher does not
// copy 8 bytes per read over PCIe when it has seen that they are all the same)
__global__ void fill_offsets_kernel(uint64_t *__restrict__ out, uint64_t n, uint64_t first, uint64_t len) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = first + i * len;
}

__global__ void init_cursors_kernel(unsigned long long *cursors, uint64_t n, uint64_t cap) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) cursors[i] = i * cap;
}

// LEVEL 2: bucket q holds keys of sub-table q (or q % sub_mod); they are grouped by page.
// LEVEL 1: bucket q holds keys of this shard in no particular order (what source rank q sent
//          in the fused multi-GPU exchange); they are grouped by sub-table.
// HLL (level 1 only): the keys also feed the cardinality sketch g_regs (the super-k-mer exchange
// sketches on the owner's side; sampled keys go straight to the global registers).
template <class K, int LEVEL, int L2S_THREADS = ktg::L2S_THREADS, int L2S_PER = ktg::L2S_PER, int MINB = 2, bool HLL = false>
__global__ void __launch_bounds__(L2S_THREADS, MINB)
scatter_buckets_kernel(const K *__restrict__ keys1, const unsigned long long *__restrict__ fill1,
                       uint64_t cap1, uint64_t tiles_per_bin, uint64_t n_tiles, uint32_t sub_mod,
                       bool skip_empty, Table<K> t, ScatterOut o, uint32_t *__restrict__ g_regs = nullptr) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int L2S_TILE = L2S_THREADS * L2S_PER;
    constexpr bool L2_STREAM = LEVEL == 2 && !(L2S_THREADS == 256 && L2S_PER == 8); // the geometries of the large tables
    const uint32_t n2 = LEVEL == 2 ? t.pages_per_sub() : t.n_sub;
    ScatterSmem<K, L2S_TILE> sm;
    sm.carve(smem, n2, false);
    for (uint32_t i = threadIdx.x; i < 2 * n2; i += L2S_THREADS) sm.cnt[i] = 0;
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t tpb = (uint32_t)tiles_per_bin;
    // bucket q = tile / tpb, tile inside it tq = tile % tpb, sub-table b = q % sub_mod, all advanced without
    // dividing: a division per tile and thread is ~25 instructions that every lane repeats (as tile / tpb
    // and q % sub_mod they were 10 % of this kernel's instructions)
    uint32_t q = blockIdx.x / tpb, tq = blockIdx.x - q * tpb;
    uint32_t b = LEVEL == 2 ? (sub_mod ? q % sub_mod : q) : 0;
    const uint32_t dq = gridDim.x / tpb, dt = gridDim.x - dq * tpb;
    auto advance = [&]() {
        uint32_t adv = dq;
        tq += dt;
        if (tq >= tpb) {
            tq -= tpb;
            ++adv;
        }
        q += adv;
        if (LEVEL == 2) {
            b += adv;
            if (sub_mod) while (b >= sub_mod) b -= sub_mod;
        }
    };
    for (uint32_t tile = blockIdx.x; tile < (uint32_t)n_tiles; tile += gridDim.x, advance()) {
        const uint64_t lim = ((uint64_t)q + 1) * cap1, fill = fill1[q];
        const uint64_t base = (uint64_t)q * cap1 + (uint64_t)tq * L2S_TILE;
        // This CTA's next tile: its keys would come straight from HBM into the registers that hash them
        // (long scoreboard was this kernel's first stall reason), so every thread asks L2 for its share of
        // them once this tile's own loads are out (C2: 1.27 -> 1.18 ms).  Only up to the bucket's fill: a
        // stage sized for more keys than it holds has whole tiles of unused capacity.
        uint64_t nbase = 0, nfill = 0;
        if (tile + gridDim.x < (uint32_t)n_tiles) {
            uint32_t qn = q + dq, tn = tq + dt;
            if (tn >= tpb) {
                tn -= tpb;
                ++qn;
            }
            nbase = (uint64_t)qn * cap1 + (uint64_t)tn * L2S_TILE;
            const uint64_t nlim = ((uint64_t)qn + 1) * cap1, f = fill1[qn];
            nfill = f < nlim ? f : nlim;
        }
        auto prefetch_next = [&]() {
            const uint64_t first = nbase + (uint64_t)threadIdx.x * L2S_PER; // the thread's L2S_PER consecutive keys
            if (first < nfill) {
                const unsigned char *nx = (const unsigned char *)(keys1 + first);
#pragma unroll
                for (int h = 0; h < (int)(L2S_PER * sizeof(K) / 32); ++h) prefetch_l2(nx + h * 32);
            }
        };
        const uint64_t end = fill < lim ? fill : lim;
        if (base >= end) continue;
        K key[L2S_PER];
        uint32_t bin[L2S_PER];
        KTG_PHASE_BEGIN();
        if (LEVEL == 2 && !skip_empty && base + L2S_TILE <= end) {
            // a whole tile (all but the last of a bucket): no per-key bounds, every key is valid.
            // (Level 1, the receive side of the multi-GPU exchange, keeps the general path: with the
            // filler test inside, this variant made its 512-thread kernel slower, 1.65 -> 1.76 ms.)
#pragma unroll
            for (int j = 0; j < L2S_PER; ++j) {
                key[j] = KeyTraits<K>::load_stream(&keys1[base + (uint32_t)(j * L2S_THREADS) + threadIdx.x]);
                bin[j] = (KeyTraits<K>::slot_hash(key[j]) & t.sub_mask) >> t.page_log2;
            }
            KTG_PHASE(0);
            prefetch_next();
            tile_scatter<K, L2S_THREADS, L2S_PER, true, false, L2_STREAM>(key, bin, (1u << L2S_PER) - 1u, sm, n2, o.cursors + (uint64_t)b * n2,
                                                               (uint64_t)b * n2, o, parity, BinByPage<K>{t.sub_mask, t.page_log2});
            parity ^= 1u;
            continue;
        }
        uint32_t vmask = 0, sampled = 0;
        const uint32_t rem = (uint32_t)(end - base < L2S_TILE ? end - base : L2S_TILE);
#pragma unroll
        for (int j = 0; j < L2S_PER; ++j) {
            const uint32_t o32 = (uint32_t)(j * L2S_THREADS) + threadIdx.x;
            const bool in = o32 < rem;
            key[j] = in ? KeyTraits<K>::load_stream(&keys1[base + o32]) : (K)0;
            if (LEVEL == 2) bin[j] = (KeyTraits<K>::slot_hash(key[j]) & t.sub_mask) >> t.page_log2;
            else {
                const uint32_t h = KeyTraits<K>::place_hash(key[j]);
                bin[j] = place_of(h, t.world, t.n_sub).part;
                if (HLL && hll_sampled(h)) sampled |= 1u << j;
            }
            if (in && !(skip_empty && key[j] == KeyTraits<K>::empty())) vmask |= 1u << j; // padding of the exchange
        }
        if (HLL) hll_update_tile<K, L2S_PER>(g_regs, key, sampled & vmask);
#ifdef KTG_PHASE_TIMERS
        if (threadIdx.x == 0 && key[L2S_PER - 1] == (K)12345) g_phase_cycles[7] = 1; // wait for the loads
#endif
        KTG_PHASE(0);
        prefetch_next();
        if (LEVEL == 2)
            tile_scatter<K, L2S_THREADS, L2S_PER, false, false, L2_STREAM>(key, bin, vmask, sm, n2, o.cursors + (uint64_t)b * n2, (uint64_t)b * n2, o,
                                                                parity, BinByPage<K>{t.sub_mask, t.page_log2});
        else
            tile_scatter<K, L2S_THREADS, L2S_PER, false, false>(key, bin, vmask, sm, n2, o.cursors, 0, o, parity,
                                                                BinByPlace<K, BIN_PART>{t.world, t.n_sub});
        parity ^= 1u;
    }
}

// ======================================================================= K3
// Streaming page update: the table is swept once, page by page.  A page is [P keys | P weights] in
// HBM and the same bytes in shared memory, so it arrives as ONE bulk copy (cp.async.bulk, the 1-D TMA
// path: an elected thread issues it, an mbarrier counts the bytes, SASS UBLKCP / SYNCS) and leaves as
// one bulk store; no thread spends instructions on moving the table.  The CTA inserts the page's
// keys with shared-memory CAS / atomicAdd (add_single_edge + create_or_modify_edge,
// hm_gir.rs:91-153, hs_gir.rs:192-203).  No L2 atomics: HBM sees one coalesced read of the keys
// and one read + write of the table.  A table that is logically empty (right after ktg_reset, its
// memory undefined) loads every page from a 1-page template of empty slots instead (L2 resident).
// NBUF == 2: one persistent CTA per SM holds two page buffers and pipelines them -- page g+1 is
// loading and page g-1 is being stored while page g is probed.  NBUF == 1: two CTAs per SM cover
// each other's copies.
constexpr int PAGE_UNROLL = 4;
template <class K> struct PageGeom;
template <> struct PageGeom<uint64_t> { static constexpr uint32_t LOG2 = 13; }; // 8192 x (8+4) B =  96 KB
template <> struct PageGeom<u128> { static constexpr uint32_t LOG2 = 12; };     // 4096 x (16+4) B = 80 KB

__device__ __forceinline__ uint64_t smem_cas(uint64_t *p, uint64_t cmp, uint64_t val) {
    return atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)val);
}
__device__ __forceinline__ u128 smem_cas(u128 *p, u128 cmp, u128 val) {
    uint64_t olo, ohi;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile(
        "{\n\t"
        ".reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.shared.relaxed.cta.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}\n"
        : "=l"(olo), "=l"(ohi)
        : "l"((uint64_t)cmp), "l"((uint64_t)(cmp >> 64)), "l"((uint64_t)val), "l"((uint64_t)(val >> 64)), "r"(addr)
        : "memory");
    return ((u128)ohi << 64) | olo;
}
__device__ __forceinline__ uint64_t smem_load(const uint64_t *p) { return *(const volatile uint64_t *)p; }
__device__ __forceinline__ u128 smem_load(const u128 *p) {
    uint64_t lo, hi;
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr) : "memory");
    return ((u128)hi << 64) | lo;
}

// One probe step for a whole warp.  Divergent per-lane probe loops cost the
// MAXIMUM chain length of the 32 lanes per row (ncu: 138 warp instructions per
// row of 32 keys, 58 % of them in the loop); instead every lane probes exactly
// one slot, and the keys that are not resolved yet go to a warp-private retry
// queue in shared memory that is drained in full rows of 32, so the work is the
// SUM of the chain lengths.  Queue entry: key + (slot | probes << 13.. | inc2 << 31).
constexpr uint32_t PQ_CAP = 64; // < 32 left over + <= 32 pushed per step
// queue state word: slot in bits 0..15, probes done in bits 16..30, "+2" flag in bit 31
constexpr uint32_t PQ_INC2 = 0x80000000u, PQ_STEP = 0x00010001u, PQ_PROBE_SHIFT = 16;

template <class K> struct PageCtx {
    K *sk;
    uint32_t *sw;
    K *qk;        // this warp's queue
    uint32_t *qi;
    uint32_t qn;  // warp-uniform
    uint32_t page_mask;
};

// One probe of one slot: resolves the key (claim / add) or advances st to the next slot and
// returns true ("still pending").  `cur` is the slot content loaded by the caller, so that
// several independent loads can be in flight before the first one is used.
template <class K>
__device__ __forceinline__ bool page_probe_once(PageCtx<K> &c, const Table<K> &t, K key, uint32_t &st, K cur) {
    typedef KeyTraits<K> T;
    const K EMPTY = T::empty();
    const uint32_t i = st & 0xFFFFu;
    if (cur != key && (cur == EMPTY || T::maybe_torn(cur))) cur = smem_cas(&c.sk[i], EMPTY, key);
    if (cur == key || cur == EMPTY) {
        atomicAdd(&c.sw[i], 1u + (st >> 31));
        return false;
    }
    st = (st + PQ_STEP) & (0xFFFF0000u | c.page_mask); // next slot (wraps inside the page), one more probe
    if (st & ((c.page_mask + 1u) << PQ_PROBE_SHIFT)) { // 2^page_log2 probes: every slot holds another key
        unsigned long long pos = atomicAdd(t.ovf_count, 1ull); // replayed after a grow
        if (pos < t.ovf_cap) {
            t.ovf_keys[pos] = key;
            t.ovf_inc[pos] = 1u + (st >> 31);
        }
        return false;
    }
    return true;
}

// all 32 lanes: append the pending lanes' (key, st) to the warp queue
template <class K>
__device__ __forceinline__ void page_push(PageCtx<K> &c, K key, uint32_t st, bool pending) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, pending);
    if (m == 0) return;
    if (pending) {
        const uint32_t pos = c.qn + __popc(m & ((1u << (threadIdx.x & 31)) - 1u));
        c.qk[pos] = key;
        c.qi[pos] = st;
    }
    c.qn += __popc(m);
}

// all 32 lanes: retry full rows of 32 queued keys until fewer than 32 are left
template <class K> __device__ __forceinline__ void page_retry_rows(PageCtx<K> &c, const Table<K> &t) {
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
    while (c.qn >= 32) {
        c.qn -= 32;
        const K key = c.qk[c.qn + lane];
        uint32_t st = c.qi[c.qn + lane];
        __syncwarp();
        const bool pending = page_probe_once(c, t, key, st, smem_load(&c.sk[st & 0xFFFFu]));
        page_push(c, key, st, pending);
        __syncwarp();
    }
}

// the partial row left in the queue at the end of a page
template <class K> __device__ __forceinline__ void page_drain(PageCtx<K> &c, const Table<K> &t) {
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
    while (c.qn) {
        const uint32_t cnt = c.qn; // < 32
        const bool active = lane < cnt;
        const K key = active ? c.qk[lane] : KeyTraits<K>::empty();
        uint32_t st = active ? c.qi[lane] : 0;
        c.qn = 0;
        __syncwarp();
        const bool pending = active && page_probe_once(c, t, key, st, smem_load(&c.sk[st & 0xFFFFu]));
        page_push(c, key, st, pending);
        __syncwarp();
    }
}

// shared memory of update_pages_kernel: NBUF page images | the warps' retry queues | NBUF mbarriers
template <class K> __host__ __device__ constexpr size_t page_kernel_smem(int threads, int nbuf, uint32_t page_log2) {
    return (size_t)nbuf * ((sizeof(K) + 4) << page_log2) + (size_t)(threads / 32) * PQ_CAP * (sizeof(K) + 4) + 16;
}

// MODE: what a key may need besides the plain update (mutually exclusive, fixed per build):
//   PAGE_PALIN    reverse_complement with even k: a k-mer equal to its reverse complement counts twice;
//   PAGE_SPECIAL  no canonicalisation at full key width: the all-ones key (T...T) lives in the special weight.
// As run-time flags the two tests were 4 % of the kernel's instructions for every build.
constexpr int PAGE_PLAIN = 0, PAGE_PALIN = 1, PAGE_SPECIAL = 2;
template <class K, int THREADS, int NBUF, int MODE = PAGE_PLAIN, int PAGE_UNROLL = ktg::PAGE_UNROLL>
__global__ void __launch_bounds__(THREADS, NBUF == 1 ? 2 : 1)
update_pages_kernel(const K *__restrict__ keys2, const unsigned long long *__restrict__ cursors2,
                    uint64_t cap2, uint32_t k, Table<K> t, const unsigned char *__restrict__ empty_page) {
    constexpr bool check_palindrome = MODE == PAGE_PALIN, has_special = MODE == PAGE_SPECIAL;
    typedef KeyTraits<K> T;
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t P = 1u << t.page_log2;
    const uint32_t pbytes = (uint32_t)t.page_bytes();
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    PageCtx<K> c;
    K *qk_all = (K *)(smem + (size_t)NBUF * pbytes);
    uint32_t *qi_all = (uint32_t *)(qk_all + (THREADS / 32) * PQ_CAP);
    uint64_t *bars = (uint64_t *)(qi_all + (THREADS / 32) * PQ_CAP);
    c.qk = qk_all + wid * PQ_CAP;
    c.qi = qi_all + wid * PQ_CAP;
    c.qn = 0;
    c.page_mask = t.page_mask;
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&bars[b], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint64_t n_pages = t.n_pages();
    auto issue_load = [&](uint64_t g, uint32_t buf) { // one thread
        mbar_expect_tx(&bars[buf], pbytes);
        bulk_g2s(smem + (size_t)buf * pbytes, empty_page ? empty_page : t.base + g * pbytes, pbytes, &bars[buf]);
    };
    if (threadIdx.x == 0 && blockIdx.x < n_pages) issue_load(blockIdx.x, 0);
    uint32_t it = 0;
    for (uint64_t g = blockIdx.x; g < n_pages; g += gridDim.x, ++it) {
        const uint32_t buf = NBUF == 1 ? 0u : (it & 1u);
        if (NBUF == 2 && threadIdx.x == 0 && g + gridDim.x < n_pages) {
            bulk_wait_read_all(); // the store of page g - gridDim.x has finished reading the other buffer
            issue_load(g + gridDim.x, buf ^ 1u);
        }
        c.sk = (K *)(smem + (size_t)buf * pbytes);
        c.sw = (uint32_t *)(smem + (size_t)buf * pbytes + ((size_t)sizeof(K) << t.page_log2));
        const uint64_t beg = g * cap2, lim = beg + cap2, cur = cursors2[g];
        const uint32_t n = (uint32_t)((cur < lim ? cur : lim) - beg);
        const K *src = keys2 + beg;
        mbar_wait(&bars[buf], (it / NBUF) & 1u);
        // a warp takes PAGE_UNROLL consecutive rows of 32 keys at a time; FULL: all of them inside
        // the bucket (all but its last rows), so no per-key bounds
        auto rows = [&](uint32_t r0, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            K my[PAGE_UNROLL];
#pragma unroll
            for (int q = 0; q < PAGE_UNROLL; ++q) {
                const uint32_t i = r0 + q * 32 + lane;
                my[q] = (FULL || i < n) ? T::load_stream(&src[i]) : T::empty();
            }
            // first probes of the PAGE_UNROLL keys: independent shared-memory loads in flight together
            uint32_t st[PAGE_UNROLL];
            K cur[PAGE_UNROLL];
            bool active[PAGE_UNROLL];
#pragma unroll
            for (int q = 0; q < PAGE_UNROLL; ++q) {
                active[q] = FULL || r0 + q * 32 + lane < n;
                if (has_special && active[q] && my[q] == T::empty()) { // all-T at full key width, no canonicalisation
                    atomicAdd(t.special_w(), 1u);
                    active[q] = false;
                }
                st[q] = T::slot_hash(my[q]) & c.page_mask;
                if (check_palindrome && revcomp(my[q], k) == my[q]) st[q] |= PQ_INC2;
                cur[q] = smem_load(&c.sk[st[q] & 0xFFFFu]);
            }
            // The first probe only takes the hit -- the slot already holds the key, 30 of 32 lanes on C2 --
            // with one compare and one atomicAdd.  Everything else (an empty slot to claim, another key
            // to step over) goes to the warp's queue with the slot NOT advanced and is handled there in
            // full rows of 32 by page_probe_once: the claim / collision code, which the whole warp used
            // to walk through for the 2-6 lanes that needed it, now runs with every lane busy
            // (57 -> 37 warp instructions per row of 32 keys on C2).
#pragma unroll
            for (int q = 0; q < PAGE_UNROLL; ++q) {
                const bool hit = active[q] && cur[q] == my[q];
                if (hit) atomicAdd(&c.sw[st[q] & 0xFFFFu], 1u + (st[q] >> 31));
                page_push(c, my[q], st[q], active[q] && !hit);
                page_retry_rows(c, t);
            }
        };
        // The keys are read once, straight from HBM (~800 cycles) into the registers that hash them, and a
        // warp has nothing else to do meanwhile (long scoreboard was the first stall reason).  Every lane
        // asks L2 for one 32-byte sector of the rows its warp takes NEXT (32 lanes cover 1 KB: the
        // PAGE_UNROLL rows of u64 keys, half of them for u128), and the warp's first rows of the CTA's
        // next page are requested when this page starts.
        constexpr uint32_t PF_GROUPS = PAGE_UNROLL * 32 * sizeof(K) / 1024; // 1 KB requests per group of rows
        auto prefetch_rows = [&](const K *p, uint32_t r, uint32_t cnt) {
#pragma unroll
            for (uint32_t h = 0; h < PF_GROUPS; ++h) {
                const uint32_t off = h * 1024 + lane * 32; // bytes from the first key of the group
                if (r * (uint32_t)sizeof(K) + off < cnt * (uint32_t)sizeof(K)) prefetch_l2((const unsigned char *)(p + r) + off);
            }
        };
        if (g + gridDim.x < n_pages) {
            const uint64_t gn = g + gridDim.x, begn = gn * cap2, curn = cursors2[gn];
            const uint32_t nn = (uint32_t)((curn < begn + cap2 ? curn : begn + cap2) - begn);
            prefetch_rows(keys2 + begn, wid * 32 * PAGE_UNROLL, nn);
        }
        for (uint32_t r0 = wid * 32 * PAGE_UNROLL; r0 < n; r0 += THREADS * PAGE_UNROLL) {
            prefetch_rows(src, r0 + THREADS * PAGE_UNROLL, n);
            if (r0 + 32 * PAGE_UNROLL <= n) rows(r0, std::true_type{});
            else rows(r0, std::false_type{});
        }
        page_drain(c, t);
        fence_proxy_async(); // this thread's shared-memory writes, before the bulk store reads them
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_s2g(t.base + g * pbytes, smem + (size_t)buf * pbytes, pbytes);
            if (NBUF == 1 && g + gridDim.x < n_pages) {
                bulk_wait_read_all();
                issue_load(g + gridDim.x, 0); // the other threads wait on the mbarrier
            }
        }
    }
    if (threadIdx.x == 0) bulk_wait_all();
    (void)P;
}

// K3 with L2 atomics over an array of canonical keys (small batches against a
// large table, spill lists, keys received from peers).  Keys arrive grouped by
// sub-table.  Tiles are handed out through a global counter (persistent CTAs,
// dynamic scheduling) instead of a static grid-stride: with a static split the
// SMs drift apart over the ~1000 iterations (near/far L2 latency differs per
// SM), the in-flight keys end up spread over many sub-tables and the working
// set falls out of L2 (ncu: 27 GB of DRAM reads for 322 M inserts).  With the
// counter every CTA works at the global frontier, i.e. on the one or two
// sub-tables that are L2 resident.
constexpr int INSERT_TILE_PER_THREAD = 8;
constexpr uint64_t INSERT_TILE = 256 * INSERT_TILE_PER_THREAD;
// bin_end == nullptr: keys[0, n) is dense (n_dev != nullptr: n = min(*n_dev, n)).
// Otherwise bin b holds keys[b*bucket_cap, min(bin_end[b], (b+1)*bucket_cap))
// (one-pass partitioner) and tile t covers bin t / tiles_per_bin.
template <class K>
__global__ void __launch_bounds__(256)
insert_keys_kernel(const K *__restrict__ keys, uint64_t n, const unsigned long long *__restrict__ n_dev,
                   unsigned long long *lost, uint32_t k, bool check_palindrome, bool skip_empty, Table<K> t, unsigned long long *tile_counter,
                   const unsigned long long *__restrict__ bin_end, uint64_t bucket_cap,
                   uint64_t tiles_per_bin, uint64_t n_tiles) {
    __shared__ unsigned long long s_tile;
    if (n_dev) {
        const uint64_t m = *n_dev;
        if (m > n && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(lost, (unsigned long long)(m - n));
        n = m < n ? m : n;
        n_tiles = (n + INSERT_TILE - 1) / INSERT_TILE;
    }
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1ull);
        __syncthreads();
        const uint64_t tile = s_tile;
        __syncthreads();
        if (tile >= n_tiles) break;
        uint64_t base, end;
        if (bin_end) {
            const uint64_t b = tile / tiles_per_bin;
            const uint64_t lim = (b + 1) * bucket_cap, fill = bin_end[b];
            base = b * bucket_cap + (tile % tiles_per_bin) * INSERT_TILE;
            end = fill < lim ? fill : lim;
        }
        else {
            base = tile * INSERT_TILE;
            end = n;
        }
        if (base >= end) continue;
        K my[INSERT_TILE_PER_THREAD];
#pragma unroll
        for (int q = 0; q < INSERT_TILE_PER_THREAD; ++q) {
            uint64_t i = base + q * 256 + threadIdx.x;
            my[q] = i < end ? KeyTraits<K>::load_stream(&keys[i]) : KeyTraits<K>::empty(); // read once
        }
#pragma unroll
        for (int q = 0; q < INSERT_TILE_PER_THREAD; ++q) {
            uint64_t i = base + q * 256 + threadIdx.x;
            if (i < end && !(skip_empty && my[q] == KeyTraits<K>::empty())) {
                uint32_t inc = (check_palindrome && revcomp(my[q], k) == my[q]) ? 2u : 1u;
                table_add(t, my[q], inc);
            }
        }
    }
}

// replay of inserts that overflowed a full sub-table, after the table grew
// ---- BFCounter input: n pre-counted k-mers of exactly k ASCII bases each.
// add_read_bfc (/root/reference/src/katome/collections/graphs/pt_graph.rs:318-329): the k-mer and,
// with reverse_complement, its reverse complement get an edge of the line's weight; lines below
// the threshold are skipped (algorithms/builder.rs:106-108).  On the canonical table that is one
// update of min(kmer, revcomp) by the weight (twice the weight for a palindrome, which the
// reference inserts twice).  counters[0] += accepted lines, counters[1] += lines with a byte
// outside ACGT (the build is void then).
template <class K, bool RC>
__global__ void __launch_bounds__(256)
insert_weighted_kmers_kernel(const uint8_t *__restrict__ kmers, const uint32_t *__restrict__ weights, uint64_t n,
                             uint32_t k, uint32_t threshold, Table<K> t, unsigned long long *__restrict__ counters) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long ok = 0, bad = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t w = weights[i];
        if (w < threshold) continue;
        if (w == 0) { // weight 0 means "not an edge" in this table (a removed edge keeps its slot that way): a
            ++ok;     // count-0 line is accepted and counted like the reference does, but adds no edge
            continue;
        }
        const uint8_t *p = kmers + i * k;
        K key = 0;
        bool valid = true;
        for (uint32_t j = 0; j < k; ++j) {
            const uint32_t c = p[j];
            valid &= c == 'A' || c == 'C' || c == 'G' || c == 'T';
            key = (key << 2) | (K)(((c >> 1) & 3u) ^ ((c >> 2) & 1u)); // A0 C1 G2 T3 (compress.rs:347-378)
        }
        if (!valid) {
            ++bad;
            continue;
        }
        uint32_t inc = w;
        if (RC) {
            const K r = revcomp(key, k);
            if (r == key) inc = 2u * w;
            else if (r < key) key = r;
        }
        table_add(t, key, inc);
        ++ok;
    }
    ok = warp_sum(ok);
    bad = warp_sum(bad);
    if ((threadIdx.x & 31) == 0) {
        if (ok) atomicAdd(&counters[0], ok);
        if (bad) atomicAdd(&counters[1], bad);
    }
}

template <class K>
__global__ void replay_overflow_kernel(const K *__restrict__ keys, const uint32_t *__restrict__ inc,
                                       uint64_t n, Table<K> t) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        table_add(t, keys[i], inc[i]);
}

// ================================================================ table scans
// All scans walk the slots in index order: slot i lives in page i >> page_log2 (Table::key_ptr /
// w_ptr), so a warp reads 32 consecutive weights (128 bytes) and, where it needs them, 32
// consecutive keys.  Index capacity() is the special entry (all-ones key, weight only).
template <class K> __device__ __forceinline__ uint32_t *slot_weight(const Table<K> &t, uint64_t i) {
    return i < t.capacity() ? t.w_ptr(i) : t.special_w();
}
template <class K> __device__ __forceinline__ K slot_key(const Table<K> &t, uint64_t i) {
    return i < t.capacity() ? KeyTraits<K>::load(t.key_ptr(i)) : KeyTraits<K>::empty();
}

// every 16 bytes of an empty table: keys all-ones, weights zero
template <class K> __global__ void init_table_kernel(Table<K> t) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t per_page = t.page_bytes() / 16, key_chunks = ((uint64_t)sizeof(K) << t.page_log2) / 16;
    const uint64_t total = t.n_pages() * per_page;
    const uint4 ones = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), zero = make_uint4(0u, 0u, 0u, 0u);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= total; i += stride) {
        const uint64_t c = i % per_page;
        ((uint4 *)t.base)[i] = (i < total && c < key_chunks) ? ones : zero; // i == total: the special entry
    }
}

template <class K> __global__ void count_occupied_kernel(Table<K> t, unsigned long long *out) {
    typedef KeyTraits<K> T;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = t.capacity();
    uint64_t acc[1] = {0};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc[0] += T::load(t.key_ptr(i)) != T::empty();
    block_accumulate<1>(acc, out);
}

template <class K> __global__ void rehash_kernel(Table<K> old, Table<K> t) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = old.capacity();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        const uint32_t w = *slot_weight(old, i);
        if (w == 0) continue; // empty, or an edge removed by the filter
        table_add(t, slot_key(old, i), w);
    }
}

// Stats over the both-strand expanded edge set: a canonical entry c with
// weight w stands for edges c and revcomp(c) (one edge if c is a palindrome);
// see SURVEY Appendix A.5-7.
struct EdgeStats {
    unsigned long long edges, sum_w, sum_w_below, max_w, digest;
};

__device__ __forceinline__ uint64_t digest_term(uint64_t hi, uint64_t lo, uint32_t w) {
    return splitmix64(splitmix64(hi) ^ lo) * (2ull * w + 1ull);
}

// Less than half of the slots are occupied and an occupied one costs ~130 instructions (four
// splitmix64 rounds and a reverse complement), so lanes do not digest "their" slots: a warp
// compacts the occupied slots it streams past into a small shared-memory queue and digests them in
// full rows of 32 (the scan was 45 % divergence-idle and instruction bound at 3.2 TB/s before).
template <class K, bool RC>
__global__ void __launch_bounds__(256)
edge_stats_kernel(Table<K> t, uint32_t k, uint32_t threshold, EdgeStats *out) {
    typedef KeyTraits<K> T;
    constexpr int U = 4, QCAP = 64; // < 32 left over + <= 32 pushed per step
    __shared__ K q_key[8][QCAP];
    __shared__ uint32_t q_w[8][QCAP];
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t acc[4] = {0, 0, 0, 0}; // edges, sum_w, sum_w_below, digest
    uint64_t mx = 0;
    uint32_t qn = 0; // warp-uniform
    auto digest_one = [&](K key, uint32_t w) {
        uint64_t mult = 1;
        uint64_t d = digest_term(T::hi(key), T::lo(key), w);
        if (RC) {
            K r = revcomp(key, k);
            if (r != key) {
                mult = 2;
                d += digest_term(T::hi(r), T::lo(r), w);
            }
        }
        acc[0] += mult;
        acc[1] += mult * w;
        if (w < threshold) acc[2] += mult * w;
        acc[3] += d;
        if (w > mx) mx = w;
    };
    // weights first (4 bytes per slot, U rows in flight per lane), keys of the occupied slots only
    const uint64_t n_slots = t.capacity() + 1;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t gw = (uint64_t)blockIdx.x * (blockDim.x >> 5) + wid;
    for (uint64_t base = gw * (32 * U); base < n_slots; base += warps * (32 * U)) {
        uint32_t w[U];
        K key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = base + u * 32 + lane;
            w[u] = i < n_slots ? __ldcs(slot_weight(t, i)) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = base + u * 32 + lane;
            key[u] = w[u] ? slot_key(t, i) : (K)0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, w[u] != 0);
            if (m == 0) continue;
            if (w[u] != 0) {
                const uint32_t pos = qn + __popc(m & ((1u << lane) - 1u));
                q_key[wid][pos] = key[u];
                q_w[wid][pos] = w[u];
            }
            qn += __popc(m);
            __syncwarp();
            if (qn >= 32) {
                qn -= 32;
                digest_one(q_key[wid][qn + lane], q_w[wid][qn + lane]);
                __syncwarp();
            }
        }
    }
    if (lane < qn) digest_one(q_key[wid][lane], q_w[wid][lane]);
    uint64_t a4[4] = {acc[0], acc[1], acc[2], acc[3]};
    // EdgeStats layout: edges, sum_w, sum_w_below, max_w, digest
    __shared__ unsigned long long sh[5];
    if (threadIdx.x < 5) sh[threadIdx.x] = 0;
    __syncthreads();
    for (int i = 0; i < 4; ++i) {
        uint64_t s = warp_sum(a4[i]);
        if ((threadIdx.x & 31) == 0 && s) atomicAdd(&sh[i < 3 ? i : 4], (unsigned long long)s);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) atomicMax(&sh[3], (unsigned long long)mx);
    __syncthreads();
    if (threadIdx.x < 5 && sh[threadIdx.x]) {
        if (threadIdx.x == 3) atomicMax(&out->max_w, sh[3]);
        else atomicAdd(((unsigned long long *)out) + threadIdx.x, sh[threadIdx.x]);
    }
}

// ======================================================================= K4
// Clean::remove_weak_edges (pruner.rs:109-118, edges.rs:51-58): keep w >= t.
// In place: a removed edge keeps its slot (probe chains stay intact) with
// weight 0 == "not in E".  Reads (and rarely writes) the weights only: 4 bytes per slot.
template <class K> __global__ void filter_kernel(Table<K> t, uint32_t threshold) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = t.capacity();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        uint32_t *wp = slot_weight(t, i);
        const uint32_t w = *wp;
        if (w != 0 && w < threshold) *wp = 0;
    }
}

// standardize_edges (standardizer.rs:56-69): w' = round_half_away(w * p) as
// u32 (saturating), w' = 1 if it rounded to 0 but w >= t; w' == 0 removes the
// edge (remove_weak_edges(1)).  p is computed on the host from the device sums.
__device__ __forceinline__ uint32_t scale_weight(uint32_t w, double p, uint32_t threshold) {
    double r = round((double)w * p); // round(): half away from zero, like f64::round
    uint32_t nw;
    if (!(r >= 0.0)) nw = 0;
    else if (r >= 4294967295.0) nw = 0xFFFFFFFFu;
    else nw = (uint32_t)r;
    if (nw == 0 && w >= threshold) nw = 1;
    return nw;
}

template <class K> __global__ void standardize_kernel(Table<K> t, double p, uint32_t threshold) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = t.capacity();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        uint32_t *wp = slot_weight(t, i);
        const uint32_t w = *wp;
        if (w != 0) *wp = scale_weight(w, p, threshold);
    }
}

// Stream compaction of the surviving edges (w >= threshold, threshold >= 1)
// into dense arrays, expanding both strands.  Block-aggregated: one atomicAdd
// on the output cursor per CTA iteration.  This is the export that
// Convert::create_from (hm_gir.rs:156-226) consumes and, with threshold > 1,
// the fused "filter + compact" of north_star kernel (4).  The output arrays may live on
// another GPU (peer memory): a sharded export compacts every shard straight into the
// gathering device's arrays over NVLink.
template <class K, bool RC>
__global__ void __launch_bounds__(256)
compact_edges_kernel(Table<K> t, uint32_t k, uint32_t threshold, uint64_t *__restrict__ out_hi,
                     uint64_t *__restrict__ out_lo, uint32_t *__restrict__ out_w, uint64_t cap,
                     unsigned long long *cursor) {
    typedef KeyTraits<K> T;
    __shared__ uint32_t s_warp[8];
    __shared__ unsigned long long s_base;
    const uint64_t n_slots = t.capacity() + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_round = (n_slots + blockDim.x - 1) / blockDim.x * blockDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        uint32_t w = i < n_slots ? *slot_weight(t, i) : 0;
        uint32_t cnt = 0;
        K key = 0, r = 0;
        if (w != 0 && w >= threshold) {
            key = slot_key(t, i);
            cnt = 1;
            if (RC) {
                r = revcomp(key, k);
                if (r != key) cnt = 2;
            }
        }
        // exclusive scan of cnt over the block
        uint32_t incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += x;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            for (int q = 0; q < 8; ++q) {
                uint32_t v = s_warp[q];
                s_warp[q] = run;
                run += v;
            }
            s_base = run ? atomicAdd(cursor, (unsigned long long)run) : 0ull;
        }
        __syncthreads();
        uint64_t pos = s_base + s_warp[wid] + incl - cnt;
        if (cnt >= 1 && pos < cap) {
            if (out_hi) out_hi[pos] = T::hi(key);
            out_lo[pos] = T::lo(key);
            out_w[pos] = w;
        }
        if (cnt == 2 && pos + 1 < cap) {
            if (out_hi) out_hi[pos + 1] = T::hi(r);
            out_lo[pos + 1] = T::lo(r);
            out_w[pos + 1] = w;
        }
        __syncthreads();
    }
}

// ============================================================ node (k-1)-mers
// Node set N = {prefix, suffix of every edge} (hm_gir.rs:99-149; sinks are keys
// with empty Outgoing) and the degree data behind CollectionStats
// (stats/collections.rs:137-168).  One entry per canonical (k-1)-mer; its u32
// holds four 8-bit degree counters: out/in of the canonical orientation and
// out/in of the reverse-complement orientation (each <= 4).
template <class KE, class KN, bool RC>
__global__ void __launch_bounds__(256)
build_nodes_kernel(Table<KE> t, uint32_t k, Table<KN> nt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = t.capacity();
    const uint32_t k1 = k - 1;
    const KE nmask = (k1 == 8 * sizeof(KE) / 2) ? ~(KE)0 : (KE)((((KE)1) << (2 * k1)) - 1);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        if (*slot_weight(t, i) == 0) continue;
        KE e = slot_key(t, i);
        KE er = RC ? revcomp(e, k) : e;
        const int n_exp = (RC && er != e) ? 2 : 1;
        for (int x = 0; x < n_exp; ++x) {
            KE edge = x ? er : e;
            KN pre = (KN)(edge >> 2), suf = (KN)(edge & nmask);
#pragma unroll
            for (int side = 0; side < 2; ++side) { // 0: prefix gains an out-edge, 1: suffix an in-edge
                KN nd = side ? suf : pre;
                uint32_t field = side;
                if (RC) {
                    KN nr = revcomp(nd, k1);
                    if (nr < nd) {
                        nd = nr;
                        field += 2;
                    }
                }
                table_add(nt, nd, 1u << (8 * field));
            }
        }
    }
}

struct NodeStats {
    unsigned long long nodes, sources, sinks, max_in, max_out;
};

// (key, degree word) of every occupied node-table entry, warp-aggregated compaction: what a shard
// sends to the owners of its nodes when the table is sharded over GPUs
template <class KN>
__global__ void __launch_bounds__(256)
compact_nodes_kernel(Table<KN> nt, KN *__restrict__ out_keys, uint32_t *__restrict__ out_deg, unsigned long long *cursor) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n_slots = nt.capacity() + 1;
    const uint64_t n_round = (n_slots + 31) & ~(uint64_t)31;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const uint32_t w = i < n_slots ? *slot_weight(nt, i) : 0;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, w != 0);
        if (m == 0) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (w) {
            const unsigned long long pos = base + __popc(m & ((1u << lane) - 1u));
            out_keys[pos] = slot_key(nt, i);
            out_deg[pos] = w;
        }
    }
}

// merge (key, degree word) pairs received from all shards: the degree fields are disjoint
// counters (<= 4 each in total), so adding the words adds the degrees
template <class KN>
__global__ void __launch_bounds__(256)
merge_nodes_kernel(const KN *__restrict__ keys, const uint32_t *__restrict__ deg, uint64_t n, Table<KN> nt) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        table_add(nt, keys[i], deg[i]);
}

template <class KN, bool RC>
__global__ void __launch_bounds__(256)
node_stats_kernel(Table<KN> nt, NodeStats *out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, n = nt.capacity();
    uint64_t acc[3] = {0, 0, 0};
    uint64_t mi = 0, mo = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        uint32_t w = *slot_weight(nt, i);
        if (w == 0) continue;
#pragma unroll
        for (int o = 0; o < (RC ? 2 : 1); ++o) {
            uint32_t od = (w >> (16 * o)) & 0xFF, id = (w >> (16 * o + 8)) & 0xFF;
            if (od + id == 0) continue;
            acc[0] += 1;
            acc[1] += id == 0;
            acc[2] += od == 0;
            if (id > mi) mi = id;
            if (od > mo) mo = od;
        }
    }
    block_accumulate<3>(acc, (unsigned long long *)out);
    mi = warp_max(mi);
    mo = warp_max(mo);
    if ((threadIdx.x & 31) == 0) {
        if (mi) atomicMax(&out->max_in, (unsigned long long)mi);
        if (mo) atomicMax(&out->max_out, (unsigned long long)mo);
    }
}

// ===================================================================== synth
// Device twin of oracle ko_synth_reads (counter based).
__global__ void synth_reads_kernel(uint8_t *__restrict__ out, uint64_t seed_g, uint64_t G,
                                   uint32_t L, uint64_t thr, uint64_t r0, uint64_t n_reads) {
    const uint64_t seed_r = seed_g ^ 0x5245414453ull, seed_e = seed_g ^ 0x4552524f52ull;
    const uint32_t chunks = (L + 3) / 4;
    const uint64_t total = n_reads * chunks, stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint64_t r = r0 + i / chunks;
        const uint32_t j0 = (uint32_t)(i % chunks) * 4;
        const uint64_t start = splitmix64(seed_r + 2 * r) % (G - L + 1);
        const bool strand = splitmix64(seed_r + 2 * r + 1) & 1;
        uint8_t *dst = out + (r - r0) * L;
        for (uint32_t j = j0; j < j0 + 4 && j < L; ++j) {
            uint32_t code = strand ? 3u - (uint32_t)(splitmix64(seed_g + start + (L - 1 - j)) & 3)
                                   : (uint32_t)(splitmix64(seed_g + start + j) & 3);
            uint64_t h = splitmix64(seed_e + r * L + j);
            if (h < thr) code = (code + 1 + (uint32_t)(splitmix64(h) % 3)) & 3;
            dst[j] = (uint8_t)((0x54474341u >> (8 * code)) & 0xFF);
        }
    }
}

// uniformly random "load key + atomicAdd weight" over a table of n_slots
// (the random-access roofline the insert kernel is compared against)
template <int SLOT_BYTES>
__global__ void random_access_probe_kernel(unsigned char *table, uint64_t n_slots,
                                           uint64_t n_updates, unsigned long long *sink) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_updates; i += stride) {
        uint64_t h = splitmix64(i);
        uint64_t s = (uint64_t)(((u128)h * n_slots) >> 64);
        unsigned char *p = table + s * SLOT_BYTES;
        acc += __ldcg((const unsigned long long *)p);
        atomicAdd((unsigned int *)(p + (SLOT_BYTES == 16 ? 8 : 16)), 1u);
    }
    if (acc == 0x1234567ull) *sink = acc;
}

} // namespace ktg
