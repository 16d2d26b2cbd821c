// common.cuh -- key types, hashing and the open-addressing table primitives.
//
// Keys are k-mers as integers, first base in the most significant 2 bits
// (A0 C1 G2 T3; /root/reference/src/katome/compress.rs:347-378), right
// aligned: k <= 32 in a u64, 33..64 in a u128.  The table replaces katome's
// HashMap<NodeSlice, Outgoing> (collections/girs/hm_gir.rs:22): instead of a
// node-keyed map with <=4 outgoing edges per node we key directly on the edge
// k-mer and keep one u32 weight per edge (create_or_modify_edge,
// collections/girs/hs_gir.rs:192-203 -> atomicAdd).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ktg {

typedef unsigned __int128 u128;

// splitmix64 finalizer; also the building block of the digest and of the
// synthetic read generator (oracle/katome_oracle.c has the CPU twins).
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    return fmix64(x + 0x9E3779B97F4A7C15ull);
}
// 64-bit mixer of the multi-GPU control paths and tests (two multiply / fold rounds)
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x *= 0x9E3779B97F4A7C15ull;
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ull;
    x ^= x >> 32;
    return x;
}

// ---- reverse complement on packed integers (compress.rs:153-169 restated as
// bit tricks: complement = NOT (:164), reversal of 2-bit symbols) ----
__host__ __device__ __forceinline__ uint64_t rev2_u64(uint64_t x) {
#ifdef __CUDA_ARCH__
    x = __brevll(x);
#else
    x = ((x >> 32) | (x << 32));
    x = ((x >> 16) & 0x0000FFFF0000FFFFull) | ((x & 0x0000FFFF0000FFFFull) << 16);
    x = ((x >> 8) & 0x00FF00FF00FF00FFull) | ((x & 0x00FF00FF00FF00FFull) << 8);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((x & 0x0F0F0F0F0F0F0F0Full) << 4);
    x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
    x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
#endif
    // all 64 bits are reversed; swap the two bits of every symbol back
    return ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
}

__host__ __device__ __forceinline__ uint64_t revcomp(uint64_t x, uint32_t k) {
    return rev2_u64(~x) >> (64 - 2 * k);
}
__host__ __device__ __forceinline__ u128 revcomp(u128 x, uint32_t k) {
    uint64_t lo = (uint64_t)x, hi = (uint64_t)(x >> 64);
    u128 r = ((u128)rev2_u64(~lo) << 64) | rev2_u64(~hi);
    return r >> (128 - 2 * k);
}

// Two independent 32-bit hashes of a key place it (they only decide WHERE a key lives, never a
// result; uniformity on sequential, shifted, low-complexity and tandem-repeat key sets is checked
// in tests/test_hashing.py against the numpy mirror katome_b200/hashing.py):
//   place_hash  owner rank and sub-table (level-1 bin), through place_of;
//   slot_hash   home slot inside the sub-table; its high bits are the PAGE (level-2 bin).
// Each is a 32-bit finalizer over a multiplicative fold of the key words: 9 instructions, against
// ~20 for a 64-bit mixer (the level-1 scatter spent 15 % of its instructions on two 64-bit
// multiplies per window in round 1).  The folds are different linear forms of the words, so that
// keys that collide under one do not collide under the other.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) { // murmur3 finalizer
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t mix32b(uint32_t x) { // "lowbias32" constants
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

template <class K> struct KeyTraits;

template <> struct KeyTraits<uint64_t> {
    static constexpr int WORDS = 1;
    __host__ __device__ static __forceinline__ uint64_t empty() { return ~0ull; }
    __host__ __device__ static __forceinline__ uint32_t place_hash(uint64_t k) {
        return mix32b((uint32_t)(k >> 32) + (uint32_t)k * 0x85EBCA77u);
    }
    __host__ __device__ static __forceinline__ uint32_t slot_hash(uint64_t k) {
        return mix32((uint32_t)k + (uint32_t)(k >> 32) * 0x9E3779B1u);
    }
    // 64 well mixed bits (cardinality sketch, sampled keys only)
    __host__ __device__ static __forceinline__ uint64_t hash64(uint64_t k) { return fmix64(k ^ 0x9E3779B97F4A7C15ull); }
    __host__ __device__ static __forceinline__ uint64_t hi(uint64_t) { return 0; }
    __host__ __device__ static __forceinline__ uint64_t lo(uint64_t k) { return k; }
    __host__ __device__ static __forceinline__ uint64_t make(uint64_t, uint64_t lo_) { return lo_; }
#ifdef __CUDACC__
    __device__ static __forceinline__ uint64_t load(const uint64_t *p) { return __ldcg((const unsigned long long *)p); }
    __device__ static __forceinline__ uint64_t load_stream(const uint64_t *p) { // read once
        return __ldcs((const unsigned long long *)p);
    }
    __device__ static __forceinline__ bool maybe_torn(uint64_t) { return false; }
    __device__ static __forceinline__ uint64_t cas(uint64_t *p, uint64_t cmp, uint64_t val) {
        return atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)val);
    }
#endif
};

template <> struct KeyTraits<u128> {
    static constexpr int WORDS = 2;
    __host__ __device__ static __forceinline__ u128 empty() { return ~(u128)0; }
    __host__ __device__ static __forceinline__ uint32_t place_hash(u128 k) {
        return mix32b((uint32_t)(k >> 96) + (uint32_t)(k >> 64) * 0x85EBCA77u + (uint32_t)(k >> 32) * 0xC2B2AE3Du +
                      (uint32_t)k * 0x27D4EB2Fu);
    }
    __host__ __device__ static __forceinline__ uint32_t slot_hash(u128 k) {
        return mix32((uint32_t)k + (uint32_t)(k >> 32) * 0x9E3779B1u + (uint32_t)(k >> 64) * 0x85EBCA77u +
                     (uint32_t)(k >> 96) * 0xC2B2AE3Du);
    }
    __host__ __device__ static __forceinline__ uint64_t hash64(u128 k) {
        return fmix64((uint64_t)k ^ ((uint64_t)(k >> 64) * 0xA24BAED4963EE407ull) ^ 0x9E3779B97F4A7C15ull);
    }
    __host__ __device__ static __forceinline__ uint64_t hi(u128 k) { return (uint64_t)(k >> 64); }
    __host__ __device__ static __forceinline__ uint64_t lo(u128 k) { return (uint64_t)k; }
    __host__ __device__ static __forceinline__ u128 make(uint64_t hi_, uint64_t lo_) {
        return ((u128)hi_ << 64) | lo_;
    }
#ifdef __CUDACC__
    __device__ static __forceinline__ u128 load(const u128 *p) {
        ulonglong2 v = __ldcg((const ulonglong2 *)p);
        return ((u128)v.y << 64) | v.x;
    }
    __device__ static __forceinline__ u128 load_stream(const u128 *p) {
        ulonglong2 v = __ldcs((const ulonglong2 *)p);
        return ((u128)v.y << 64) | v.x;
    }
    // A 16-byte vector load is one transaction in practice but is not promised
    // to be single-copy atomic against the 128-bit CAS; a value with an
    // all-ones half could be a half-written key, so it is re-read with the CAS.
    __device__ static __forceinline__ bool maybe_torn(u128 k) {
        return (uint64_t)k == ~0ull || (uint64_t)(k >> 64) == ~0ull;
    }
    __device__ static __forceinline__ u128 cas(u128 *p, u128 cmp, u128 val) {
        uint64_t olo, ohi;
        asm volatile(
            "{\n\t"
            ".reg .b128 c, v, o;\n\t"
            "mov.b128 c, {%2, %3};\n\t"
            "mov.b128 v, {%4, %5};\n\t"
            "atom.global.relaxed.gpu.cas.b128 o, [%6], c, v;\n\t"
            "mov.b128 {%0, %1}, o;\n\t"
            "}\n"
            : "=l"(olo), "=l"(ohi)
            : "l"((uint64_t)cmp), "l"((uint64_t)(cmp >> 64)), "l"((uint64_t)val),
              "l"((uint64_t)(val >> 64)), "l"(p)
            : "memory");
        return ((u128)ohi << 64) | olo;
    }
#endif
};

// Where a key lives: owner rank (hash sharding over GPUs), sub-table (the
// L2-resident partition, = level-1 bin of the partitioner) and home slot inside
// it (slot_hash & sub_mask); the high bits of the slot select the PAGE (the shared-memory sized unit of
// the streaming update, = level-2 bin) and linear probing wraps inside the page.
// Owner and sub-table are disjoint pieces of place_hash, range-reduced by multiply-shift so that
// world and n_sub need not be powers of two; its low 9 bits pick the keys the sketch samples.
struct Place {
    uint32_t owner, part;
};
__host__ __device__ __forceinline__ Place place_of(uint32_t h, uint32_t world, uint32_t n_sub) {
    uint64_t t = (uint64_t)h * world;
    Place p;
    p.owner = (uint32_t)(t >> 32);
    p.part = (uint32_t)(((uint64_t)(uint32_t)t * n_sub) >> 32);
    return p;
}

// The edge table in HBM: open addressing, linear probing inside a page.  Structure of arrays PER
// PAGE -- page g is [2^page_log2 keys | 2^page_log2 u32 weights], 12 bytes per slot for u64 keys and
// 20 for u128 -- which is exactly the image the page update keeps in shared memory, so a page
// moves between HBM and an SM as ONE bulk copy each way (cp.async.bulk), the weight-only passes
// (filter, standardize, node statistics) read 4 bytes per slot, and nothing is spent on padding
// (round 1 had 16 / 32-byte array-of-structs slots).  One "special" entry past the last page holds
// the weight of the all-ones key (T...T at full key width without canonicalisation).
template <class K> struct Table {
    unsigned char *base; // n_pages() pages, then the special weight (16 bytes)
    uint32_t n_sub, sub_log2, sub_mask;
    uint32_t page_log2, page_mask; // page_log2 <= sub_log2; probing wraps inside a page
    uint32_t world, rank;
    uint32_t max_probe;
    // inserts that could not be placed (sub-table full): replayed after a grow
    K *ovf_keys;
    uint32_t *ovf_inc;
    unsigned long long *ovf_count;
    uint64_t ovf_cap;
    static constexpr uint32_t SLOT_BYTES = sizeof(K) + 4;
    __host__ __device__ uint64_t capacity() const { return (uint64_t)n_sub << sub_log2; }
    __host__ __device__ uint32_t pages_per_sub() const { return 1u << (sub_log2 - page_log2); }
    __host__ __device__ uint64_t n_pages() const { return (uint64_t)n_sub << (sub_log2 - page_log2); }
    __host__ __device__ uint64_t page_bytes() const { return (uint64_t)SLOT_BYTES << page_log2; }
    __host__ __device__ uint64_t bytes() const { return n_pages() * page_bytes() + 16; }
    __host__ __device__ K *page_keys(uint64_t g) const { return (K *)(base + g * page_bytes()); }
    __host__ __device__ uint32_t *page_weights(uint64_t g) const {
        return (uint32_t *)(base + g * page_bytes() + ((uint64_t)sizeof(K) << page_log2));
    }
    // slot i of the whole table, i < capacity()
    __host__ __device__ K *key_ptr(uint64_t i) const { return page_keys(i >> page_log2) + (i & page_mask); }
    __host__ __device__ uint32_t *w_ptr(uint64_t i) const { return page_weights(i >> page_log2) + (i & page_mask); }
    __host__ __device__ uint32_t *special_w() const { return (uint32_t *)(base + n_pages() * page_bytes()); }
};

#ifdef __CUDACC__
// weight[key] += inc.  The net effect of add_single_edge + create_or_modify_edge
// (hm_gir.rs:91-153, hs_gir.rs:192-203); u32 wraps like the release build.
template <class K>
__device__ __forceinline__ void table_add(const Table<K> &t, K key, uint32_t inc) {
    typedef KeyTraits<K> T;
    const K EMPTY = T::empty();
    if (key == EMPTY) { // all-T at full key width (only without canonicalisation)
        atomicAdd(t.special_w(), inc);
        return;
    }
    const Place p = place_of(T::place_hash(key), t.world, t.n_sub);
    const uint32_t slot = T::slot_hash(key) & t.sub_mask;
    const uint64_t g = (((uint64_t)p.part << t.sub_log2) + slot) >> t.page_log2;
    K *pk = t.page_keys(g);
    uint32_t *pw = t.page_weights(g);
    uint32_t i = slot & t.page_mask;
    for (uint32_t n = 0; n < t.max_probe; ++n) {
        K cur = T::load(pk + i);
        if (cur != key && (cur == EMPTY || T::maybe_torn(cur))) cur = T::cas(pk + i, EMPTY, key);
        if (cur == key || cur == EMPTY) {
            atomicAdd(pw + i, inc);
            return;
        }
        i = (i + 1) & t.page_mask;
    }
    unsigned long long pos = atomicAdd(t.ovf_count, 1ull);
    if (pos < t.ovf_cap) {
        t.ovf_keys[pos] = key;
        t.ovf_inc[pos] = inc;
    }
}
#endif

} // namespace ktg
