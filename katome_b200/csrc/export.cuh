// export.cuh -- the hand-off to Convert::create_from (collections/girs/hm_gir.rs:156-226,
// hs_gir.rs:205-262): the graph in a canonical numbering, built on the device.
//
//   nodes  = distinct (k-1)-mers, sorted ascending (the reference numbers them in HashMap
//            iteration order, hm_gir.rs:160-171, which is not part of the parity contract)
//   edges  = sorted by k-mer; edge e goes from node src[e] (its prefix) to node dst[e] (its suffix)
//   bytes  = every edge in compress_edge format (compress.rs:250-271): [padding count][bases,
//            first base in the most significant bits, zero padded], what SEQUENCES holds after
//            kmer_to_edge
//
// This is not the build hot path: sorting is cub::DeviceRadixSort (library), everything
// around it (splitting edges into nodes, head flags, ranks, byte strings) is ours.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace ktg {

// keys travel as (hi, lo) u64 arrays; hi == nullptr for keys of at most 64 bits
struct KeyArr {
    uint64_t *hi, *lo;
};

__global__ void iota_kernel(uint32_t *p, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (uint32_t)i;
}
template <class T> __global__ void gather_kernel(const T *__restrict__ src, const uint32_t *__restrict__ idx,
                                                 T *__restrict__ dst, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[idx[i]];
}

// prefix and suffix (k-1)-mers of every edge: cand[e] = prefix, cand[n + e] = suffix
__global__ void split_nodes_kernel(const uint64_t *__restrict__ e_hi, const uint64_t *__restrict__ e_lo, uint64_t n,
                                   uint32_t k, uint64_t *__restrict__ c_hi, uint64_t *__restrict__ c_lo) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t k1 = k - 1;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const u128 edge = ((u128)(e_hi ? e_hi[e] : 0) << 64) | e_lo[e];
        const u128 pre = edge >> 2;
        const u128 suf = k1 >= 64 ? edge : (edge & ((((u128)1) << (2 * k1)) - 1));
        c_lo[e] = (uint64_t)pre;
        c_lo[n + e] = (uint64_t)suf;
        if (c_hi) {
            c_hi[e] = (uint64_t)(pre >> 64);
            c_hi[n + e] = (uint64_t)(suf >> 64);
        }
    }
}

// flag[i] = 1 iff sorted key i differs from key i-1
__global__ void head_flags_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo, uint64_t n,
                                  uint32_t *__restrict__ flag) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        flag[i] = i == 0 || lo[i] != lo[i - 1] || (hi && hi[i] != hi[i - 1]);
}
__global__ void scatter_heads_kernel(const uint64_t *__restrict__ hi, const uint64_t *__restrict__ lo,
                                     const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos, uint64_t n,
                                     uint64_t *__restrict__ out_hi, uint64_t *__restrict__ out_lo) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (flag[i]) {
            out_lo[pos[i]] = lo[i];
            if (out_hi) out_hi[pos[i]] = hi[i];
        }
}

// index of every edge's prefix / suffix in the sorted node array (both are in it by construction)
__global__ void node_ids_kernel(const uint64_t *__restrict__ e_hi, const uint64_t *__restrict__ e_lo, uint64_t n_edges,
                                uint32_t k, const uint64_t *__restrict__ n_hi, const uint64_t *__restrict__ n_lo,
                                uint64_t n_nodes, uint64_t *__restrict__ src, uint64_t *__restrict__ dst) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t k1 = k - 1;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
        const u128 edge = ((u128)(e_hi ? e_hi[e] : 0) << 64) | e_lo[e];
        const u128 want[2] = {edge >> 2, k1 >= 64 ? edge : (edge & ((((u128)1) << (2 * k1)) - 1))};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            uint64_t lo = 0, hi = n_nodes; // first node >= want
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                const u128 v = ((u128)(n_hi ? n_hi[mid] : 0) << 64) | n_lo[mid];
                if (v < want[s]) lo = mid + 1;
                else hi = mid;
            }
            (s ? dst : src)[e] = lo;
        }
    }
}

// compress_edge (compress.rs:250-271): byte 0 = number of unused 2-bit places in the last byte,
// then ceil(k/4) bytes of bases, first base in the most significant bits
__global__ void edge_bytes_kernel(const uint64_t *__restrict__ e_hi, const uint64_t *__restrict__ e_lo, uint64_t n,
                                  uint32_t k, uint8_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t nb = (k + 3) / 4, pad = (4 - k % 4) % 4, rec = nb + 1;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const u128 v = (((u128)(e_hi ? e_hi[e] : 0) << 64) | e_lo[e]) << (2 * pad); // now 8 * nb bits
        uint8_t *o = out + e * rec;
        o[0] = (uint8_t)pad;
        for (uint32_t j = 0; j < nb; ++j) o[1 + j] = (uint8_t)(v >> (8 * (nb - 1 - j)));
    }
}

// ---- Externals (pruner.rs:165-195): nodes without an incoming edge (Input), else without an outgoing
// edge (Output), in node index order.  deg[v] bit 0: v has an outgoing edge, bit 1: an incoming one.
__global__ void mark_degrees_kernel(const uint64_t *__restrict__ src, const uint64_t *__restrict__ dst, uint64_t n_edges,
                                    uint32_t *__restrict__ deg) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += stride) {
        atomicOr(&deg[src[e]], 1u);
        atomicOr(&deg[dst[e]], 2u);
    }
}
__global__ void external_flags_kernel(const uint32_t *__restrict__ deg, uint64_t n_nodes, uint32_t *__restrict__ flag) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v <= n_nodes; v += stride)
        flag[v] = v < n_nodes && deg[v] != 3u;
}
__global__ void scatter_externals_kernel(const uint32_t *__restrict__ deg, const uint32_t *__restrict__ flag,
                                         const uint32_t *__restrict__ pos, uint64_t n_nodes, uint64_t *__restrict__ ids,
                                         uint8_t *__restrict__ kinds) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_nodes; v += stride)
        if (flag[v]) {
            ids[pos[v]] = v;
            kinds[pos[v]] = (deg[v] & 2u) ? 1 : 0; // no incoming edge: Input, whatever goes out (pruner.rs:181-186)
        }
}

// ---- host helpers -------------------------------------------------------------------------
struct Scratch { // device allocations of one export, freed together
    std::vector<void *> ptrs;
    ~Scratch() {
        for (void *p : ptrs) cudaFree(p);
    }
    template <class T> int alloc(T **p, size_t n) {
        void *q = nullptr;
        cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(KTG_ERR_CUDA, "cudaMalloc(%zu) for the export failed: %s", n * sizeof(T), cudaGetErrorString(e));
        }
        ptrs.push_back(q);
        *p = (T *)q;
        return KTG_OK;
    }
};

inline int export_grid(uint64_t n) { return (int)std::min<uint64_t>((n + 255) / 256 + 1, 148 * 8); }

// Stable LSD radix sort of n (hi, lo) keys: returns the permutation and the sorted keys.
// Two passes of cub::DeviceRadixSort::SortPairs (by lo, then by hi) over an index array.
inline int sort_keys(const KeyArr &in, uint64_t n, uint32_t key_bits, Scratch &sc, cudaStream_t st,
                     uint32_t **perm_out, KeyArr *sorted) {
    if (n >= 0xFFFFFFFFull) return fail(KTG_ERR_INVALID, "export of more than 2^32 entries is not supported");
    uint32_t *idx0, *idx1;
    uint64_t *k0, *k1;
    KTG_TRY(sc.alloc(&idx0, n));
    KTG_TRY(sc.alloc(&idx1, n));
    KTG_TRY(sc.alloc(&k0, n));
    KTG_TRY(sc.alloc(&k1, n));
    iota_kernel<<<export_grid(n), 256, 0, st>>>(idx0, n);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, in.lo, k0, idx0, idx1, (int)n, 0, 64, st);
    void *tmp;
    KTG_TRY(sc.alloc((uint8_t **)&tmp, tb));
    const int lo_bits = (int)std::min<uint32_t>(key_bits, 64);
    KTG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, in.lo, k0, idx0, idx1, (int)n, 0, lo_bits, st)); // idx1: by lo
    uint32_t *perm = idx1;
    if (in.hi && key_bits > 64) {
        gather_kernel<uint64_t><<<export_grid(n), 256, 0, st>>>(in.hi, idx1, k1, n); // hi in lo-sorted order
        KTG_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, k1, k0, idx1, idx0, (int)n, 0, (int)key_bits - 64, st));
        perm = idx0; // stable: ties in hi keep the lo order
    }
    uint64_t *s_lo, *s_hi = nullptr;
    KTG_TRY(sc.alloc(&s_lo, n));
    gather_kernel<uint64_t><<<export_grid(n), 256, 0, st>>>(in.lo, perm, s_lo, n);
    if (in.hi) {
        KTG_TRY(sc.alloc(&s_hi, n));
        gather_kernel<uint64_t><<<export_grid(n), 256, 0, st>>>(in.hi, perm, s_hi, n);
    }
    *perm_out = perm;
    sorted->hi = s_hi;
    sorted->lo = s_lo;
    KTG_CUDA(cudaGetLastError());
    return KTG_OK;
}

// distinct keys of a sorted array
inline int unique_sorted(const KeyArr &sorted, uint64_t n, Scratch &sc, cudaStream_t st, KeyArr *out, uint64_t *n_out) {
    uint32_t *flag, *pos;
    KTG_TRY(sc.alloc(&flag, n + 1));
    KTG_TRY(sc.alloc(&pos, n + 1));
    KTG_CUDA(cudaMemsetAsync(flag + n, 0, 4, st));
    head_flags_kernel<<<export_grid(n), 256, 0, st>>>(sorted.hi, sorted.lo, n, flag);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, (int)(n + 1), st);
    void *tmp;
    KTG_TRY(sc.alloc((uint8_t **)&tmp, tb));
    KTG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, flag, pos, (int)(n + 1), st));
    uint32_t total = 0;
    KTG_CUDA(cudaMemcpyAsync(&total, pos + n, 4, cudaMemcpyDeviceToHost, st));
    KTG_CUDA(cudaStreamSynchronize(st));
    KTG_TRY(sc.alloc(&out->lo, total));
    out->hi = nullptr;
    if (sorted.hi) KTG_TRY(sc.alloc(&out->hi, total));
    scatter_heads_kernel<<<export_grid(n), 256, 0, st>>>(sorted.hi, sorted.lo, flag, pos, n, out->hi, out->lo);
    KTG_CUDA(cudaGetLastError());
    *n_out = total;
    return KTG_OK;
}

} // namespace ktg
