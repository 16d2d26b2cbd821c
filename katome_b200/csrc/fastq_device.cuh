// fastq_device.cuh -- FASTQ record parsing on the device.
//
// Replaces the record iteration of create_fastq (algorithms/builder.rs:142-160), i.e. rust-bio
// 0.10's fastq::Reader (Cargo.lock:22-25; restated for the host in host_reader.h): strict 4-line
// records, the header must start with '@', seq() is line 2 with trailing whitespace trimmed, the
// '+' line and the quality line are read and dropped, end of file inside a record is an error.
//
// The host only moves raw file bytes (fread into pinned memory, one H2D copy); the device finds
// the newlines, cuts the records, checks the headers and gathers the sequence lines into the dense
// (bases, offsets) batch that ktg_add_reads_device takes.  A chunk always starts at a record
// boundary: the bytes after the last complete record are carried over to the next chunk.
#pragma once
#include <cub/cub.cuh>

#include "common.cuh"

namespace ktg {

constexpr int FQ_BLOCK_BYTES = 8192; // bytes scanned by one CTA of 256 threads (32 per thread)

__device__ __forceinline__ uint32_t newline_mask32(const uint8_t *p, uint64_t pos, uint64_t n) {
    // bit i: byte pos + i is '\n' (bytes past n do not count); p is 16-byte aligned, pos % 32 == 0
    uint32_t m = 0;
    if (pos + 32 <= n) {
        const uint4 a = *(const uint4 *)(p + pos), b = *(const uint4 *)(p + pos + 16);
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t x = w[q] ^ 0x0A0A0A0Au; // zero byte <=> newline
            const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); // 0x80 per zero byte
            m |= ((((z >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * q);
        }
    }
    else {
        for (int i = 0; i < 32 && pos + i < n; ++i) m |= (uint32_t)(p[pos + i] == '\n') << i;
    }
    return m;
}

// pass 1: newlines per block
__global__ void __launch_bounds__(256)
fq_count_kernel(const uint8_t *__restrict__ raw, uint64_t n, uint32_t *__restrict__ block_count) {
    __shared__ uint32_t s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    const uint64_t pos = (uint64_t)blockIdx.x * FQ_BLOCK_BYTES + threadIdx.x * 32;
    uint32_t c = pos < n ? __popc(newline_mask32(raw, pos, n)) : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s, c);
    __syncthreads();
    if (threadIdx.x == 0) block_count[blockIdx.x] = s;
}

// pass 2: positions of the newlines, in order (block_start = exclusive scan of block_count)
__global__ void __launch_bounds__(256)
fq_positions_kernel(const uint8_t *__restrict__ raw, uint64_t n, const uint32_t *__restrict__ block_start,
                    uint32_t *__restrict__ nl) {
    __shared__ uint32_t warp_base[8];
    const uint64_t pos = (uint64_t)blockIdx.x * FQ_BLOCK_BYTES + threadIdx.x * 32;
    const uint32_t m = pos < n ? newline_mask32(raw, pos, n) : 0;
    const uint32_t c = __popc(m), lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = c;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t x = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += x;
    }
    if (lane == 31) warp_base[wid] = incl;
    __syncthreads();
    uint32_t base = block_start[blockIdx.x];
    for (uint32_t w = 0; w < wid; ++w) base += warp_base[w];
    uint32_t out = base + incl - c, mm = m;
    while (mm) {
        const int b = __ffs(mm) - 1;
        mm &= mm - 1;
        nl[out++] = (uint32_t)(pos + b);
    }
}

struct FastqChunkInfo {
    unsigned long long consumed;   // bytes up to the end of the last complete record
    unsigned long long bad_header; // position of the first header that does not start with '@' (or ~0)
    unsigned long long min_len, max_len, windows_ub; // of the sequence lines (batch hint)
};

__device__ __forceinline__ bool fq_space(uint8_t c) {
    return c == '\n' || c == '\r' || c == ' ' || c == '\t' || c == '\v' || c == '\f';
}

// one thread per record: header check, sequence line trimmed on the right
__global__ void __launch_bounds__(256)
fq_records_kernel(const uint8_t *__restrict__ raw, const uint32_t *__restrict__ nl, uint64_t n_records, uint32_t k,
                  uint32_t *__restrict__ seq_start, uint64_t *__restrict__ seq_len, FastqChunkInfo *info) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long lmin = ~0ull, lmax = 0, wub = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_records; r += stride) {
        const uint32_t h = r ? nl[4 * r - 1] + 1 : 0;
        if (raw[h] != '@') atomicMin(&info->bad_header, (unsigned long long)h);
        const uint32_t s = nl[4 * r] + 1;
        uint32_t e = nl[4 * r + 1];
        while (e > s && fq_space(raw[e - 1])) --e;
        seq_start[r] = s;
        const unsigned long long len = e - s;
        seq_len[r] = len;
        lmin = len < lmin ? len : lmin;
        lmax = len > lmax ? len : lmax;
        wub += len >= k ? len - k + 1 : 0;
        if (r == n_records - 1) info->consumed = (unsigned long long)nl[4 * r + 3] + 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xFFFFFFFFu, lmin, o), b = __shfl_xor_sync(0xFFFFFFFFu, lmax, o);
        lmin = a < lmin ? a : lmin;
        lmax = b > lmax ? b : lmax;
        wub += __shfl_xor_sync(0xFFFFFFFFu, wub, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&info->min_len, lmin);
        atomicMax(&info->max_len, lmax);
        if (wub) atomicAdd(&info->windows_ub, wub);
    }
}

// one warp per read: dense[offsets[r] + i] = raw[seq_start[r] + i]
__global__ void __launch_bounds__(256)
fq_gather_kernel(const uint8_t *__restrict__ raw, const uint32_t *__restrict__ seq_start,
                 const uint64_t *__restrict__ offsets, uint64_t n_records, uint8_t *__restrict__ dense) {
    const uint64_t warps = (uint64_t)gridDim.x * blockDim.x / 32;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32; r < n_records; r += warps) {
        const uint64_t o = offsets[r], len = offsets[r + 1] - o;
        const uint8_t *src = raw + seq_start[r];
        for (uint64_t i = lane; i < len; i += 32) dense[o + i] = src[i];
    }
}

} // namespace ktg
