// fastq_device.cuh -- FASTQ and FASTA record parsing on the device.
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

// ---- FASTA (create_fasta, builder.rs:118-140; rust-bio 0.10's fasta::Reader restated in host_reader.h):
// a line that starts with '>' opens a record, every other line is appended to the open record after
// trimming trailing whitespace; the first line of a file must be a header.  A record is one read.
// Lines are cut with the same newline index as FASTQ; two scans turn the per-line facts into the dense
// batch: rec_idx = headers before the line, base_off = sequence bytes before the line.

// one thread per line: hdr[i] (starts with '>'), len[i] (trimmed length, 0 for a header)
__global__ void __launch_bounds__(256)
fa_lines_kernel(const uint8_t *__restrict__ raw, const uint32_t *__restrict__ nl, uint64_t n_lines,
                uint32_t *__restrict__ hdr, uint64_t *__restrict__ len) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_lines; i += stride) {
        if (i == n_lines) { // scan sentinels
            hdr[i] = 0;
            len[i] = 0;
            continue;
        }
        const uint32_t s = i ? nl[i - 1] + 1 : 0;
        uint32_t e = nl[i];
        const bool h = raw[s] == '>';
        while (e > s && fq_space(raw[e - 1])) --e;
        hdr[i] = h;
        len[i] = h ? 0 : e - s;
    }
}

// offsets[r] = sequence bytes before record r, r <= n_rec; consumed = where the first record that is
// NOT part of this batch starts (the header after the last complete record)
__global__ void __launch_bounds__(256)
fa_offsets_kernel(const uint32_t *__restrict__ nl, const uint32_t *__restrict__ hdr, const uint32_t *__restrict__ rec_idx,
                  const uint64_t *__restrict__ base_off, uint64_t n_lines, uint64_t n_rec, uint64_t *__restrict__ offsets,
                  FastqChunkInfo *info) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n_lines; i += stride) {
        if (i == n_lines) {
            if (rec_idx[i] == n_rec) offsets[n_rec] = base_off[i]; // the chunk ends with the file
        }
        else if (hdr[i] && rec_idx[i] <= n_rec) {
            offsets[rec_idx[i]] = base_off[i];
            if (rec_idx[i] == n_rec) info->consumed = i ? (unsigned long long)nl[i - 1] + 1 : 0ull;
        }
    }
}

// one warp per sequence line of a complete record
__global__ void __launch_bounds__(256)
fa_gather_kernel(const uint8_t *__restrict__ raw, const uint32_t *__restrict__ nl, const uint32_t *__restrict__ hdr,
                 const uint32_t *__restrict__ rec_idx, const uint64_t *__restrict__ base_off, uint64_t n_lines, uint64_t n_rec,
                 uint8_t *__restrict__ dense) {
    const uint64_t warps = (uint64_t)gridDim.x * blockDim.x / 32;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32; i < n_lines; i += warps) {
        if (hdr[i] || rec_idx[i] == 0 || rec_idx[i] - 1 >= n_rec) continue;
        const uint64_t o = base_off[i], len = base_off[i + 1] - o;
        const uint8_t *src = raw + (i ? nl[i - 1] + 1 : 0);
        for (uint64_t j = lane; j < len; j += 32) dense[o + j] = src[j];
    }
}

// batch hint of the records: one length or ragged, windows if every read is accepted
__global__ void __launch_bounds__(256)
fa_hint_kernel(const uint64_t *__restrict__ offsets, uint64_t n_rec, uint32_t k, FastqChunkInfo *info) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long lmin = ~0ull, lmax = 0, wub = 0;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += stride) {
        const unsigned long long len = offsets[r + 1] - offsets[r];
        lmin = len < lmin ? len : lmin;
        lmax = len > lmax ? len : lmax;
        wub += len >= k ? len - k + 1 : 0;
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

} // namespace ktg
