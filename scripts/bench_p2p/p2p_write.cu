// Micro-benchmark: NVLink write bandwidth from a kernel, by store shape.
// Single process, devices 0 -> 1.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p_write.cu -o p2p_write
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int BYTES> __global__ void store_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n) {
    // every warp copies contiguous chunks of 32 * BYTES
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * BYTES;
    const size_t stride = (size_t)gridDim.x * blockDim.x * BYTES;
    for (; i + BYTES <= n; i += stride) {
        if (BYTES == 8) *(uint2 *)(dst + i) = *(const uint2 *)(src + i);
        else *(uint4 *)(dst + i) = *(const uint4 *)(src + i);
    }
}

// tile-wise: load 16 KB into shared memory, then write it out in runs of RUN bytes with 8-byte stores
template <int RUN> __global__ void tile_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n) {
    __shared__ __align__(16) uint8_t sm[16384];
    const size_t n_tiles = n / 16384;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int i = threadIdx.x * 16; i < 16384; i += blockDim.x * 16) *(uint4 *)(sm + i) = *(const uint4 *)(src + t * 16384 + i);
        __syncthreads();
        for (int i = threadIdx.x * 8; i < 16384; i += blockDim.x * 8) *(uint2 *)(dst + t * 16384 + i) = *(const uint2 *)(sm + i);
        __syncthreads();
    }
}

// bulk (TMA) store: one thread issues cp.async.bulk.global.shared::cta of CH bytes
template <int CH> __global__ void bulk_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, size_t n) {
    __shared__ __align__(128) uint8_t sm[16384];
    const size_t n_tiles = n / 16384;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int i = threadIdx.x * 16; i < 16384; i += blockDim.x * 16) *(uint4 *)(sm + i) = *(const uint4 *)(src + t * 16384 + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 16384 / CH) {
            uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * CH);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t * 16384 + (size_t)threadIdx.x * CH), "r"(s), "r"(CH) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t n = 1ull << 30;
    uint8_t *src, *loc, *rem;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, n));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&src, n)); CK(cudaMalloc(&loc, n)); CK(cudaMemset(src, 1, n));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, auto launch) {
        for (int target = 0; target < 2; ++target) {
            uint8_t *dst = target ? rem : loc;
            launch(dst); cudaDeviceSynchronize();
            cudaEventRecord(a);
            for (int it = 0; it < 5; ++it) launch(dst);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("%-28s %-6s %8.1f GB/s  (%s)\n", name, target ? "peer" : "local", 5.0 * n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    };
    run("store 8B/lane  grid 148x8", [&](uint8_t *d) { store_kernel<8><<<148 * 8, 256>>>(src, d, n); });
    run("store 16B/lane grid 148x8", [&](uint8_t *d) { store_kernel<16><<<148 * 8, 256>>>(src, d, n); });
    run("store 8B/lane  grid 148x3", [&](uint8_t *d) { store_kernel<8><<<148 * 3, 256>>>(src, d, n); });
    run("smem tile, 8B stores  x3", [&](uint8_t *d) { tile_kernel<0><<<148 * 3, 256>>>(src, d, n); });
    run("smem tile, 8B stores  x8", [&](uint8_t *d) { tile_kernel<0><<<148 * 8, 256>>>(src, d, n); });
    run("bulk store 4 KB       x3", [&](uint8_t *d) { bulk_kernel<4096><<<148 * 3, 256>>>(src, d, n); });
    run("bulk store 1 KB       x3", [&](uint8_t *d) { bulk_kernel<1024><<<148 * 3, 256>>>(src, d, n); });
    run("bulk store 4 KB       x8", [&](uint8_t *d) { bulk_kernel<4096><<<148 * 8, 256>>>(src, d, n); });
    {
        cudaEventRecord(a);
        for (int it = 0; it < 5; ++it) cudaMemcpyPeerAsync(rem, 1, src, 0, n);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-28s %-6s %8.1f GB/s\n", "cudaMemcpyPeerAsync", "peer", 5.0 * n / ms / 1e6);
    }
    return 0;
}
