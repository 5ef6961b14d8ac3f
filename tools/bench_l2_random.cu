// Microbenchmark (B200): throughput of random single-sector accesses to a working set of W entries laid out `stride` bytes apart
// (32 = dense sectors, 128 = one sector per L2 line), read only or read + RED on the same sector.  Guides the layout of the
// pretoken tables: how many entries does L2 really hold, and what does a DRAM-bound probe cost?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/bench_l2_random tools/bench_l2_random.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int MODE>   // 0 read, 1 read + red, 2 red only
__global__ void k(u64 *tab, u64 W, u64 stride8, int iters, u64 seed, u64 *sink) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x, acc = 0;
    for (int i = 0; i < iters; i += 4) {
        u64 a[4], v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = (mix(seed + t * 1315423911ull + i + j) % W) * stride8;
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = MODE != 2 ? tab[a[j]] : 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { acc += v[j]; if (MODE) atomicAdd(&tab[a[j] + 1], 1ull + (v[j] & 0)); }
    }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    u64 *tab, *sink; size_t bytes = 4ull << 30;
    cudaMalloc(&tab, bytes); cudaMalloc(&sink, 8); cudaMemset(tab, 0, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256, iters = 256;
    for (int mode = 0; mode < 3; mode++)
        for (u64 stride : {32ull, 128ull})
            for (u64 W : {1ull << 16, 1ull << 18, 1ull << 19, 1ull << 20, 1ull << 21, 1ull << 22, 1ull << 23, 1ull << 24, 1ull << 25}) {
                if (W * stride > bytes) continue;
                float best = 1e30f;
                for (int rep = 0; rep < 3; rep++) {
                    cudaEventRecord(e0);
                    if (mode == 0) k<0><<<grid, block>>>(tab, W, stride / 8, iters, rep * 7919, sink);
                    else if (mode == 1) k<1><<<grid, block>>>(tab, W, stride / 8, iters, rep * 7919, sink);
                    else k<2><<<grid, block>>>(tab, W, stride / 8, iters, rep * 7919, sink);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
                }
                double ops = (double)grid * block * iters;
                printf("mode %d (%s) stride %3llu W %9llu (%7.1f MB of lines, %7.1f MB of sectors): %7.2f G accesses/s\n", mode, mode == 0 ? "read" : mode == 1 ? "read+red" : "red", stride,
                       W, W * (stride < 128 ? stride : 128) / 1e6, W * 32 / 1e6, ops / best / 1e6);
            }
    return cudaGetLastError() != cudaSuccess;
}
