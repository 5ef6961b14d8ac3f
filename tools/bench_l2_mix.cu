// Microbenchmark (B200): a dense hot table (bucket = one 32-byte sector of two 16-byte keys, counts in a separate array) probed
// with probability P_HOT, a sparse cold table (one 32-byte slot per 128-byte line, read + RED) otherwise, 16 bytes of streaming
// text per access -- the access mix of the count stage -- with and without L2 eviction-priority hints.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/bench_l2_mix tools/bench_l2_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
__device__ __forceinline__ ulonglong2 ld_hint(const void *p, u64 pol) { ulonglong2 r; asm volatile("ld.global.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(pol)); return r; }
__device__ __forceinline__ void red_hint(u64 *p, u64 v, u64 pol) { asm volatile("red.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory"); }
template <int HINT>
__global__ void k(ulonglong2 *hot, u64 *hcnt, u64 n_buckets, u64 *cold, u64 n_cold, const uint4 *stream, u64 n_stream, int iters, int p_hot, u64 seed, u64 *sink) {
    u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x, acc = 0, nt = (u64)gridDim.x * blockDim.x;
    u64 pl, pf;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pl));
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pf));
    for (int i = 0; i < iters; i++) {
        u64 r = mix(seed + t * 1315423911ull + i);
        uint4 s = __ldcs(stream + ((u64)i * nt + t) % n_stream);
        acc += s.x;
        if ((int)(r % 100) < p_hot) {
            u64 b = (r >> 8) % n_buckets;
            ulonglong2 a0, a1;
            if (HINT) { a0 = ld_hint(hot + 2 * b, pl); a1 = ld_hint(hot + 2 * b + 1, pl); } else { a0 = hot[2 * b]; a1 = hot[2 * b + 1]; }
            u64 which = (a0.x + a1.x + r) & 1;
            if (HINT) red_hint(hcnt + 2 * b + which, 1, pl); else atomicAdd(hcnt + 2 * b + which, 1ull);
        } else {
            u64 c = ((r >> 8) % n_cold) * 16;
            ulonglong2 a;
            if (HINT) a = ld_hint(cold + c, pf); else a = *(const ulonglong2 *)(cold + c);
            if (HINT) red_hint(cold + c + 2, 1 + (a.x & 0), pf); else atomicAdd(cold + c + 2, 1ull + (a.x & 0));
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    ulonglong2 *hot; u64 *hcnt, *cold, *sink; uint4 *stream;
    const u64 n_cold = 24ull << 20;                      // 3 GB of lines
    const u64 n_stream = (1ull << 30) / 16;
    cudaMalloc(&hot, 256ull << 20); cudaMalloc(&hcnt, 128ull << 20); cudaMalloc(&cold, n_cold * 128); cudaMalloc(&stream, n_stream * 16); cudaMalloc(&sink, 8);
    cudaMemset(hot, 0, 256ull << 20); cudaMemset(hcnt, 0, 128ull << 20); cudaMemset(cold, 0, n_cold * 128); cudaMemset(stream, 0, n_stream * 16);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 96ull << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, block = 256, iters = 128;
    for (int p_hot : {78, 100, 0})
        for (u64 nb : {1ull << 18, 1ull << 19, 1ull << 20, 1ull << 21})
            for (int hint = 0; hint < 2; hint++) {
                float best = 1e30f;
                for (int rep = 0; rep < 4; rep++) {
                    cudaEventRecord(e0);
                    if (hint) k<1><<<grid, block>>>(hot, hcnt, nb, cold, n_cold, stream, n_stream, iters, p_hot, rep * 7919, sink);
                    else k<0><<<grid, block>>>(hot, hcnt, nb, cold, n_cold, stream, n_stream, iters, p_hot, rep * 7919, sink);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
                }
                printf("p_hot %3d buckets %8llu (keys %5.1f MB + counts %5.1f MB) hints %d: %7.2f G accesses/s\n", p_hot, nb, nb * 32 / 1e6, nb * 16 / 1e6, hint,
                       (double)grid * block * iters / best / 1e6);
            }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
