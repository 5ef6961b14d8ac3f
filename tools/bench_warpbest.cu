// Micro-benchmark: cycles per warp arg-max (redux-based warp_best vs the shuffle butterfly) on B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I transformer-lm_b200/csrc -o /tmp/bench_warpbest tools/bench_warpbest.cu
#include <cstdio>
#include "../transformer-lm_b200/csrc/merge.cuh"
unsigned long long g_bpe_launches = 0;

template <int MODE>
__global__ void k(int iters, int tie_mode, unsigned long long *out) {
    u32 lane = threadIdx.x & 31;
    Best b;
    b.cnt = tie_mode ? 7 : (i64)(lane * 3 % 17);             // tie_mode: every lane ties on the count
    b.key = ((u64)(300 + lane) << 32) | (400 + lane);
    b.ka = mix64(lane + 1) | (1ull << 63); b.kb = mix64(lane + 77);
    if (tie_mode == 2) b.ka = 12345;                         // ties on count and first token prefix (same token needed)
    if (tie_mode == 2) b.key = ((u64)300 << 32) | (400 + lane);
    long long t0 = clock64();
    u64 acc = 0;
    for (int i = 0; i < iters; i++) {
        Best r = MODE == 0 ? warp_best(b) : warp_best_butterfly(b);
        acc += r.key;
        b.cnt += (i64)(r.key & 1);                          // dependent chain
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = acc; }
}
int main() {
    unsigned long long *out, h[2];
    cudaMalloc(&out, 16);
    int iters = 10000;
    for (int tie = 0; tie < 3; tie++) {
        k<0><<<1, 32>>>(iters, tie, out); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        double a = (double)h[0] / iters;
        k<1><<<1, 32>>>(iters, tie, out); cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
        double b = (double)h[0] / iters;
        printf("tie_mode %d: redux warp_best %.0f cycles, butterfly %.0f cycles  (%s)\n", tie, a, b, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
