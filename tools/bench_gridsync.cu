// Micro-benchmark: cost of a grid-wide barrier on B200 (cooperative groups vs a hand-written one).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_gridsync tools/bench_gridsync.cu && /tmp/bench_gridsync
#include <cooperative_groups.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void k_cg(int iters, unsigned long long *sink) {
    cg::grid_group g = cg::this_grid();
    for (int i = 0; i < iters; i++) g.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}

// sense-free counting barrier: every CTA adds 1, waits until the counter reaches (epoch+1)*G
__device__ __forceinline__ void grid_barrier(unsigned int *counter, unsigned int &epoch, unsigned int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += G;
        __threadfence();
        atomicAdd(counter, 1u);
        while (*((volatile unsigned int *)counter) < epoch) { }
        __threadfence();
    }
    __syncthreads();
}
__global__ void k_custom(int iters, unsigned int *counter, unsigned long long *sink) {
    unsigned int epoch = 0;
    for (int i = 0; i < iters; i++) grid_barrier(counter, epoch, gridDim.x);
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
// variant: ld.acquire / red.release instead of fences
__device__ __forceinline__ void grid_barrier2(unsigned int *counter, unsigned int &epoch, unsigned int G) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += G;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < epoch);
    }
    __syncthreads();
}
__global__ void k_custom2(int iters, unsigned int *counter, unsigned long long *sink) {
    unsigned int epoch = 0;
    for (int i = 0; i < iters; i++) grid_barrier2(counter, epoch, gridDim.x);
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}

// ---- cluster variants: hardware cluster barrier, and a hierarchical grid barrier (cluster barrier + one global
// arrival per cluster) ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned int cluster_ctarank() { unsigned int r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__global__ void k_cluster(int iters, unsigned long long *sink) {
    for (int i = 0; i < iters; i++) cluster_sync_all();
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
__global__ void k_hier(int iters, unsigned int *counter, unsigned int n_clusters, unsigned long long *sink) {
    unsigned int epoch = 0;
    const bool leader = cluster_ctarank() == 0;
    for (int i = 0; i < iters; i++) {
        cluster_sync_all();
        if (leader && threadIdx.x == 0) {
            epoch += n_clusters;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            unsigned int v;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < epoch);
        }
        cluster_sync_all();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = 1;
}
static float run_cluster(void *fn, int grid, int nt, int csize, void **args) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(nt);
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, args);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("  launch failed: %s\n", cudaGetErrorString(e)); return -1; }
    return ms;
}

int main() {
    int iters = 20000;
    unsigned long long *sink; unsigned int *counter;
    cudaMalloc(&sink, 8); cudaMalloc(&counter, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int nt : {128, 256, 512, 1024}) {
        for (int g : {sms / 4, sms}) {
            float ms[3];
            for (int v = 0; v < 3; v++) {
                cudaMemset(counter, 0, 4);
                void *a0[] = {&iters, &sink};
                void *a1[] = {&iters, &counter, &sink};
                cudaEventRecord(e0);
                if (v == 0) cudaLaunchCooperativeKernel((void *)k_cg, dim3(g), dim3(nt), a0, 0, 0);
                if (v == 1) cudaLaunchCooperativeKernel((void *)k_custom, dim3(g), dim3(nt), a1, 0, 0);
                if (v == 2) cudaLaunchCooperativeKernel((void *)k_custom2, dim3(g), dim3(nt), a1, 0, 0);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                cudaEventElapsedTime(&ms[v], e0, e1);
            }
            printf("G=%3d NT=%4d: cg %.2f us  custom(fence) %.2f us  custom(acq/rel) %.2f us   err=%s\n", g, nt,
                   1e3 * ms[0] / iters, 1e3 * ms[1] / iters, 1e3 * ms[2] / iters, cudaGetErrorString(cudaGetLastError()));
        }
    }
    for (int nt : {512, 1024}) {
        for (int cs : {8, 16}) {
            void *a0[] = {&iters, &sink};
            float ms = run_cluster((void *)k_cluster, cs, nt, cs, a0);
            printf("single cluster of %2d CTAs NT=%4d: cluster barrier %.3f us\n", cs, nt, 1e3 * ms / iters);
            unsigned int ncl = (unsigned)(sms / cs);
            int grid = (int)ncl * cs;
            cudaMemset(counter, 0, 4);
            void *a1[] = {&iters, &counter, &ncl, &sink};
            ms = run_cluster((void *)k_hier, grid, nt, cs, a1);
            printf("hierarchical: %u clusters x %d CTAs (grid %d) NT=%4d: %.3f us\n", ncl, cs, grid, nt, 1e3 * ms / iters);
        }
    }
    for (int g : {8, 16, 32, 64}) {
        cudaMemset(counter, 0, 4);
        void *a1[] = {&iters, &counter, &sink};
        cudaEventRecord(e0);
        cudaLaunchCooperativeKernel((void *)k_custom2, dim3(g), dim3(512), a1, 0, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("G=%3d NT= 512: custom(acq/rel) %.2f us\n", g, 1e3 * ms / iters);
    }
    return 0;
}
