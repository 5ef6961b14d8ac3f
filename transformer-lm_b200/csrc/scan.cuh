// scan.cuh -- small hand-written exclusive prefix sums used for compaction (u32 counts -> u64 offsets).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__

__device__ __forceinline__ u32 warp_incl_scan_u32(u32 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (u32)d) v += t;
    }
    return v;
}
__device__ __forceinline__ u64 warp_incl_scan_u64(u64 v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (u32)d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one u32 per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *total receives the block sum (valid for all threads).
__device__ __forceinline__ u32 block_excl_scan_u32(u32 v, u32 *total, u32 *s_warp /* >= 33 u32 */) {
    u32 inc = warp_incl_scan_u32(v);
    u32 w = threadIdx.x >> 5, l = lane_id();
    if (l == 31) s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 nw = (blockDim.x + 31) >> 5;
        u32 x = l < nw ? s_warp[l] : 0;
        u32 xi = warp_incl_scan_u32(x);
        s_warp[l] = xi - x;
        if (l == 31) s_warp[32] = xi;
    }
    __syncthreads();
    u32 res = inc - v + s_warp[w];
    *total = s_warp[32];
    __syncthreads();
    return res;
}

#define SCAN_ITEMS 4
#define SCAN_NT 1024
#define SCAN_TILE (SCAN_ITEMS * SCAN_NT)

static __global__ void __launch_bounds__(SCAN_NT) k_scan_blocksums(const u32 *__restrict__ in, u64 n, u64 *__restrict__ bsum) {
    __shared__ u32 s_warp[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) if (base + k < n) v += in[base + k];
    u32 tot;
    block_excl_scan_u32(v, &tot, s_warp);
    if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single block: exclusive scan of bsum[0..nb) in place; bsum[nb] = grand total
static __global__ void __launch_bounds__(1024) k_scan_blocksums_scan(u64 *bsum, u64 nb) {
    __shared__ u64 s_w[33];
    __shared__ u64 s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (u64 base = 0; base < nb; base += 1024) {
        u64 i = base + threadIdx.x;
        u64 v = i < nb ? bsum[i] : 0;
        u64 inc = warp_incl_scan_u64(v);
        u32 w = threadIdx.x >> 5, l = lane_id();
        if (l == 31) s_w[w] = inc;
        __syncthreads();
        if (w == 0) {
            u64 x = s_w[l];
            u64 xi = warp_incl_scan_u64(x);
            s_w[l] = xi - x;
            if (l == 31) s_w[32] = xi;
        }
        __syncthreads();
        u64 carry = s_carry;
        if (i < nb) bsum[i] = carry + inc - v + s_w[w];
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[nb] = s_carry;
}

static __global__ void __launch_bounds__(SCAN_NT) k_scan_apply(const u32 *__restrict__ in, u64 n, const u64 *__restrict__ bsum,
                                                              u64 *__restrict__ out) {
    __shared__ u32 s_warp[33];
    u64 base = (u64)blockIdx.x * SCAN_TILE + (u64)threadIdx.x * SCAN_ITEMS;
    u32 x[SCAN_ITEMS]; u32 v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { x[k] = base + k < n ? in[base + k] : 0; v += x[k]; }
    u32 tot;
    u32 ex = block_excl_scan_u32(v, &tot, s_warp);
    u64 o = bsum[blockIdx.x] + ex;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) { if (base + k < n) out[base + k] = o; o += x[k]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = bsum[gridDim.x];
}

// out must hold n+1 u64; bsum_tmp must hold (ceil(n/SCAN_TILE)+1) u64.  out[n] = total.
static inline void launch_excl_scan_u32_to_u64(const u32 *in, u64 n, u64 *out, u64 *bsum_tmp, cudaStream_t st) {
    if (n == 0) { cudaMemsetAsync(out, 0, sizeof(u64), st); return; }
    u64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    KLAUNCH(k_scan_blocksums, (unsigned)nb, SCAN_NT, 0, st, in, n, bsum_tmp);
    KLAUNCH(k_scan_blocksums_scan, 1, 1024, 0, st, bsum_tmp, nb);
    KLAUNCH(k_scan_apply, (unsigned)nb, SCAN_NT, 0, st, in, n, bsum_tmp, out);
}
static inline size_t scan_tmp_elems(u64 n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE + 2); }

#endif
