// hashtab.cuh -- device helpers shared by the pretoken tables of training (train.cu) and encoding (encode.cu).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
__device__ __forceinline__ u64 hash_long(const uint8_t *p, u32 len) {
    u64 h = 0x9E3779B97F4A7C15ull ^ len;
    u32 i = 0;
    for (; i + 8 <= len; i += 8) {
        u64 v = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) v |= (u64)p[i + k] << (8 * k);
        h = (h ^ v) * 0x9FB21C651E98DF25ull;
        h ^= h >> 29;
    }
    u64 v = 0;
    for (u32 k = 0; i + k < len; k++) v |= (u64)p[i + k] << (8 * k);
    h = (h ^ v) * 0x9FB21C651E98DF25ull;
    return mix64(h) | 1ull;                      // never 0 ("hash not published yet")
}

__device__ __forceinline__ bool bytes_equal(const uint8_t *a, const uint8_t *b, u32 len) {
    for (u32 i = 0; i < len; i++) if (a[i] != b[i]) return false;
    return true;
}

__device__ __forceinline__ u64 flags_next_start(const u32 *__restrict__ flags, u64 from, u64 n) {
    // first start bit at position >= from (< n), else n
    if (from >= n) return n;
    u64 w = from >> 5;
    u32 cur = flags[w] & (0xFFFFFFFFu << (from & 31));
    for (;;) {
        if (cur) { u64 p = (w << 5) + (__ffs(cur) - 1); return p < n ? p : n; }
        w++;
        if ((w << 5) >= n) return n;
        cur = flags[w];
    }
}

#endif
