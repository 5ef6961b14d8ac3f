// hashtab.cuh -- device helpers shared by the pretoken tables of training (train.cu) and encoding (encode.cu).
#pragma once
#include "common.cuh"

#ifdef __CUDACC__
// 8 bytes at an arbitrary address through aligned 64-bit loads (up to 15 bytes past p are touched: every buffer
// these helpers see -- text arena, key pools -- is padded accordingly).  Byte loops cost ~4 instructions per byte
// and diverge on the length; this is ~8 instructions per 8 bytes.
__device__ __forceinline__ u64 load8_unaligned(const uint8_t *p) {
    const u64 *q = reinterpret_cast<const u64 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)7);
    const u32 sh = (u32)(reinterpret_cast<uintptr_t>(p) & 7u) * 8u;
    const u64 lo = q[0], hi = q[1];
    return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
}
__device__ __forceinline__ u64 low_bytes_mask(u32 n_bytes) { return n_bytes >= 8 ? ~0ull : ((1ull << (8u * n_bytes)) - 1ull); }

__device__ __forceinline__ u64 hash_long(const uint8_t *p, u32 len) {
    u64 h = 0x9E3779B97F4A7C15ull ^ len;
    u32 i = 0;
    for (; i + 8 <= len; i += 8) {
        h = (h ^ load8_unaligned(p + i)) * 0x9FB21C651E98DF25ull;
        h ^= h >> 29;
    }
    const u64 v = i < len ? load8_unaligned(p + i) & low_bytes_mask(len - i) : 0;
    h = (h ^ v) * 0x9FB21C651E98DF25ull;
    return mix64(h) | 1ull;                      // never 0 ("hash not published yet")
}

__device__ __forceinline__ bool bytes_equal(const uint8_t *a, const uint8_t *b, u32 len) {
    u32 i = 0;
    for (; i + 8 <= len; i += 8) if (load8_unaligned(a + i) != load8_unaligned(b + i)) return false;
    if (i < len) return ((load8_unaligned(a + i) ^ load8_unaligned(b + i)) & low_bytes_mask(len - i)) == 0;
    return true;
}

// key of a pretoken of <= 7 bytes: the bytes little-endian in the low 56 bits, the length in the top byte
__device__ __forceinline__ u64 short_key(const uint8_t *p, u32 len) {
    return (load8_unaligned(p) & low_bytes_mask(len)) | ((u64)len << 56);
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
