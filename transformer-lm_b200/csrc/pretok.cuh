// pretok.cuh -- GPT-2 pretokenizer as a local stencil on UTF-8 bytes (device side).
//
// Replaces regex.finditer with
//   '(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
// (reference: models/tokenizer/train.py:143-146 == models/tokenizer/tokenizer.py:26-27).
// Pretoken starts are a function of <= 4 code points of left context and <= 2 of right context
// (SURVEY Appendix B), so every byte decides independently whether a pretoken starts on it.
//
// Classes: 0 S=\s, 1 L=\p{L}, 2 N=\p{N}, 3 P=other, 4 B=boundary (padding / special-token bytes).
#pragma once
#include "common.cuh"
#include "unicode_tables.h"

#define CLS_S 0u
#define CLS_L 1u
#define CLS_N 2u
#define CLS_P 3u
#define CLS_B 4u

#define PT_NT 256                 // threads per CTA
#define PT_CHUNK 16               // bytes per thread
#define PT_TILE (PT_NT * PT_CHUNK)

#ifdef __CUDACC__

__device__ __constant__ uint32_t c_uc_pages[BPE_UC_NPAGES * 16];
__device__ __constant__ uint8_t c_uc_index[BPE_UC_NINDEX];
__device__ __constant__ uint8_t c_uc_ascii[128];

// Decode the code point whose lead byte is b0 (>= 0xC0) followed by b1,b2,b3.
// Returns its length; *ok = strict validity (CPython's utf-8 codec rules); *cp = code point.
__device__ __forceinline__ int utf8_decode_multi(u32 b0, u32 b1, u32 b2, u32 b3, u32 *cp, bool *ok) {
    bool c1 = (b1 & 0xC0u) == 0x80u, c2 = (b2 & 0xC0u) == 0x80u, c3 = (b3 & 0xC0u) == 0x80u;
    if (b0 < 0xE0u) {
        *cp = ((b0 & 0x1Fu) << 6) | (b1 & 0x3Fu);
        *ok = b0 >= 0xC2u && c1;
        return 2;
    }
    if (b0 < 0xF0u) {
        *cp = ((b0 & 0x0Fu) << 12) | ((b1 & 0x3Fu) << 6) | (b2 & 0x3Fu);
        *ok = c1 && c2 && !(b0 == 0xE0u && b1 < 0xA0u) && !(b0 == 0xEDu && b1 >= 0xA0u);
        return 3;
    }
    *cp = ((b0 & 0x07u) << 18) | ((b1 & 0x3Fu) << 12) | ((b2 & 0x3Fu) << 6) | (b3 & 0x3Fu);
    *ok = b0 < 0xF5u && c1 && c2 && c3 && !(b0 == 0xF0u && b1 < 0x90u) && !(b0 == 0xF4u && b1 >= 0x90u);
    return 4;
}

__device__ __forceinline__ bool is_sdmt(u32 b) { return b == 's' || b == 'd' || b == 'm' || b == 't'; }
__device__ __forceinline__ bool is_llvere(u32 x, u32 y) {
    return (x == 'l' && y == 'l') || (x == 'v' && y == 'e') || (x == 'r' && y == 'e');
}

// =============================================================================================
// Bit-parallel formulation of SURVEY Appendix B.1 on bytes:
//  - whitespace run: start at its first char, and at its last char when the run (len >= 2) is followed by \S
//  - L / N / P run: start at its first char unless the previous char is U+0020
//  - contractions 's 'd 'm 't 'll 've 're: letters inside are not starts, the first letter after is
//  - a char after a boundary byte (text start, after a special token) is always a start
// A continuation byte carries the class of its character and is never a start.
//
// Per text byte one "info" byte, one-hot:  bit0 S  bit1 L  bit2 N  bit3 P  bit4 B(boundary)  bit5 LEAD
//                                          bit6 the byte is U+0020  bit7 the byte is an apostrophe
// ASCII bytes get theirs from a 256-entry table (bytes >= 0x80 map to 0 and are patched by a loop that runs once per
// non-ASCII character).  The 16 info bytes of a chunk are transposed into eight 16-bit masks; with the neighbours'
// masks they form 32-bit windows [8 bytes before | 16 | 8 after] on which the start rules are ~40 bitwise operations.
// =============================================================================================
#define INF_S 0x01u
#define INF_L 0x02u
#define INF_N 0x04u
#define INF_P 0x08u
#define INF_B 0x10u
#define INF_LEAD 0x20u
#define INF_SP 0x40u
#define INF_AP 0x80u

struct PretokTables {            // shared memory: Unicode class pages + index (for non-ASCII) and the ASCII info table
    uint32_t pages[BPE_UC_NPAGES * 16];
    uint8_t index[BPE_UC_NINDEX];
    uint8_t info[256];
};

__device__ __forceinline__ void pretok_load_tables(PretokTables *t) {
    for (int i = threadIdx.x; i < BPE_UC_NPAGES * 16; i += blockDim.x) t->pages[i] = c_uc_pages[i];
    for (int i = threadIdx.x; i < BPE_UC_NINDEX; i += blockDim.x) t->index[i] = c_uc_index[i];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t v = 0;
        if (i < 128) v = (1u << c_uc_ascii[i]) | INF_LEAD | (i == 0x20 ? INF_SP : 0u) | (i == 0x27 ? INF_AP : 0u);
        t->info[i] = (uint8_t)v;
    }
}
__device__ __forceinline__ u32 cp_class_smem(const PretokTables *t, u32 cp) {
    if (cp >= 0x110000u) return CLS_P;
    u32 page = t->index[cp >> 8];
    u32 w = t->pages[page * 16 + ((cp & 255u) >> 4)];
    return (w >> (2 * (cp & 15u))) & 3u;
}
// bit k of each of the 4 bytes of w -> 4 adjacent bits
__device__ __forceinline__ u32 movemask4(u32 w, u32 k) { return ((((w >> k) & 0x01010101u) * 0x00204081u) >> 21) & 0xFu; }
__device__ __forceinline__ bool has_byte(u32 w, u32 b) { u32 x = w ^ (b * 0x01010101u); return ((x - 0x01010101u) & ~x & 0x80808080u) != 0; }

// Info bytes of the ASCII bytes of the 16-byte chunk at byte offset `at` of the shared-memory text tile `tx` (bytes >= 0x80 get
// 0 and are patched by chunk_patch_non_ascii).  *hi: mask of the chunk's bytes >= 0x80; *has_cr: the chunk contains '\r'.
__device__ __forceinline__ uint4 chunk_info_ascii(const PretokTables *tb, const uint8_t *tx, u32 at, u32 *hi_out, bool *has_cr) {
    const uint4 c = *reinterpret_cast<const uint4 *>(tx + at);
    const u32 w[4] = {c.x, c.y, c.z, c.w};
    u32 iw[4];
    u32 hi = 0;
    bool cr = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const u32 x = w[q];
        iw[q] = (u32)tb->info[x & 0xFFu] | ((u32)tb->info[(x >> 8) & 0xFFu] << 8) | ((u32)tb->info[(x >> 16) & 0xFFu] << 16) |
                ((u32)tb->info[x >> 24] << 24);
        hi |= movemask4(x, 7) << (4 * q);
        cr |= has_byte(x, 0x0Du);
    }
    *hi_out = hi; *has_cr = cr;
    return make_uint4(iw[0], iw[1], iw[2], iw[3]);
}
// The non-ASCII bytes (mask hi) of that chunk: leads are decoded, validated and classified through the Unicode table in shared
// memory, continuation bytes inherit the class of their character.  `tx` needs at least 3 readable bytes before and 4 after the
// chunk.  Returns the smallest chunk-relative offset (0..15) of an ill-formed sequence that STARTS in this chunk, or 0xFF.
__device__ __forceinline__ u32 chunk_patch_non_ascii(const PretokTables *tb, const uint8_t *tx, u32 at, u32 hi, uint4 *info) {
    u32 iw[4] = {info->x, info->y, info->z, info->w};
    u32 e = 0xFFu;
    int carry_left = 0; u32 carry_cls = 0, expect = 0;
    if ((tx[at] & 0xC0u) == 0x80u) {                               // first byte continues a character of the previous chunk?
#pragma unroll
        for (int d = 1; d <= 3; d++) {
            const u32 l0 = tx[at - d];
            if (carry_left == 0 && l0 >= 0xC0u && l0 < 0xF8u) {
                u32 cp; bool ok;
                int len = utf8_decode_multi(l0, tx[at - d + 1], tx[at - d + 2], tx[at - d + 3], &cp, &ok);
                if (ok && len > d) { carry_left = len - d; carry_cls = cp_class_smem(tb, cp); expect = 0; }
            }
        }
    }
    u32 m = hi;
    while (m) {
        const u32 j = __ffs(m) - 1; m &= m - 1;
        const u32 b = tx[at + j];
        u32 inf;
        if (b >= 0xF8u) {                                          // never occurs in UTF-8: padding / boundary
            inf = INF_B | INF_LEAD; if (e == 0xFFu) e = j; carry_left = 0;
        } else if (b >= 0xC0u) {
            u32 cp; bool ok;
            int len = utf8_decode_multi(b, tx[at + j + 1], tx[at + j + 2], tx[at + j + 3], &cp, &ok);
            carry_cls = ok ? cp_class_smem(tb, cp) : CLS_P;
            inf = (1u << carry_cls) | INF_LEAD;
            carry_left = ok ? len - 1 : 0; expect = j + 1;
            if (!ok && e == 0xFFu) e = j;
        } else if (carry_left > 0 && j == expect) {                // continuation byte of the character in progress
            inf = 1u << carry_cls; carry_left--; expect++;
        } else {                                                   // orphan continuation byte (or tail of an invalid sequence)
            inf = INF_P; if (e == 0xFFu) e = j; carry_left = 0;
        }
        iw[j >> 2] |= inf << ((j & 3u) * 8u);
    }
    *info = make_uint4(iw[0], iw[1], iw[2], iw[3]);
    return e;
}

// 16 info bytes -> eight 16-bit masks packed as {S|L<<16, N|P<<16, B|LEAD<<16, SP|AP<<16}: bit i of mask k = bit k of info byte i.
// Two 8 x 8 bit-matrix transposes (three masked swap steps on a 64-bit word each: byte k of the result collects bit k of the eight
// input bytes), then byte k of both halves side by side is mask k.  ~50 instructions; extracting the eight bit planes one by one
// with multiply-and-shift cost ~170 and was a third of all instructions of the flags kernel.
__device__ __forceinline__ u64 transpose8x8(u64 x) {
    u64 t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAull; x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCull; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ull; x = x ^ t ^ (t << 28);
    return x;
}
__device__ __forceinline__ uint4 info_to_masks(uint4 inf) {
    const u64 a = transpose8x8((u64)inf.x | ((u64)inf.y << 32)), b = transpose8x8((u64)inf.z | ((u64)inf.w << 32));
    const u32 al = (u32)a, ah = (u32)(a >> 32), bl = (u32)b, bh = (u32)(b >> 32);
    return make_uint4(__byte_perm(al, bl, 0x5140), __byte_perm(al, bl, 0x7362), __byte_perm(ah, bh, 0x5140), __byte_perm(ah, bh, 0x7362));
}
// bytes covered by a special-token occurrence become boundary bytes
__device__ __forceinline__ uint4 masks_apply_boundary(uint4 mk, u32 m16) {
    const u32 keep = ~(m16 | (m16 << 16));
    mk.x &= keep; mk.y &= keep; mk.w &= keep;
    mk.z = (mk.z & keep) | m16 | (m16 << 16);
    return mk;
}
// class of the next LEAD byte after each position, for a one-hot class mask X (window form)
__device__ __forceinline__ u32 next_lead_has(u32 X, u32 LEADm) {
    const u32 A = LEADm >> 1;                    // bit i: byte i+1 starts a character
    u32 f = (X >> 1) & A;
    f |= ~A & (f >> 1); f |= ~A & (f >> 1); f |= ~A & (f >> 1);
    return f;
}
__device__ __forceinline__ u32 window32(u32 prev16, u32 cur16, u32 next16) { return (prev16 >> 8) | (cur16 << 8) | ((next16 & 0xFFu) << 24); }

// Start bits of the 16 bytes of a chunk from its masks and its neighbours' (same packing as info_to_masks).
// tx + at = the chunk's text in the shared-memory tile (raw bytes are only needed around apostrophes).
__device__ __forceinline__ u32 flags_from_masks(uint4 pm, uint4 cm, uint4 nm, const uint8_t *tx, u32 at) {
    const u32 S = window32(pm.x & 0xFFFFu, cm.x & 0xFFFFu, nm.x & 0xFFFFu), L = window32(pm.x >> 16, cm.x >> 16, nm.x >> 16);
    const u32 N = window32(pm.y & 0xFFFFu, cm.y & 0xFFFFu, nm.y & 0xFFFFu), P = window32(pm.y >> 16, cm.y >> 16, nm.y >> 16);
    const u32 B = window32(pm.z & 0xFFFFu, cm.z & 0xFFFFu, nm.z & 0xFFFFu), LEADm = window32(pm.z >> 16, cm.z >> 16, nm.z >> 16);
    const u32 SP = window32(pm.w & 0xFFFFu, cm.w & 0xFFFFu, nm.w & 0xFFFFu) & ~B, AP = window32(pm.w >> 16, cm.w >> 16, nm.w >> 16) & P;
    const u32 after_S_or_B = next_lead_has(S | B, LEADm);
    const u32 no_space_before = ~(SP << 1);
    u32 st = S & (~(S << 1) | ~after_S_or_B);                        // whitespace run: first char; last char when followed by \S
    st |= L & ~(L << 1) & no_space_before;                           // L / N / P runs: first char unless a single space attaches in front
    st |= N & ~(N << 1) & no_space_before;
    st |= P & ~(P << 1) & no_space_before;
    st &= LEADm;
    // contractions 's 'd 'm 't 'll 've 're: an apostrophe opens one only if the character before it is neither P nor U+0020
    u32 ap = AP & ~((P | SP) << 1) & 0x007FFFE0u;                    // apostrophes at window bytes 5..22 can touch bytes 8..23
    const u32 Llead = L & LEADm;
    while (ap) {
        const u32 i = __ffs(ap) - 1; ap &= ap - 1;
        const u32 t1 = tx[at + i - 8 + 1], t2 = tx[at + i - 8 + 2], t3 = tx[at + i - 8 + 3];
        const bool c2 = is_sdmt(t1), c3 = is_llvere(t1, t2);
        if ((c2 || c3) && ((Llead >> (i + 1)) & 1u)) st &= ~(1u << (i + 1));          // first letter inside
        if ((Llead >> (i + 2)) & 1u) {
            if (c2) st |= 1u << (i + 2);                                             // first letter after 's 'd 'm 't
            if (is_llvere(t1, t2)) st &= ~(1u << (i + 2));                           // second letter inside
        }
        if (c3 && ((Llead >> (i + 3)) & 1u)) st |= 1u << (i + 3);                    // first letter after 'll 've 're
        (void)t3;
    }
    return (st >> 8) & 0xFFFFu;
}

#endif  // __CUDACC__
