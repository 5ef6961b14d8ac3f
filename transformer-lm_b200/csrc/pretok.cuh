// pretok.cuh -- GPT-2 pretokenizer as a local stencil on UTF-8 bytes (device side).
//
// Replaces regex.finditer with
//   '(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+
// (reference: models/tokenizer/train.py:143-146 == models/tokenizer/tokenizer.py:26-27).
// Pretoken starts are a function of <= 4 code points of left context and <= 2 of right context
// (SURVEY Appendix B), so every byte decides independently whether a pretoken starts on it.
//
// Per-byte class byte:  bits 0-2 = class of the code point the byte belongs to
//   (0 S=\s, 1 L=\p{L}, 2 N=\p{N}, 3 P=other, 4 B=boundary: padding / special-token bytes),
//   bit 3 = this byte is the first byte of a code point.
#pragma once
#include "common.cuh"
#include "unicode_tables.h"

#define CLS_S 0u
#define CLS_L 1u
#define CLS_N 2u
#define CLS_P 3u
#define CLS_B 4u
#define CLS_LEAD 8u

#define PT_NT 256                 // threads per CTA
#define PT_CHUNK 16               // bytes per thread
#define PT_TILE (PT_NT * PT_CHUNK)

struct PretokTables {             // shared-memory copy of the Unicode class tables (12.9 KB)
    uint32_t pages[BPE_UC_NPAGES * 16];
    uint8_t index[BPE_UC_NINDEX];
    uint8_t ascii[128];
};

#ifdef __CUDACC__

__device__ __constant__ uint32_t c_uc_pages[BPE_UC_NPAGES * 16];
__device__ __constant__ uint8_t c_uc_index[BPE_UC_NINDEX];
__device__ __constant__ uint8_t c_uc_ascii[128];

__device__ __forceinline__ void pretok_load_tables(PretokTables *t) {
    for (int i = threadIdx.x; i < BPE_UC_NPAGES * 16; i += blockDim.x) t->pages[i] = c_uc_pages[i];
    for (int i = threadIdx.x; i < BPE_UC_NINDEX; i += blockDim.x) t->index[i] = c_uc_index[i];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) t->ascii[i] = c_uc_ascii[i];
}

__device__ __forceinline__ u32 cp_class_smem(const PretokTables *t, u32 cp) {
    if (cp >= 0x110000u) return CLS_P;
    u32 page = t->index[cp >> 8];
    u32 w = t->pages[page * 16 + ((cp & 255u) >> 4)];
    return (w >> (2 * (cp & 15u))) & 3u;
}

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

// byte k (0..47) of a 3-chunk register window {prev, cur, next}; k is a compile-time constant after unrolling
__device__ __forceinline__ u32 win_byte(const u32 (&w)[12], int k) { return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu; }

// Classify the 16 bytes of chunk `cur` (window byte 16..31).  prev/next chunks supply the <=3 bytes of
// look-back / look-ahead a straddling code point needs.  Returns 16 class bytes packed in a uint4.
// err_rel: smallest chunk-relative offset (0..15) of an ill-formed sequence that STARTS in this chunk, or 0xFF.
// has_cr: set when the chunk contains '\r'.
__device__ __forceinline__ uint4 classify_chunk(const PretokTables *tb, uint4 prev, uint4 cur, uint4 next,
                                                u32 *err_rel, bool *has_cr) {
    u32 w[12] = {prev.x, prev.y, prev.z, prev.w, cur.x, cur.y, cur.z, cur.w, next.x, next.y, next.z, next.w};
    u32 out[4] = {0, 0, 0, 0};
    u32 e = 0xFFu;
    bool cr = false;
    int carry_left = 0; u32 carry_cls = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const int k = 16 + j;
        u32 b = win_byte(w, k);
        u32 c;
        if (b < 0x80u) {
            c = tb->ascii[b] | CLS_LEAD;
            cr |= (b == 0x0Du);
            carry_left = 0;
        } else if (b >= 0xF8u) {                 // 0xF8..0xFF never occur in UTF-8: padding / boundary
            c = CLS_B | CLS_LEAD;
            if (e == 0xFFu) e = j;               // callers ignore errors at offsets outside the payload
            carry_left = 0;
        } else if (b >= 0xC0u) {
            u32 cp; bool ok;
            int len = utf8_decode_multi(b, win_byte(w, k + 1), win_byte(w, k + 2), win_byte(w, k + 3), &cp, &ok);
            carry_cls = ok ? cp_class_smem(tb, cp) : CLS_P;
            c = carry_cls | CLS_LEAD;
            carry_left = ok ? len - 1 : 0;
            if (!ok && e == 0xFFu) e = j;
        } else {                                 // continuation byte 0x80..0xBF
            if (carry_left > 0) { c = carry_cls; carry_left--; }
            else {
                // the lead byte sits in the previous chunk (or this byte is an orphan)
                bool found = false; u32 cls2 = CLS_P;
#pragma unroll
                for (int d = 1; d <= 3; d++) {
                    if (!found && d > j) {       // only positions before this chunk can hold an unseen lead
                        u32 l0 = win_byte(w, k - d);
                        if (l0 >= 0xC0u && l0 < 0xF8u) {
                            u32 cp; bool ok;
                            int len = utf8_decode_multi(l0, win_byte(w, k - d + 1), win_byte(w, k - d + 2), win_byte(w, k - d + 3), &cp, &ok);
                            if (ok && len > d) { found = true; cls2 = cp_class_smem(tb, cp); carry_cls = cls2; carry_left = len - 1 - d; }
                        }
                    }
                }
                c = cls2;
                if (!found && e == 0xFFu) e = j;  // orphan continuation byte (or tail of an invalid sequence)
            }
        }
        out[j >> 2] |= c << ((j & 3) * 8);
    }
    *err_rel = e;
    *has_cr = cr;
    return make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ bool is_sdmt(u32 b) { return b == 's' || b == 'd' || b == 'm' || b == 't'; }
__device__ __forceinline__ bool is_llvere(u32 x, u32 y) {
    return (x == 'l' && y == 'l') || (x == 'v' && y == 'e') || (x == 'r' && y == 'e');
}

// Pretoken-start bits for the 16 bytes of chunk `cur`, given text and class windows {prev,cur,next}.
// Bit j set <=> a pretoken starts at chunk byte j.  Implements SURVEY Appendix B.1 on bytes:
//  - whitespace run: start at its first char, and at its last char when the run (len>=2) is followed by \S
//  - L/N/P run: start at its first char unless the previous char is U+0020
//  - contractions 's 'd 'm 't 'll 've 're: letters inside are not starts, the first letter after is
//  - a char after a boundary byte (text start, after a special token) is always a start
__device__ __forceinline__ u32 flags_chunk(uint4 tp, uint4 tc, uint4 tn, uint4 cp_, uint4 cc, uint4 cn) {
    u32 t[12] = {tp.x, tp.y, tp.z, tp.w, tc.x, tc.y, tc.z, tc.w, tn.x, tn.y, tn.z, tn.w};
    u32 c[12] = {cp_.x, cp_.y, cp_.z, cp_.w, cc.x, cc.y, cc.z, cc.w, cn.x, cn.y, cn.z, cn.w};
    u32 bits = 0;
    // class of the first code point that starts after the current byte: walk right-to-left
    u32 next_cls = CLS_B;
#pragma unroll
    for (int k = 36; k >= 32; k--) {             // the next lead is at most 4 bytes into the next chunk
        u32 ck = win_byte(c, k);
        if (ck & CLS_LEAD) next_cls = ck & 7u;
    }
#pragma unroll
    for (int j = 15; j >= 0; j--) {
        const int k = 16 + j;
        u32 ck = win_byte(c, k);
        if (ck & CLS_LEAD) {
            u32 cl = ck & 7u;
            u32 pc = win_byte(c, k - 1) & 7u;
            u32 t0 = win_byte(t, k), t1 = win_byte(t, k + 1);
            u32 tm1 = win_byte(t, k - 1), tm2 = win_byte(t, k - 2), tm3 = win_byte(t, k - 3);
            bool st;
            if (cl == CLS_B) st = false;
            else if (pc == CLS_B) st = true;
            else if (cl == CLS_S) st = (pc != CLS_S) || (next_cls != CLS_S && next_cls != CLS_B);
            else {
                st = (pc != cl) && (tm1 != ' ');
                if (cl == CLS_L) {
                    // ok_mD: an apostrophe at k-D opens a contraction only if the char before it (k-D-1) is
                    // neither class P nor U+0020 (a boundary byte there counts as "text start": ok)
                    u32 c_m2 = win_byte(c, k - 2) & 7u, c_m3 = win_byte(c, k - 3) & 7u, c_m4 = win_byte(c, k - 4) & 7u;
                    u32 tm4 = win_byte(t, k - 4);
                    bool ok_m1 = !(c_m2 == CLS_P || (tm2 == ' ' && c_m2 != CLS_B));
                    bool ok_m2 = !(c_m3 == CLS_P || (tm3 == ' ' && c_m3 != CLS_B));
                    bool ok_m3 = !(c_m4 == CLS_P || (tm4 == ' ' && c_m4 != CLS_B));
                    bool a1 = tm1 == '\'' && pc == CLS_P;          // the byte really is an in-text apostrophe
                    bool a2 = tm2 == '\'' && c_m2 == CLS_P;
                    bool a3 = tm3 == '\'' && c_m3 == CLS_P;
                    if (a2 && ok_m2 && is_sdmt(tm1)) st = true;                        // first letter after 's 'd 'm 't
                    if (a3 && ok_m3 && is_llvere(tm2, tm1)) st = true;                  // first letter after 'll 've 're
                    if (a1 && ok_m1 && (is_sdmt(t0) || is_llvere(t0, t1))) st = false;  // 1st letter inside
                    if (a2 && ok_m2 && is_llvere(tm1, t0)) st = false;                  // 2nd letter inside
                }
            }
            bits |= (st ? 1u : 0u) << j;
            next_cls = cl;
        }
    }
    return bits;
}

#endif  // __CUDACC__
