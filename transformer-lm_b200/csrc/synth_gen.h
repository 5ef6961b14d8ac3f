// synth_gen.h -- deterministic synthetic corpora (bench / test infrastructure).
//
// The reference ships no corpus for its OWT / TinyStories runs (perf/bpe/owt.py:4-8, perf/bpe/tiny.py:4-8 read
// /data/*.txt), so the benchmark inputs are generated.  The text is a pure function of
// (shape, seed, block index): it is produced in independent SYNTH_BLOCK-byte blocks with integer-only
// arithmetic, so the CUDA kernel and the host function emit identical bytes and any prefix whose
// length is a multiple of SYNTH_BLOCK equals the shorter corpus generated on its own.
//
// TinyStories shape: 2^15 lowercase word types, Zipf-like (log-uniform magnitude, biased to the head),
//   short sentences, dialogue in curly and straight quotes, contractions, "\n" paragraphs,
//   documents separated by "\n<|endoftext|>\n".
// OWT shape: 2^21 word types with a long tail, mixed case, numbers, URLs / e-mails / hashtags,
//   ~1.5 % non-ASCII (Latin-1 accents, Cyrillic, CJK, emoji, NBSP / thin space), "\n\n" paragraphs,
//   occasional space / tab runs and trailing spaces, documents separated by "<|endoftext|>".
#pragma once
#include <stdint.h>

#define SYNTH_BLOCK 4096u

#ifdef __CUDACC__
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD static inline
#endif

struct SynthRng { uint64_t s; };
SYNTH_HD uint64_t synth_mix(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
SYNTH_HD uint32_t synth_next(SynthRng &r) { r.s += 0x9E3779B97F4A7C15ull; return (uint32_t)(synth_mix(r.s) >> 32); }
SYNTH_HD uint32_t synth_below(SynthRng &r, uint32_t n) { return (uint32_t)(((uint64_t)synth_next(r) * n) >> 32); }

struct SynthOut { uint8_t *p; uint32_t pos, limit; };
SYNTH_HD void synth_put(SynthOut &o, uint8_t b) { if (o.pos < o.limit) o.p[o.pos++] = b; }
SYNTH_HD void synth_puts(SynthOut &o, const char *s) { while (*s) synth_put(o, (uint8_t)*s++); }
SYNTH_HD void synth_put_cp(SynthOut &o, uint32_t cp) {
    if (cp < 0x80) synth_put(o, (uint8_t)cp);
    else if (cp < 0x800) { synth_put(o, 0xC0 | (cp >> 6)); synth_put(o, 0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { synth_put(o, 0xE0 | (cp >> 12)); synth_put(o, 0x80 | ((cp >> 6) & 0x3F)); synth_put(o, 0x80 | (cp & 0x3F)); }
    else { synth_put(o, 0xF0 | (cp >> 18)); synth_put(o, 0x80 | ((cp >> 12) & 0x3F)); synth_put(o, 0x80 | ((cp >> 6) & 0x3F)); synth_put(o, 0x80 | (cp & 0x3F)); }
}

// word type id: log-uniform magnitude m in [0, bits), then uniform inside [2^m, 2^(m+1))  (~ Zipf s=1)
SYNTH_HD uint32_t synth_word_id(SynthRng &r, uint32_t bits, bool head_bias) {
    uint32_t m = synth_below(r, bits);
    if (head_bias && (synth_next(r) & 3u) == 0) { uint32_t m2 = synth_below(r, bits); if (m2 < m) m = m2; }
    return (1u << m) + (synth_next(r) & ((1u << m) - 1u));
}

// spelling of word type k: consonant/vowel syllables, length grows with log2(k)
SYNTH_HD void synth_spell(SynthOut &o, uint32_t k, uint32_t salt, bool capital, uint32_t accent /* 0 none, 1 latin-1, 2 cyrillic */) {
    const char *cons = "tnshrdlcmwfgypbvkjxqz";     // 21
    const char *vow = "eaoiu";                      // 5
    uint64_t h = synth_mix(((uint64_t)salt << 32) | k);
    uint32_t m = 31 - (uint32_t)
#ifdef __CUDA_ARCH__
        __clz((int)k);
#else
        __builtin_clz(k);
#endif
    uint32_t letters = 1 + m / 3 + (uint32_t)(h & 1);            // 1..9
    h >>= 1;
    bool vowel_first = (h & 7) == 0; h >>= 3;
    for (uint32_t i = 0; i < letters; i++) {
        bool v = ((i & 1u) == 0) == vowel_first;
        uint32_t c;
        if (v) { c = (uint32_t)vow[h % 5]; h /= 5; } else { c = (uint32_t)cons[h % 21]; h /= 21; }
        if (h < 64) h = synth_mix(h + k + i);
        if (accent == 2) { uint32_t cy = 0x0430u + (c - 'a'); if (capital && i == 0) cy -= 0x20u; synth_put_cp(o, cy); continue; }
        if (accent == 1 && v && i == 1) { synth_put_cp(o, c == 'e' ? 0xE9u : c == 'a' ? 0xE0u : c == 'o' ? 0xF6u : c == 'i' ? 0xEFu : 0xFCu); continue; }
        if (capital && i == 0) c -= 32;
        synth_put(o, (uint8_t)c);
    }
}

SYNTH_HD void synth_number(SynthOut &o, SynthRng &r) {
    uint32_t kind = synth_below(r, 4);
    char buf[12]; int n = 0;
    uint32_t v = kind == 0 ? 1900 + synth_below(r, 130) : kind == 1 ? synth_below(r, 100) : kind == 2 ? synth_below(r, 1000) : synth_word_id(r, 20, false);
    do { buf[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (int i = n - 1; i >= 0; i--) {
        synth_put(o, (uint8_t)buf[i]);
        if (kind == 3 && i > 0 && i % 3 == 0) synth_put(o, ',');
    }
    if (kind == 2) { synth_put(o, '.'); synth_put(o, (uint8_t)('0' + synth_below(r, 10))); synth_put(o, (uint8_t)('0' + synth_below(r, 10))); }
}

// One block of `limit` (<= SYNTH_BLOCK) bytes.
SYNTH_HD void synth_block(int shape, uint64_t seed, uint64_t block, uint8_t *dst, uint32_t limit) {
    SynthRng r; r.s = synth_mix(seed ^ synth_mix(block * 2 + (uint64_t)shape));
    SynthOut o; o.p = dst; o.pos = 0; o.limit = limit;
    const bool owt = shape == 1;
    const uint32_t bits = owt ? 21 : 15;
    const uint32_t salt = owt ? 0x4f57u : 0x5453u;
    // documents: separators at a quarter of the block starts and after ~1/18 (tiny) or ~1/90 (owt) of the sentences
    if (synth_below(r, 4u) == 0) {
        if (owt) synth_puts(o, "<|endoftext|>"); else synth_puts(o, "<|endoftext|>\n");
    }
    while (o.pos + 80 < o.limit) {
        // ---- one sentence ----
        uint32_t n_words = 6 + synth_below(r, 11);
        bool quoted = synth_below(r, owt ? 40u : 12u) == 0;
        bool curly = !owt && (synth_next(r) & 1u);
        if (quoted) { if (curly) synth_put_cp(o, 0x201C); else synth_put(o, '"'); }
        for (uint32_t w = 0; w < n_words && o.pos + 80 < o.limit; w++) {
            if (w) {
                uint32_t sp = synth_below(r, 1000);
                if (owt && sp < 3) synth_puts(o, "  ");
                else if (owt && sp < 4) synth_put(o, '\t');
                else if (owt && sp < 5) synth_put_cp(o, 0x00A0);
                else if (owt && sp < 6) synth_put_cp(o, 0x2009);
                else synth_put(o, ' ');
            }
            uint32_t kind = synth_below(r, 1000);
            bool cap = w == 0 || (owt && synth_below(r, 12) == 0);
            if (owt && kind < 30) synth_number(o, r);
            else if (owt && kind < 33) {                     // URL
                synth_puts(o, (synth_next(r) & 1u) ? "https://" : "http://www.");
                synth_spell(o, synth_word_id(r, bits, false), salt, false, 0);
                synth_puts(o, (synth_next(r) & 1u) ? ".com/" : ".org/");
                synth_spell(o, synth_word_id(r, bits, false), salt, false, 0);
                if (synth_next(r) & 1u) { synth_puts(o, "?id="); synth_number(o, r); }
            } else if (owt && kind < 35) {                   // e-mail / hashtag / handle
                uint32_t t = synth_below(r, 3);
                if (t == 0) { synth_spell(o, synth_word_id(r, bits, false), salt, false, 0); synth_put(o, '@'); synth_spell(o, synth_word_id(r, 12, false), salt, false, 0); synth_puts(o, ".com"); }
                else { synth_put(o, t == 1 ? '#' : '@'); synth_spell(o, synth_word_id(r, bits, false), salt, cap, 0); }
            } else if (owt && kind < 41) synth_spell(o, synth_word_id(r, bits, false), salt, cap, 1);      // Latin-1 accents
            else if (owt && kind < 45) synth_spell(o, synth_word_id(r, 14, false), salt, false, 2);        // Cyrillic
            else if (owt && kind < 48) {                     // CJK run
                uint32_t len = 1 + synth_below(r, 4);
                for (uint32_t i = 0; i < len; i++) synth_put_cp(o, 0x4E00 + synth_word_id(r, 11, false));
            } else if (kind < (owt ? 50u : 3u)) synth_put_cp(o, 0x1F600 + synth_below(r, 64));             // emoji
            else {
                synth_spell(o, synth_word_id(r, bits, !owt), salt, cap, 0);
                uint32_t c = synth_below(r, 350);
                if (c < 7) { const char *sfx[7] = {"'s", "'t", "'ll", "'re", "'ve", "'d", "'m"}; synth_puts(o, sfx[c]); }
                else if (owt && c < 9) { synth_put(o, '-'); synth_spell(o, synth_word_id(r, bits, false), salt, false, 0); }
            }
            if (w + 1 < n_words && synth_below(r, 12) == 0) synth_put(o, ',');
            if (owt && w + 1 < n_words && synth_below(r, 150) == 0) { synth_puts(o, " ("); synth_spell(o, synth_word_id(r, bits, false), salt, false, 0); synth_put(o, ')'); }
        }
        uint32_t e = synth_below(r, 20);
        synth_put(o, e < 15 ? '.' : e < 18 ? '!' : '?');
        if (owt && e == 19) synth_puts(o, "..");
        if (quoted) { if (curly) synth_put_cp(o, 0x201D); else synth_put(o, '"'); }
        uint32_t nl = synth_below(r, owt ? 7u : 5u);
        if (nl == 0) { if (owt && synth_below(r, 20) == 0) synth_put(o, ' '); synth_puts(o, owt ? "\n\n" : "\n"); }
        else synth_put(o, ' ');
        if (synth_below(r, owt ? 90u : 18u) == 0 && o.pos + 100 < o.limit) {
            if (owt) synth_puts(o, "<|endoftext|>"); else synth_puts(o, "\n<|endoftext|>\n");
        }
    }
    // fill the tail of the block with one filler word and a newline so the block ends exactly at `limit`
    while (o.pos + 1 < o.limit) synth_put(o, 'x');
    synth_put(o, '\n');
}
