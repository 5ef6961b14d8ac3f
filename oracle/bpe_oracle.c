/*
 * bpe_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the byte-level BPE path of gashon/transformer-lm
 * (pure Python in the reference).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (transformer-lm_b200/) never links, imports or calls it.
 *
 * Parity pin: tests/test_oracle_pins.py checks this file against
 *   - the reference's own golden vectors (tests/fixtures/train-bpe-reference-*.{txt,json},
 *     tiktoken-GPT-2 ids for the reference's test strings/fixtures), and
 *   - outputs of the reference itself (imported from /root/reference in the build
 *     container; the generated vectors are committed under tests/golden/ by
 *     tools/make_golden.py), and
 *   - the `regex` module on fuzzed strings for the pretokenizer.
 *
 * Every function cites the reference file:line it restates
 * (paths relative to /root/reference).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../transformer-lm_b200/csrc/unicode_tables.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* small utilities                                                            */
/* ------------------------------------------------------------------------- */

static void *xmalloc(size_t n) {
    void *p = malloc(n ? n : 1);
    if (!p) { fprintf(stderr, "bpe_oracle: out of memory (%zu)\n", n); abort(); }
    return p;
}
static void *xrealloc(void *q, size_t n) {
    void *p = realloc(q, n ? n : 1);
    if (!p) { fprintf(stderr, "bpe_oracle: out of memory (%zu)\n", n); abort(); }
    return p;
}

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}
static uint64_t hash_bytes(const uint8_t *p, size_t n, uint64_t seed) {
    uint64_t h = 0xcbf29ce484222325ULL ^ seed;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return mix64(h ^ n);
}

/* class of a code point: 0=S 1=L 2=N 3=P (tables generated from `regex`) */
static inline int cp_class(uint32_t cp) {
    if (cp < 128) return bpe_uc_ascii[cp];
    if (cp >= 0x110000) return BPE_CLS_P;
    uint32_t page = bpe_uc_index[cp >> 8];
    uint32_t w = bpe_uc_pages[page * 16 + ((cp & 255) >> 4)];
    return (w >> (2 * (cp & 15))) & 3;
}

ORC_API int orc_cp_class(uint32_t cp) { return cp_class(cp); }

/* ------------------------------------------------------------------------- */
/* UTF-8 strict validation / decoding                                         */
/* (open(path, "r", encoding="utf-8").read() -- models/tokenizer/train.py:21-23) */
/* ------------------------------------------------------------------------- */

/* Decode one code point at s[i]; returns its length (1-4) or 0 if invalid. */
static inline int utf8_decode1(const uint8_t *s, size_t n, size_t i, uint32_t *cp) {
    uint8_t b0 = s[i];
    if (b0 < 0x80) { *cp = b0; return 1; }
    if (b0 < 0xC2) return 0;
    if (b0 < 0xE0) {
        if (i + 1 >= n || (s[i + 1] & 0xC0) != 0x80) return 0;
        *cp = ((uint32_t)(b0 & 0x1F) << 6) | (s[i + 1] & 0x3F);
        return 2;
    }
    if (b0 < 0xF0) {
        if (i + 2 >= n) return 0;
        uint8_t b1 = s[i + 1], b2 = s[i + 2];
        if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80) return 0;
        if (b0 == 0xE0 && b1 < 0xA0) return 0;          /* overlong */
        if (b0 == 0xED && b1 >= 0xA0) return 0;         /* surrogates */
        *cp = ((uint32_t)(b0 & 0x0F) << 12) | ((uint32_t)(b1 & 0x3F) << 6) | (b2 & 0x3F);
        return 3;
    }
    if (b0 < 0xF5) {
        if (i + 3 >= n) return 0;
        uint8_t b1 = s[i + 1], b2 = s[i + 2], b3 = s[i + 3];
        if ((b1 & 0xC0) != 0x80 || (b2 & 0xC0) != 0x80 || (b3 & 0xC0) != 0x80) return 0;
        if (b0 == 0xF0 && b1 < 0x90) return 0;          /* overlong */
        if (b0 == 0xF4 && b1 >= 0x90) return 0;         /* > U+10FFFF */
        *cp = ((uint32_t)(b0 & 0x07) << 18) | ((uint32_t)(b1 & 0x3F) << 12) |
              ((uint32_t)(b2 & 0x3F) << 6) | (b3 & 0x3F);
        return 4;
    }
    return 0;
}

/* Returns -1 when `s` is valid UTF-8, else the byte offset of the first
 * ill-formed sequence (== UnicodeDecodeError.start in CPython). */
ORC_API int64_t orc_utf8_validate(const uint8_t *s, uint64_t n) {
    size_t i = 0;
    while (i < n) {
        uint32_t cp;
        int l = utf8_decode1(s, n, i, &cp);
        if (!l) return (int64_t)i;
        i += l;
    }
    return -1;
}

/* Universal-newline translation of text-mode open() (train.py:21-23; SURVEY A-2):
 * "\r\n" -> "\n", lone "\r" -> "\n".  Returns the new length (dst may alias src). */
ORC_API uint64_t orc_universal_newlines(const uint8_t *src, uint64_t n, uint8_t *dst) {
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; i++) {
        uint8_t c = src[i];
        if (c == '\r') {
            dst[o++] = '\n';
            if (i + 1 < n && src[i + 1] == '\n') i++;
        } else dst[o++] = c;
    }
    return o;
}

/* ------------------------------------------------------------------------- */
/* GPT-2 pretokenizer as a literal ordered-alternation matcher               */
/*   '(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+ */
/* (train.py:143-146 == tokenizer.py:26-27; finditer semantics: leftmost,    */
/*  alternatives tried in order, greedy with backtracking)                   */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint32_t *cp;     /* code points */
    uint8_t  *cls;    /* class per code point */
    uint64_t *off;    /* byte offset per code point, off[n] = n_bytes */
    size_t    n;
} cptext;

static int cptext_decode(const uint8_t *s, size_t nbytes, cptext *t) {
    t->cp = xmalloc(sizeof(uint32_t) * (nbytes + 1));
    t->cls = xmalloc(nbytes + 1);
    t->off = xmalloc(sizeof(uint64_t) * (nbytes + 1));
    size_t i = 0, k = 0;
    while (i < nbytes) {
        uint32_t cp;
        int l = utf8_decode1(s, nbytes, i, &cp);
        if (!l) { free(t->cp); free(t->cls); free(t->off); return -1; }
        t->cp[k] = cp; t->cls[k] = (uint8_t)cp_class(cp); t->off[k] = i;
        k++; i += l;
    }
    t->off[k] = nbytes; t->n = k;
    return 0;
}
static void cptext_free(cptext *t) { free(t->cp); free(t->cls); free(t->off); }

/* Length (in code points) of the regex match starting at code point i (always >= 1). */
static size_t gpt2_match_len(const cptext *t, size_t i) {
    const uint32_t *c = t->cp; const uint8_t *k = t->cls; size_t n = t->n;
    /* alt 1: '(?:[sdmt]|ll|ve|re) */
    if (c[i] == '\'' && i + 1 < n) {
        uint32_t x = c[i + 1];
        if (x == 's' || x == 'd' || x == 'm' || x == 't') return 2;
        if (i + 2 < n) {
            uint32_t y = c[i + 2];
            if ((x == 'l' && y == 'l') || (x == 'v' && y == 'e') || (x == 'r' && y == 'e')) return 3;
        }
    }
    /* alts 2-4: " ?" + run of one class among L, N, P */
    for (int want = BPE_CLS_L; want <= BPE_CLS_P; want++) {
        size_t j = i;
        if (c[j] == ' ' && j + 1 < n && k[j + 1] == want) j++;   /* optional space is greedy;
                                         backtracking to "no space" is tried next */
        if (k[j] == want) {
            size_t e = j;
            while (e < n && k[e] == want) e++;
            return e - i;
        }
        /* " ?" matched empty: class run must start at i itself (handled above when c[i]!=' ';
           when c[i]==' ' its class is S, never L/N/P, so the alternative fails) */
    }
    /* alt 5: \s+(?!\S) with backtracking, alt 6: \s+ */
    if (k[i] == BPE_CLS_S) {
        size_t e = i;
        while (e < n && k[e] == BPE_CLS_S) e++;
        if (e == n) return e - i;                /* (?!\S) holds at end of text */
        if (e - i >= 2) return e - i - 1;        /* give back one: next char is \s */
        return 1;                                /* alt 5 fails, alt 6 takes the single \s */
    }
    return 1; /* unreachable for valid input: every code point is S, L, N or P */
}

/* Pretokenize text[0..n): writes the BYTE offset of every pretoken start into
 * starts[] (up to cap) and returns the number of pretokens, or -1 on invalid UTF-8. */
ORC_API int64_t orc_pretokenize(const uint8_t *text, uint64_t n, uint64_t *starts, uint64_t cap) {
    cptext t;
    if (cptext_decode(text, n, &t)) return -1;
    size_t i = 0; int64_t cnt = 0;
    while (i < t.n) {
        if ((uint64_t)cnt < cap) starts[cnt] = t.off[i];
        cnt++;
        i += gpt2_match_len(&t, i);
    }
    cptext_free(&t);
    return cnt;
}

/* ------------------------------------------------------------------------- */
/* byte-string -> value hash map (python dict[str|bytes, int])               */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint64_t *hash;       /* 0 = empty */
    uint64_t *koff;       /* key offset into pool */
    uint32_t *klen;
    int64_t  *val;
    uint32_t *extra;      /* second discriminator (e.g. split position), part of the key */
    size_t cap, cnt;
    uint8_t *pool; size_t pool_n, pool_cap;
    uint32_t *order;      /* slots in insertion order */
} bmap;

static void bmap_init(bmap *m, size_t cap_pow2) {
    m->cap = cap_pow2; m->cnt = 0;
    m->hash = calloc(m->cap, sizeof(uint64_t));
    m->koff = xmalloc(m->cap * sizeof(uint64_t));
    m->klen = xmalloc(m->cap * sizeof(uint32_t));
    m->val = xmalloc(m->cap * sizeof(int64_t));
    m->extra = xmalloc(m->cap * sizeof(uint32_t));
    m->order = xmalloc(m->cap * sizeof(uint32_t));
    m->pool_cap = 1 << 16; m->pool_n = 0; m->pool = xmalloc(m->pool_cap);
    if (!m->hash) abort();
}
static void bmap_free(bmap *m) {
    free(m->hash); free(m->koff); free(m->klen); free(m->val); free(m->extra); free(m->order); free(m->pool);
}
static void bmap_grow(bmap *m);
/* find-or-insert; returns slot; *is_new set */
static size_t bmap_get(bmap *m, const uint8_t *k, uint32_t len, uint32_t extra, int insert, int *is_new) {
    uint64_t h = hash_bytes(k, len, extra) | 1;
    size_t s = h & (m->cap - 1);
    for (;;) {
        if (m->hash[s] == 0) {
            if (is_new) *is_new = 1;
            if (!insert) return (size_t)-1;
            if ((m->cnt + 1) * 2 > m->cap) { bmap_grow(m); return bmap_get(m, k, len, extra, insert, is_new); }
            if (m->pool_n + len > m->pool_cap) {
                while (m->pool_n + len > m->pool_cap) m->pool_cap *= 2;
                m->pool = xrealloc(m->pool, m->pool_cap);
            }
            memcpy(m->pool + m->pool_n, k, len);
            m->hash[s] = h; m->koff[s] = m->pool_n; m->klen[s] = len; m->extra[s] = extra; m->val[s] = 0;
            m->pool_n += len;
            m->order[m->cnt++] = (uint32_t)s;
            return s;
        }
        if (m->hash[s] == h && m->klen[s] == len && m->extra[s] == extra &&
            memcmp(m->pool + m->koff[s], k, len) == 0) {
            if (is_new) *is_new = 0;
            return s;
        }
        s = (s + 1) & (m->cap - 1);
    }
}
static void bmap_grow(bmap *m) {
    bmap o = *m;
    m->cap = o.cap * 2;
    m->hash = calloc(m->cap, sizeof(uint64_t));
    m->koff = xmalloc(m->cap * sizeof(uint64_t));
    m->klen = xmalloc(m->cap * sizeof(uint32_t));
    m->val = xmalloc(m->cap * sizeof(int64_t));
    m->extra = xmalloc(m->cap * sizeof(uint32_t));
    m->order = xmalloc(m->cap * sizeof(uint32_t));
    if (!m->hash) abort();
    for (size_t i = 0; i < o.cnt; i++) {         /* re-insert in insertion order */
        size_t so = o.order[i];
        uint64_t h = o.hash[so];
        size_t s = h & (m->cap - 1);
        while (m->hash[s]) s = (s + 1) & (m->cap - 1);
        m->hash[s] = h; m->koff[s] = o.koff[so]; m->klen[s] = o.klen[so];
        m->extra[s] = o.extra[so]; m->val[s] = o.val[so];
        m->order[i] = (uint32_t)s;
    }
    free(o.hash); free(o.koff); free(o.klen); free(o.val); free(o.extra); free(o.order);
}

/* ------------------------------------------------------------------------- */
/* extract_subword_frequencies  (train.py:16-28)                              */
/* ------------------------------------------------------------------------- */

static int is_special(const uint8_t *p, size_t len, const uint8_t *blob, const uint32_t *offs, int ns) {
    for (int i = 0; i < ns; i++) {
        size_t l = offs[i + 1] - offs[i];
        if (l == len && memcmp(blob + offs[i], p, len) == 0) return 1;
    }
    return 0;
}

/* Count unique pretokens of `text` (already newline-translated) into `m`;
 * pretokens equal to a special token are skipped (train.py:25). */
static int count_pretokens(const uint8_t *text, size_t n, const uint8_t *sp_blob,
                           const uint32_t *sp_offs, int n_sp, bmap *m) {
    cptext t;
    if (cptext_decode(text, n, &t)) return -1;
    size_t i = 0;
    while (i < t.n) {
        size_t l = gpt2_match_len(&t, i);
        const uint8_t *p = text + t.off[i];
        size_t blen = t.off[i + l] - t.off[i];
        if (!is_special(p, blen, sp_blob, sp_offs, n_sp)) {
            size_t s = bmap_get(m, p, (uint32_t)blen, 0, 1, NULL);
            m->val[s] += 1;
        }
        i += l;
    }
    cptext_free(&t);
    return 0;
}

/* Exposed for parity tests of the count stage: returns number of unique pretokens and
 * fills (if non-NULL, sized by a first call) blob/offs/counts in first-occurrence order. */
ORC_API int64_t orc_count_pretokens(const uint8_t *text, uint64_t n,
                                    const uint8_t *sp_blob, const uint32_t *sp_offs, int n_sp,
                                    uint8_t *out_blob, uint64_t *out_offs, int64_t *out_counts,
                                    uint64_t *blob_bytes) {
    bmap m; bmap_init(&m, 1 << 12);
    if (count_pretokens(text, n, sp_blob, sp_offs, n_sp, &m)) { bmap_free(&m); return -1; }
    if (blob_bytes) *blob_bytes = m.pool_n;
    if (out_blob && out_offs && out_counts) {
        uint64_t o = 0;
        for (size_t i = 0; i < m.cnt; i++) {
            size_t s = m.order[i];
            memcpy(out_blob + o, m.pool + m.koff[s], m.klen[s]);
            out_offs[i] = o; o += m.klen[s];
            out_counts[i] = m.val[s];
        }
        out_offs[m.cnt] = o;
    }
    int64_t r = (int64_t)m.cnt;
    bmap_free(&m);
    return r;
}

/* ------------------------------------------------------------------------- */
/* train_bpe  (train.py:142-231)                                              */
/* ------------------------------------------------------------------------- */

typedef struct { int32_t *v; size_t n, cap; } ivec;
static void ivec_push(ivec *a, int32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 4; a->v = xrealloc(a->v, a->cap * sizeof(int32_t)); }
    a->v[a->n++] = x;
}

typedef struct {
    int32_t a, b;
    int64_t count;        /* byte_pair_frequencies[(a,b)] */
    int alive;            /* key currently present in byte_pair_frequencies */
    ivec words;           /* token_indices[(a,b)] (word indices; duplicates harmless, see below) */
} pair_ent;

typedef struct {
    /* tokens (symbols): 0..255 single bytes, 256+k product of merge k */
    uint8_t **tok; uint32_t *tok_len; size_t n_tok, tok_cap;
    bmap tok_by_bytes;                 /* bytes -> symbol id (token identity is the byte string, SURVEY A-6) */
    /* pair dict */
    pair_ent *pairs; size_t n_pairs, pairs_cap;
    uint64_t *pm_key; int32_t *pm_val; size_t pm_cap, pm_cnt;   /* (a,b) -> index in pairs */
} trainer;

static void pm_grow(trainer *T);
static int32_t pm_find(trainer *T, int32_t a, int32_t b, int insert) {
    uint64_t key = ((uint64_t)(uint32_t)a << 32) | (uint32_t)b;
    size_t s = mix64(key) & (T->pm_cap - 1);
    for (;;) {
        if (T->pm_val[s] < 0) {
            if (!insert) return -1;
            if ((T->pm_cnt + 1) * 2 > T->pm_cap) { pm_grow(T); return pm_find(T, a, b, insert); }
            if (T->n_pairs == T->pairs_cap) {
                T->pairs_cap *= 2; T->pairs = xrealloc(T->pairs, T->pairs_cap * sizeof(pair_ent));
            }
            pair_ent *p = &T->pairs[T->n_pairs];
            memset(p, 0, sizeof(*p)); p->a = a; p->b = b;
            T->pm_key[s] = key; T->pm_val[s] = (int32_t)T->n_pairs; T->pm_cnt++;
            return (int32_t)T->n_pairs++;
        }
        if (T->pm_key[s] == key) return T->pm_val[s];
        s = (s + 1) & (T->pm_cap - 1);
    }
}
static void pm_grow(trainer *T) {
    size_t oc = T->pm_cap; uint64_t *ok = T->pm_key; int32_t *ov = T->pm_val;
    T->pm_cap = oc * 2;
    T->pm_key = xmalloc(T->pm_cap * sizeof(uint64_t));
    T->pm_val = xmalloc(T->pm_cap * sizeof(int32_t));
    for (size_t i = 0; i < T->pm_cap; i++) T->pm_val[i] = -1;
    for (size_t i = 0; i < oc; i++) if (ov[i] >= 0) {
        size_t s = mix64(ok[i]) & (T->pm_cap - 1);
        while (T->pm_val[s] >= 0) s = (s + 1) & (T->pm_cap - 1);
        T->pm_key[s] = ok[i]; T->pm_val[s] = ov[i];
    }
    free(ok); free(ov);
}

/* frequencies[(a,b)] += delta with defaultdict(int) semantics: touching a missing key creates it
 * (train.py:36, 65-78). */
static pair_ent *freq_touch(trainer *T, int32_t a, int32_t b) {
    int32_t i = pm_find(T, a, b, 1);
    pair_ent *p = &T->pairs[i];
    if (!p->alive) { p->alive = 1; p->count = 0; p->words.n = 0; }
    return p;
}

/* python: (bytes_a1, bytes_b1) < (bytes_a2, bytes_b2) tuple compare (train.py:187-189, SURVEY A-10) */
static int bytes_cmp(const uint8_t *x, uint32_t nx, const uint8_t *y, uint32_t ny) {
    uint32_t m = nx < ny ? nx : ny;
    int c = memcmp(x, y, m);
    if (c) return c;
    return (nx > ny) - (nx < ny);
}
static int pair_key_greater(const trainer *T, const pair_ent *p, const pair_ent *q) {
    if (p->count != q->count) return p->count > q->count;
    int c = bytes_cmp(T->tok[p->a], T->tok_len[p->a], T->tok[q->a], T->tok_len[q->a]);
    if (c) return c > 0;
    c = bytes_cmp(T->tok[p->b], T->tok_len[p->b], T->tok[q->b], T->tok_len[q->b]);
    return c > 0;
}

static int32_t add_token(trainer *T, int32_t a, int32_t b) {
    uint32_t la = T->tok_len[a], lb = T->tok_len[b];
    if (T->n_tok == T->tok_cap) {
        T->tok_cap *= 2;
        T->tok = xrealloc(T->tok, T->tok_cap * sizeof(uint8_t *));
        T->tok_len = xrealloc(T->tok_len, T->tok_cap * sizeof(uint32_t));
    }
    uint8_t *nb = xmalloc(la + lb);
    memcpy(nb, T->tok[a], la); memcpy(nb + la, T->tok[b], lb);
    int32_t id = (int32_t)T->n_tok;
    T->tok[id] = nb; T->tok_len[id] = la + lb; T->n_tok++;
    /* identity by bytes: reuse the first symbol with these bytes if one exists */
    int is_new;
    size_t s = bmap_get(&T->tok_by_bytes, nb, la + lb, 0, 1, &is_new);
    if (is_new) T->tok_by_bytes.val[s] = id;
    return (int32_t)T->tok_by_bytes.val[s];
}

/*
 * Train.  text = raw file bytes (newline translation and strict UTF-8 validation happen here,
 * like the text-mode read at train.py:21-23).
 * Outputs: merge_pairs[2*k], merge_pairs[2*k+1] = symbol ids of merge k (0..255 = bytes,
 * 256+j = product of merge j, canonicalised by bytes); *n_done = number of merges performed.
 * Returns 0, or -1 on invalid UTF-8 (*err_off = offset in the ORIGINAL bytes).
 */
ORC_API int orc_train_bpe(const uint8_t *text_in, uint64_t n_in,
                          const uint8_t *sp_blob, const uint32_t *sp_offs, int n_sp,
                          int n_merges, int32_t *merge_pairs, int *n_done, int64_t *err_off) {
    *n_done = 0;
    int64_t bad = orc_utf8_validate(text_in, n_in);
    if (bad >= 0) { if (err_off) *err_off = bad; return -1; }
    uint8_t *text = xmalloc(n_in);
    uint64_t n = orc_universal_newlines(text_in, n_in, text);

    /* extract_subword_frequencies (train.py:16-28) */
    bmap words; bmap_init(&words, 1 << 12);
    count_pretokens(text, n, sp_blob, sp_offs, n_sp, &words);
    free(text);

    trainer T; memset(&T, 0, sizeof(T));
    T.tok_cap = 512; T.tok = xmalloc(T.tok_cap * sizeof(uint8_t *)); T.tok_len = xmalloc(T.tok_cap * sizeof(uint32_t));
    bmap_init(&T.tok_by_bytes, 1 << 12);
    for (int i = 0; i < 256; i++) {
        T.tok[i] = xmalloc(1); T.tok[i][0] = (uint8_t)i; T.tok_len[i] = 1;
        size_t s = bmap_get(&T.tok_by_bytes, T.tok[i], 1, 0, 1, NULL); T.tok_by_bytes.val[s] = i;
    }
    T.n_tok = 256;
    T.pairs_cap = 1024; T.pairs = xmalloc(T.pairs_cap * sizeof(pair_ent));
    T.pm_cap = 1 << 12; T.pm_key = xmalloc(T.pm_cap * sizeof(uint64_t)); T.pm_val = xmalloc(T.pm_cap * sizeof(int32_t));
    for (size_t i = 0; i < T.pm_cap; i++) T.pm_val[i] = -1;

    /* encode_subwords (train.py:31-32): word -> list of single-byte symbols */
    size_t W = words.cnt;
    int32_t **w = xmalloc(W * sizeof(int32_t *));
    uint32_t *wl = xmalloc(W * sizeof(uint32_t));
    int64_t *wf = xmalloc(W * sizeof(int64_t));
    for (size_t i = 0; i < W; i++) {
        size_t s = words.order[i];
        wl[i] = words.klen[s]; wf[i] = words.val[s];
        w[i] = xmalloc(sizeof(int32_t) * (wl[i] ? wl[i] : 1));
        for (uint32_t j = 0; j < wl[i]; j++) w[i][j] = words.pool[words.koff[s] + j];
    }
    /* calculate_byte_pair_frequencies (train.py:35-49) */
    for (size_t i = 0; i < W; i++)
        for (uint32_t j = 0; j + 1 < wl[i]; j++) {
            pair_ent *p = freq_touch(&T, w[i][j], w[i][j + 1]);
            p->count += wf[i];
            if (p->words.n == 0 || p->words.v[p->words.n - 1] != (int32_t)i) ivec_push(&p->words, (int32_t)i);
        }

    /* merge loop (train.py:183-228) */
    for (int step = 0; step < n_merges; step++) {
        /* len(byte_pair_frequencies) == 0 -> break (184-185); max over ALL keys incl. count 0 (187-189) */
        int32_t best = -1;
        for (size_t i = 0; i < T.n_pairs; i++) {
            if (!T.pairs[i].alive) continue;
            if (best < 0 || pair_key_greater(&T, &T.pairs[i], &T.pairs[best])) best = (int32_t)i;
        }
        if (best < 0) break;
        int32_t a = T.pairs[best].a, b = T.pairs[best].b;
        int32_t nw = add_token(&T, a, b);                 /* new_byte = a + b (190) */
        /* subword_indices = list(token_indices[best_pair].keys()) (192): snapshot */
        size_t nidx = T.pairs[best].words.n;
        int32_t *idx = xmalloc(sizeof(int32_t) * (nidx ? nidx : 1));
        memcpy(idx, T.pairs[best].words.v, sizeof(int32_t) * nidx);
        for (size_t q = 0; q < nidx; q++) {
            int32_t wi = idx[q];
            int32_t *sw = w[wi]; int64_t c = wf[wi];
            uint32_t bi = 0;
            while (wl[wi] >= 2 && bi < wl[wi] - 1) {      /* 196-224 */
                if (sw[bi] == a && sw[bi + 1] == b) {
                    /* update_frequencies_after_merge (52-78) on the not-yet-spliced word */
                    if (bi > 0) {
                        freq_touch(&T, sw[bi - 1], sw[bi])->count -= c;
                        freq_touch(&T, sw[bi - 1], nw)->count += c;
                    }
                    if (bi + 2 < wl[wi]) {
                        freq_touch(&T, sw[bi + 1], sw[bi + 2])->count -= c;
                        freq_touch(&T, nw, sw[bi + 2])->count += c;
                    }
                    /* update_token_indices (81-104) never deletes anything (SURVEY A-9) */
                    /* merge_subwords (132-139) */
                    sw[bi] = nw;
                    memmove(sw + bi + 1, sw + bi + 2, sizeof(int32_t) * (wl[wi] - bi - 2));
                    wl[wi]--;
                    /* create_new_token_indices (107-129) */
                    if (bi > 0) ivec_push(&freq_touch(&T, sw[bi - 1], sw[bi])->words, wi);
                    if (bi + 1 < wl[wi]) ivec_push(&freq_touch(&T, sw[bi], sw[bi + 1])->words, wi);
                }
                bi++;
            }
        }
        free(idx);
        /* pop best from both dicts (226-227), merges.append (228) */
        pair_ent *bp = &T.pairs[pm_find(&T, a, b, 0)];
        bp->alive = 0; bp->words.n = 0;
        merge_pairs[2 * step] = a; merge_pairs[2 * step + 1] = b;
        *n_done = step + 1;
    }

    for (size_t i = 0; i < W; i++) free(w[i]);
    free(w); free(wl); free(wf);
    for (size_t i = 0; i < T.n_pairs; i++) free(T.pairs[i].words.v);
    free(T.pairs); free(T.pm_key); free(T.pm_val);
    for (size_t i = 0; i < T.n_tok; i++) free(T.tok[i]);
    free(T.tok); free(T.tok_len); bmap_free(&T.tok_by_bytes); bmap_free(&words);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Tokenizer  (tokenizer.py:12-157)                                           */
/* ------------------------------------------------------------------------- */

typedef struct {
    bmap vocab_inv;       /* bytes -> id                      (tokenizer.py:19) */
    bmap merges;          /* (a+b bytes, len(a)) -> rank, last duplicate wins (tokenizer.py:115) */
    uint8_t *sp_blob; uint32_t *sp_offs; int n_sp;  /* specials, longest first (tokenizer.py:29-30) */
    uint8_t **vocab; uint32_t *vocab_len; int64_t vocab_n;  /* id -> bytes, NULL if absent */
} orc_tok;

ORC_API orc_tok *orc_tok_create(const uint8_t *vocab_blob, const uint64_t *vocab_offs, const int64_t *vocab_ids, int64_t n_vocab,
                                const uint8_t *merge_blob, const uint64_t *merge_offs /* 2*n_merges+1 */, int64_t n_merges,
                                const uint8_t *sp_blob, const uint32_t *sp_offs, int n_sp) {
    orc_tok *t = xmalloc(sizeof(orc_tok)); memset(t, 0, sizeof(*t));
    bmap_init(&t->vocab_inv, 1 << 12); bmap_init(&t->merges, 1 << 12);
    int64_t max_id = -1;
    for (int64_t i = 0; i < n_vocab; i++) if (vocab_ids[i] > max_id) max_id = vocab_ids[i];
    t->vocab_n = max_id + 1;
    t->vocab = calloc((size_t)(t->vocab_n ? t->vocab_n : 1), sizeof(uint8_t *));
    t->vocab_len = calloc((size_t)(t->vocab_n ? t->vocab_n : 1), sizeof(uint32_t));
    for (int64_t i = 0; i < n_vocab; i++) {            /* {v: k for k, v in vocab.items()}: last wins */
        const uint8_t *p = vocab_blob + vocab_offs[i]; uint32_t l = (uint32_t)(vocab_offs[i + 1] - vocab_offs[i]);
        size_t s = bmap_get(&t->vocab_inv, p, l, 0, 1, NULL);
        t->vocab_inv.val[s] = vocab_ids[i];
        int64_t id = vocab_ids[i];
        if (id >= 0) { t->vocab[id] = xmalloc(l ? l : 1); memcpy(t->vocab[id], p, l); t->vocab_len[id] = l; }
    }
    for (int64_t i = 0; i < n_merges; i++) {
        const uint8_t *pa = merge_blob + merge_offs[2 * i];
        uint32_t la = (uint32_t)(merge_offs[2 * i + 1] - merge_offs[2 * i]);
        uint32_t lb = (uint32_t)(merge_offs[2 * i + 2] - merge_offs[2 * i + 1]);
        size_t s = bmap_get(&t->merges, pa, la + lb, la, 1, NULL);   /* blob stores a then b contiguously */
        t->merges.val[s] = i;
    }
    t->n_sp = n_sp;
    t->sp_offs = xmalloc(sizeof(uint32_t) * (n_sp + 1));
    memcpy(t->sp_offs, sp_offs, sizeof(uint32_t) * (n_sp + 1));
    t->sp_blob = xmalloc(sp_offs[n_sp] ? sp_offs[n_sp] : 1);
    memcpy(t->sp_blob, sp_blob, sp_offs[n_sp]);
    return t;
}
ORC_API void orc_tok_destroy(orc_tok *t) {
    if (!t) return;
    bmap_free(&t->vocab_inv); bmap_free(&t->merges);
    for (int64_t i = 0; i < t->vocab_n; i++) free(t->vocab[i]);
    free(t->vocab); free(t->vocab_len); free(t->sp_blob); free(t->sp_offs); free(t);
}

typedef struct { int64_t *v; size_t n, cap; } lvec;
static void lvec_push(lvec *a, int64_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = xrealloc(a->v, a->cap * sizeof(int64_t)); }
    a->v[a->n++] = x;
}

/* BPE of one pretoken (tokenizer.py:124-136).  Returns 0, or -2 + sets (*kerr_p,*kerr_len) when a
 * final byte string is missing from vocab_inv (KeyError at tokenizer.py:135). */
static int encode_pretoken(const orc_tok *t, const uint8_t *p, uint32_t len, lvec *out,
                           const uint8_t **kerr_p, uint32_t *kerr_len) {
    /* tokens are contiguous slices: cut[i]..cut[i+1] */
    uint32_t stackcut[257]; uint32_t *cut = stackcut;
    if (len + 1 > 257) cut = xmalloc(sizeof(uint32_t) * (len + 1));
    uint32_t nt = len;
    for (uint32_t i = 0; i <= len; i++) cut[i] = i;
    while (nt > 1) {
        /* pair = min(pairs, key=rank); first minimal pair in order (all equal-rank pairs are the same pair) */
        int64_t best_rank = -1; uint32_t best_i = 0;
        for (uint32_t i = 0; i + 1 < nt; i++) {
            size_t s = bmap_get((bmap *)&t->merges, p + cut[i], cut[i + 2] - cut[i], cut[i + 1] - cut[i], 0, NULL);
            if (s == (size_t)-1) continue;
            int64_t r = t->merges.val[s];
            if (best_rank < 0 || r < best_rank) { best_rank = r; best_i = i; }
        }
        if (best_rank < 0) break;
        /* self.merge (tokenizer.py:92-109): replace all non-overlapping occurrences left to right */
        const uint8_t *pa = p + cut[best_i]; uint32_t la = cut[best_i + 1] - cut[best_i];
        const uint8_t *pb = p + cut[best_i + 1]; uint32_t lb = cut[best_i + 2] - cut[best_i + 1];
        uint32_t o = 0, i = 0;
        /* in-place compaction of cut[]: new cut list */
        while (i < nt) {
            uint32_t li = cut[i + 1] - cut[i];
            if (li == la && memcmp(p + cut[i], pa, la) == 0 && i + 1 < nt &&
                cut[i + 2] - cut[i + 1] == lb && memcmp(p + cut[i + 1], pb, lb) == 0) {
                cut[o++] = cut[i]; i += 2;
            } else { cut[o++] = cut[i]; i += 1; }
        }
        cut[o] = len; nt = o;
    }
    int rc = 0;
    for (uint32_t i = 0; i < nt; i++) {
        size_t s = bmap_get((bmap *)&t->vocab_inv, p + cut[i], cut[i + 1] - cut[i], 0, 0, NULL);
        if (s == (size_t)-1) { *kerr_p = p + cut[i]; *kerr_len = cut[i + 1] - cut[i]; rc = -2; break; }
        lvec_push(out, t->vocab_inv.val[s]);
    }
    if (cut != stackcut) free(cut);
    return rc;
}

/* match (tokenizer.py:68-77) + per-pretoken encode for one ordinary segment */
static int encode_segment(const orc_tok *t, const uint8_t *seg, size_t n, lvec *out,
                          const uint8_t **kerr_p, uint32_t *kerr_len) {
    cptext ct;
    if (cptext_decode(seg, n, &ct)) return -1;
    size_t i = 0; int rc = 0;
    while (i < ct.n && rc == 0) {
        size_t l = gpt2_match_len(&ct, i);
        const uint8_t *p = seg + ct.off[i]; size_t blen = ct.off[i + l] - ct.off[i];
        if (!is_special(p, blen, t->sp_blob, t->sp_offs, t->n_sp))          /* tokenizer.py:73-74 */
            rc = encode_pretoken(t, p, (uint32_t)blen, out, kerr_p, kerr_len);
        i += l;
    }
    cptext_free(&ct);
    return rc;
}

/*
 * Tokenizer.encode (tokenizer.py:111-138).  Two-call idiom: ids==NULL -> only *n_out.
 * Returns 0; -1 invalid UTF-8; -2 KeyError (kerr_off/kerr_len = offending byte string within text).
 */
ORC_API int orc_encode(const orc_tok *t, const uint8_t *text, uint64_t n, int64_t *ids, uint64_t cap,
                       uint64_t *n_out, uint64_t *kerr_off, uint32_t *kerr_len) {
    lvec out = {0};
    int rc = 0; const uint8_t *kp = NULL; uint32_t kl = 0;
    /* segment (tokenizer.py:63-66): re.split on "(s1|s2|...)", specials sorted longest first */
    size_t seg_start = 0, i = 0;
    while (i <= n && rc == 0) {
        int hit = -1;
        if (i < n) for (int s = 0; s < t->n_sp; s++) {
            size_t l = t->sp_offs[s + 1] - t->sp_offs[s];
            if (l && i + l <= n && memcmp(text + i, t->sp_blob + t->sp_offs[s], l) == 0) { hit = s; break; }
        }
        if (hit >= 0 || i == n) {
            if (i > seg_start) rc = encode_segment(t, text + seg_start, i - seg_start, &out, &kp, &kl);
            if (rc) break;
            if (hit >= 0) {
                size_t l = t->sp_offs[hit + 1] - t->sp_offs[hit];
                size_t s = bmap_get((bmap *)&t->vocab_inv, text + i, (uint32_t)l, 0, 0, NULL);  /* tokenizer.py:120 */
                if (s == (size_t)-1) { kp = text + i; kl = (uint32_t)l; rc = -2; break; }
                lvec_push(&out, t->vocab_inv.val[s]);
                i += l; seg_start = i;
                continue;
            }
            break;
        }
        i++;
    }
    if (rc == -2 && kerr_off) { *kerr_off = (uint64_t)(kp - text); *kerr_len = kl; }
    *n_out = out.n;
    if (rc == 0 && ids) memcpy(ids, out.v, sizeof(int64_t) * (out.n < cap ? out.n : cap));
    free(out.v);
    return rc;
}

/* Tokenizer.decode, byte part (tokenizer.py:155-157): b"".join(vocab[i] for i in ids).
 * Returns 0, or -2 with *bad_index set when an id is not in vocab (KeyError). */
ORC_API int orc_decode_bytes(const orc_tok *t, const int64_t *ids, uint64_t n, uint8_t *out, uint64_t cap,
                             uint64_t *n_out, uint64_t *bad_index) {
    uint64_t o = 0;
    for (uint64_t i = 0; i < n; i++) {
        int64_t id = ids[i];
        if (id < 0 || id >= t->vocab_n || !t->vocab[id]) { if (bad_index) *bad_index = i; return -2; }
        uint32_t l = t->vocab_len[id];
        if (out && o + l <= cap) memcpy(out + o, t->vocab[id], l);
        o += l;
    }
    *n_out = o;
    return 0;
}
