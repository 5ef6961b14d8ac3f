// train.cu -- BPE training on the device: pretoken hash-count, word list, initial pair table,
// persistent cooperative merge loop.  Entry points: bpe_train*, bpe_count_* (include/bpe_sm100.h).
//
// Reference being replaced: models/tokenizer/train.py:142-231 (see each kernel for the exact lines).
#include <cooperative_groups.h>
#include <algorithm>
#include "kernels.h"
#include "ctx.h"

namespace cg = cooperative_groups;

// =============================================================================================
// 1. Pretoken counting  (extract_subword_frequencies, train.py:16-28)
//
// Two open-addressing tables in HBM:
//   short table: pretokens of <= 7 bytes.  The key IS the token: bytes packed little-endian in the
//                low 56 bits, length in the top byte, so one 64-bit CAS claims and publishes a slot
//                and equality is an integer compare (exact, no fingerprints).
//   long table:  pretokens of >= 8 bytes.  meta = (offset:40 | len:24) of a representative occurrence,
//                claimed by CAS; a 64-bit hash is kept beside it as a filter and every hash hit is
//                confirmed by comparing the bytes (exact).
// Counts are 64-bit atomics.  Offsets with bit 39 set address the persistent pool (imported /
// re-homed words), others the text arena of the current shard.
// =============================================================================================
#define SHORT_MAX 7u
#define META_EMPTY 0xFFFFFFFFFFFFFFFFull
#define META_LEN_BITS 24
#define META_LEN_MASK ((1ull << META_LEN_BITS) - 1)
#define META_POOL_BIT (1ull << 39)
#define MAX_TOKEN_LEN ((1u << META_LEN_BITS) - 2)

struct CountTables {
    u64 *skey; u64 *scnt; u64 scap;              // short
    u64 *lmeta; u64 *lhash; u64 *lcnt; u64 lcap; // long
    const uint8_t *text;                         // payload of the current text arena
    const uint8_t *pool;                         // persistent bytes of long words
    u64 *counters;                               // [0]=n_short [1]=n_long [2]=long_bytes [3]=overflow [4]=n_pretokens [5]=too_long
};

struct CountState {
    DevBuf skey, scnt, lmeta, lhash, lcnt, pool, counters;
    u64 scap = 0, lcap = 0, pool_used = 0;
    bool active = false;
    u64 n_pretokens = 0;
};

__device__ __forceinline__ const uint8_t *rep_ptr(const CountTables &t, u64 meta) {
    u64 off = meta >> META_LEN_BITS;
    return (off & META_POOL_BIT) ? t.pool + (off & ~META_POOL_BIT) : t.text + off;
}

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

__device__ __forceinline__ void short_add(const CountTables &t, u64 key, u64 delta) {
    u64 mask = t.scap - 1;
    u64 s = mix64(key) & mask;
    for (u64 probes = 0; probes < t.scap; probes++) {
        u64 k = t.skey[s];
        if (k == 0) {
            u64 old = atomicCAS(&t.skey[s], 0ull, key);
            if (old == 0) { atomicAdd(&t.counters[0], 1ull); k = key; }
            else k = old;
        }
        if (k == key) { atomicAdd(&t.scnt[s], delta); return; }
        s = (s + 1) & mask;
    }
    t.counters[3] = 1;
}

// p = bytes of the occurrence, off_meta = its offset in the space META addresses (text or pool bit set)
__device__ __forceinline__ void long_add(const CountTables &t, const uint8_t *p, u32 len, u64 off_meta, u64 delta) {
    u64 h = hash_long(p, len);
    u64 mask = t.lcap - 1;
    u64 s = h & mask;
    u64 mine = (off_meta << META_LEN_BITS) | len;
    for (u64 probes = 0; probes < t.lcap; probes++) {
        u64 m = t.lmeta[s];
        if (m == META_EMPTY) {
            u64 old = atomicCAS(&t.lmeta[s], META_EMPTY, mine);
            if (old == META_EMPTY) {
                t.lhash[s] = h;
                atomicAdd(&t.counters[1], 1ull);
                atomicAdd(&t.counters[2], (u64)len);
                atomicAdd(&t.lcnt[s], delta);
                return;
            }
            m = old;
        }
        if ((m & META_LEN_MASK) == len) {
            u64 hh = *((volatile u64 *)&t.lhash[s]);
            if ((hh == 0 || hh == h) && bytes_equal(rep_ptr(t, m), p, len)) { atomicAdd(&t.lcnt[s], delta); return; }
        }
        s = (s + 1) & mask;
    }
    t.counters[3] = 1;
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

// One thread per 32-byte flag word; every start bit in [own_begin, own_end) is one pretoken occurrence.
__global__ void __launch_bounds__(256) k_count_pretokens(CountTables t, const u32 *__restrict__ flags, u64 n,
                                                        u64 word_begin, u64 word_end, u64 own_begin, u64 own_end) {
    u64 n_tok = 0;
    for (u64 w = word_begin + (u64)blockIdx.x * blockDim.x + threadIdx.x; w < word_end; w += (u64)gridDim.x * blockDim.x) {
        u32 bits = flags[w];
        while (bits) {
            u32 j = __ffs(bits) - 1; bits &= bits - 1;
            u64 pos = (w << 5) + j;
            if (pos < own_begin || pos >= own_end) continue;
            u64 end = bits ? (w << 5) + (__ffs(bits) - 1) : flags_next_start(flags, (w + 1) << 5, n);
            u64 len = end - pos;
            n_tok++;
            const uint8_t *p = t.text + pos;
            if (len <= SHORT_MAX) {
                u64 key = 0;
                for (u32 k = 0; k < (u32)len; k++) key |= (u64)p[k] << (8 * k);
                key |= len << 56;
                short_add(t, key, 1);
            } else if (len <= MAX_TOKEN_LEN) {
                long_add(t, p, (u32)len, pos, 1);
            } else {
                t.counters[5] = 1;
            }
        }
    }
    // one atomic per warp for the occurrence counter
    for (int d = 16; d; d >>= 1) n_tok += __shfl_down_sync(0xffffffffu, n_tok, d);
    if (lane_id() == 0 && n_tok) atomicAdd(&t.counters[4], n_tok);
}

// upper bound of new unique words a range of flag words can create: its number of start bits
__global__ void __launch_bounds__(256) k_popc_ranges(const u32 *__restrict__ flags, u64 n_words, u64 words_per_range,
                                                    u64 *__restrict__ out) {
    u64 r = blockIdx.y;
    u64 lo = r * words_per_range, hi = lo + words_per_range;
    if (hi > n_words) hi = n_words;
    u64 c = 0;
    for (u64 w = lo + (u64)blockIdx.x * blockDim.x + threadIdx.x; w < hi; w += (u64)gridDim.x * blockDim.x) c += __popc(flags[w]);
    for (int d = 16; d; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if (lane_id() == 0 && c) atomicAdd(&out[r], c);
}

// ---- table growth: re-insert every entry of an old table into a bigger one -------------------
__global__ void __launch_bounds__(256) k_rehash_short(const u64 *__restrict__ okey, const u64 *__restrict__ ocnt, u64 ocap, CountTables t) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ocap; i += (u64)gridDim.x * blockDim.x)
        if (okey[i]) short_add(t, okey[i], ocnt[i]);
}
__global__ void __launch_bounds__(256) k_rehash_long(const u64 *__restrict__ ometa, const u64 *__restrict__ ohash,
                                                    const u64 *__restrict__ ocnt, u64 ocap, CountTables t) {
    u64 mask = t.lcap - 1;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ocap; i += (u64)gridDim.x * blockDim.x) {
        u64 m = ometa[i];
        if (m == META_EMPTY) continue;
        u64 s = ohash[i] & mask;                 // keys are unique: no comparison needed, just find a hole
        for (;;) {
            if (t.lmeta[s] == META_EMPTY && atomicCAS(&t.lmeta[s], META_EMPTY, m) == META_EMPTY) {
                t.lhash[s] = ohash[i]; t.lcnt[s] = ocnt[i];
                break;
            }
            s = (s + 1) & mask;
        }
    }
}

// ---- re-homing: copy representatives that still point into the text arena to the pool ---------
__global__ void __launch_bounds__(256) k_rehome_sizes(CountTables t, u64 *__restrict__ need /* [0] */) {
    u64 c = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.lcap; i += (u64)gridDim.x * blockDim.x) {
        u64 m = t.lmeta[i];
        if (m != META_EMPTY && !((m >> META_LEN_BITS) & META_POOL_BIT)) c += m & META_LEN_MASK;
    }
    for (int d = 16; d; d >>= 1) c += __shfl_down_sync(0xffffffffu, c, d);
    if (lane_id() == 0 && c) atomicAdd(&need[0], c);
}
__global__ void __launch_bounds__(256) k_rehome_copy(CountTables t, uint8_t *__restrict__ pool, u64 *__restrict__ cursor) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < t.lcap; i += (u64)gridDim.x * blockDim.x) {
        u64 m = t.lmeta[i];
        if (m == META_EMPTY || ((m >> META_LEN_BITS) & META_POOL_BIT)) continue;
        u32 len = (u32)(m & META_LEN_MASK);
        u64 o = atomicAdd(cursor, (u64)len);
        const uint8_t *src = t.text + (m >> META_LEN_BITS);
        for (u32 k = 0; k < len; k++) pool[o + k] = src[k];
        t.lmeta[i] = ((o | META_POOL_BIT) << META_LEN_BITS) | len;
    }
}

// ---- import of another rank's table: words = blob/offs/counts ---------------------------------
__global__ void __launch_bounds__(256) k_import_words(CountTables t, const uint8_t *__restrict__ blob_in_pool, u64 pool_base,
                                                     const u64 *__restrict__ offs, const i64 *__restrict__ counts, u64 n_words) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (u64)gridDim.x * blockDim.x) {
        u64 o = offs[i]; u64 len = offs[i + 1] - o;
        const uint8_t *p = blob_in_pool + o;
        if (len == 0) continue;
        if (len <= SHORT_MAX) {
            u64 key = 0;
            for (u32 k = 0; k < (u32)len; k++) key |= (u64)p[k] << (8 * k);
            key |= len << 56;
            short_add(t, key, (u64)counts[i]);
        } else if (len <= MAX_TOKEN_LEN) {
            long_add(t, p, (u32)len, (pool_base + o) | META_POOL_BIT, (u64)counts[i]);
        } else t.counters[5] = 1;
    }
}

// ---- export: (bytes, offs, counts) of every word ----------------------------------------------
__global__ void __launch_bounds__(256) k_export_lens(CountTables t, u32 *__restrict__ lens /* scap + lcap */) {
    u64 total = t.scap + t.lcap;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u32 l = 0;
        if (i < t.scap) { if (t.skey[i]) l = (u32)(t.skey[i] >> 56); }
        else { u64 m = t.lmeta[i - t.scap]; if (m != META_EMPTY) l = (u32)(m & META_LEN_MASK); }
        lens[i] = l;
    }
}
__global__ void __launch_bounds__(256) k_export_flags(const u32 *__restrict__ lens, u64 total, u32 *__restrict__ present) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) present[i] = lens[i] ? 1u : 0u;
}
__global__ void __launch_bounds__(256) k_export_write(CountTables t, const u32 *__restrict__ lens, const u64 *__restrict__ byte_off,
                                                     const u64 *__restrict__ word_idx, uint8_t *__restrict__ blob,
                                                     u64 *__restrict__ offs, i64 *__restrict__ counts) {
    u64 total = t.scap + t.lcap;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u32 l = lens[i];
        if (!l) continue;
        u64 o = byte_off[i], w = word_idx[i];
        offs[w] = o;
        if (i < t.scap) {
            u64 k = t.skey[i];
            for (u32 j = 0; j < l; j++) blob[o + j] = (uint8_t)(k >> (8 * j));
            counts[w] = (i64)t.scnt[i];
        } else {
            const uint8_t *src = rep_ptr(t, t.lmeta[i - t.scap]);
            for (u32 j = 0; j < l; j++) blob[o + j] = src[j];
            counts[w] = (i64)t.lcnt[i - t.scap];
        }
    }
}

// =============================================================================================
// 2. Word list + initial pair statistics
//    encode_subwords (train.py:31-32) and calculate_byte_pair_frequencies (train.py:35-49)
// =============================================================================================
struct Words {
    int32_t *sym;        // symbols of all words, word w occupies [off[w], off[w]+len[w])
    u32 *off; u32 *len;
    i64 *cnt;            // word frequency
    u32 *stamp;          // last merge step (+1) that processed the word
    u64 *counters;       // [0]=n_words [1]=n_syms [2]=max_len
};

__device__ __forceinline__ bool equals_special(const uint8_t *p, u32 len, const uint8_t *sp_blob, const u32 *sp_offs, int n_sp) {
    for (int s = 0; s < n_sp; s++) {
        u32 o = sp_offs[s], l = sp_offs[s + 1] - o;
        if (l == len && bytes_equal(p, sp_blob + o, len)) return true;
    }
    return false;
}

// Pretokens equal to a special token are dropped (train.py:25).  Words of one byte carry no pair and are
// skipped: they can never be touched by the merge loop.
__global__ void __launch_bounds__(256) k_build_words(CountTables t, Words W, const uint8_t *__restrict__ sp_blob,
                                                    const u32 *__restrict__ sp_offs, int n_sp) {
    u64 total = t.scap + t.lcap;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (u64)gridDim.x * blockDim.x) {
        u32 l = 0; const uint8_t *src = nullptr; uint8_t tmp[8]; u64 c = 0;
        if (i < t.scap) {
            u64 k = t.skey[i];
            if (!k) continue;
            l = (u32)(k >> 56);
            for (u32 j = 0; j < l; j++) tmp[j] = (uint8_t)(k >> (8 * j));
            src = tmp; c = t.scnt[i];
        } else {
            u64 m = t.lmeta[i - t.scap];
            if (m == META_EMPTY) continue;
            l = (u32)(m & META_LEN_MASK); src = rep_ptr(t, m); c = t.lcnt[i - t.scap];
        }
        if (l < 2 || c == 0) continue;
        if (n_sp && equals_special(src, l, sp_blob, sp_offs, n_sp)) continue;
        u64 w = atomicAdd(&W.counters[0], 1ull);
        u64 o = atomicAdd(&W.counters[1], (u64)l);
        atomicMax(&W.counters[2], (u64)l);
        W.off[w] = (u32)o; W.len[w] = l; W.cnt[w] = (i64)c; W.stamp[w] = 0;
        for (u32 j = 0; j < l; j++) W.sym[o + j] = src[j];
    }
}

__global__ void __launch_bounds__(256) k_init_pair_counts(Words W, u64 n_words, u64 *__restrict__ dense /* 65536 */,
                                                         u32 *__restrict__ hist /* 65536 */) {
    for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (u64)gridDim.x * blockDim.x) {
        const int32_t *s = W.sym + W.off[w]; u32 l = W.len[w]; u64 c = (u64)W.cnt[w];
        for (u32 j = 0; j + 1 < l; j++) {
            u32 p = ((u32)s[j] << 8) | (u32)s[j + 1];
            atomicAdd(&dense[p], c);
            atomicAdd(&hist[p], 1u);
        }
    }
}
// single block: exclusive scan of hist[65536] -> csr_off[65537]; also resets hist to 0 for the fill pass
__global__ void __launch_bounds__(1024) k_csr_scan(u32 *__restrict__ hist, u32 *__restrict__ csr_off) {
    __shared__ u32 s_part[1024];
    u32 tid = threadIdx.x;
    u32 sum = 0;
    for (u32 k = 0; k < 64; k++) sum += hist[tid * 64 + k];
    s_part[tid] = sum;
    __syncthreads();
    if (tid == 0) { u32 a = 0; for (u32 i = 0; i < 1024; i++) { u32 v = s_part[i]; s_part[i] = a; a += v; } csr_off[65536] = a; }
    __syncthreads();
    u32 a = s_part[tid];
    for (u32 k = 0; k < 64; k++) { u32 v = hist[tid * 64 + k]; csr_off[tid * 64 + k] = a; a += v; hist[tid * 64 + k] = 0; }
}
__global__ void __launch_bounds__(256) k_csr_fill(Words W, u64 n_words, const u32 *__restrict__ csr_off, u32 *__restrict__ fill,
                                                 u32 *__restrict__ csr_words) {
    for (u64 w = (u64)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (u64)gridDim.x * blockDim.x) {
        const int32_t *s = W.sym + W.off[w]; u32 l = W.len[w];
        for (u32 j = 0; j + 1 < l; j++) {
            u32 p = ((u32)s[j] << 8) | (u32)s[j + 1];
            csr_words[csr_off[p] + atomicAdd(&fill[p], 1u)] = (u32)w;
        }
    }
}

// =============================================================================================
// 3. Merge loop  (train.py:183-228)
// =============================================================================================
#define PAIR_EMPTY 0xFFFFFFFFFFFFFFFFull
#define CNT_DEAD ((i64)0x8000000000000000ll)     // key was popped from the dict (train.py:226)
#define PB 256u                                  // pair-table slots per cached-maximum block
#define MG_NT 512

struct Best { i64 cnt; u64 key; };

struct MergeState {
    Words W; u32 n_words;
    // pair table = byte_pair_frequencies: key (a<<32|b) -> count; a key stays (count 0 allowed) until merged
    u64 *pkey; i64 *pcnt; u64 pcap;
    Best *bmax; uint8_t *dirty; u32 n_blocks;   // per-block cached argmax
    // token_indices for pairs of two initial bytes: CSR over the initial words
    const u32 *csr_off; const u32 *csr_words;
    // token_indices for every other pair: per-step append log of (side|other symbol, word)
    uint2 *log; u64 *log_begin; u64 log_cap;
    // token byte strings and their lexicographic ranks (tie-break of train.py:187-189)
    u32 *tok_off; u32 *tok_len; uint8_t *tok_bytes; int32_t *lexrank; u64 tok_bytes_cap;
    Best *cta_best;
    int32_t *merges_out; int n_merges;
    u64 *ctr;      // [0]=log cursor [1]=n_done [2]=pairs created [3]=error flags [4]=tok bytes cursor [5]=duplicate tokens
};

// (count, (bytes_a, bytes_b)) ordering of train.py:187-189 through the rank table
__device__ __forceinline__ bool best_greater(const Best &x, const Best &y, const int32_t *__restrict__ rank) {
    if (x.cnt != y.cnt) return x.cnt > y.cnt;
    if (x.cnt == CNT_DEAD) return false;
    u32 xa = (u32)(x.key >> 32), ya = (u32)(y.key >> 32);
    if (xa != ya) return rank[xa] > rank[ya];
    u32 xb = (u32)x.key, yb = (u32)y.key;
    return rank[xb] > rank[yb];
}
__device__ __forceinline__ Best best_shfl_xor(const Best &b, int d) {
    Best o;
    o.cnt = __shfl_xor_sync(0xffffffffu, b.cnt, d);
    o.key = __shfl_xor_sync(0xffffffffu, b.key, d);
    return o;
}
__device__ __forceinline__ Best warp_best(Best b, const int32_t *rank) {
#pragma unroll
    for (int d = 16; d; d >>= 1) { Best o = best_shfl_xor(b, d); if (best_greater(o, b, rank)) b = o; }
    return b;
}

// frequencies[key] += delta with defaultdict semantics (train.py:36,65-78): a missing key is created.
__device__ __forceinline__ void pair_add(const MergeState &M, u32 a, u32 b, i64 delta) {
    u64 key = ((u64)a << 32) | b;
    u64 mask = M.pcap - 1;
    u64 s = mix64(key) & mask;
    for (u64 probes = 0; probes < M.pcap; probes++) {
        u64 k = M.pkey[s];
        if (k == PAIR_EMPTY) {
            u64 old = atomicCAS(&M.pkey[s], PAIR_EMPTY, key);
            if (old == PAIR_EMPTY) { atomicAdd(&M.ctr[2], 1ull); k = key; }
            else k = old;
        }
        if (k == key) {
            i64 old = (i64)atomicAdd((u64 *)&M.pcnt[s], (u64)delta);
            if (old == CNT_DEAD) atomicAdd((u64 *)&M.pcnt[s], (u64)CNT_DEAD);   // popped key touched again: back to 0 + delta
            M.dirty[s / PB] = 1;
            return;
        }
        s = (s + 1) & mask;
    }
    M.ctr[3] = 1;                                // table full
}

__device__ __forceinline__ void log_append(const MergeState &M, u32 key, u32 word) {
    u64 i = atomicAdd(&M.ctr[0], 1ull);
    if (i < M.log_cap) M.log[i] = make_uint2(key, word);
    else M.ctr[3] = 2;
}

// Apply merge (a,b)->nw to word w: the left-to-right scan of train.py:196-224 with
// update_frequencies_after_merge (52-78), merge_subwords (132-139) and create_new_token_indices (107-129).
__device__ __forceinline__ void apply_merge_to_word(const MergeState &M, u32 w, u32 a, u32 b, u32 nw) {
    int32_t *s = M.W.sym + M.W.off[w];
    u32 len = M.W.len[w];
    i64 c = M.W.cnt[w];
    u32 o = 0, r = 0;
    while (r + 1 < len) {
        if ((u32)s[r] == a && (u32)s[r + 1] == b) {
            if (o > 0) {
                u32 left = (u32)s[o - 1];
                pair_add(M, left, a, -c);
                pair_add(M, left, nw, c);
                log_append(M, left, w);                          // (left, nw): side L, other = left
            }
            if (r + 2 < len) {
                u32 right = (u32)s[r + 2];
                pair_add(M, b, right, -c);
                pair_add(M, nw, right, c);
                log_append(M, 0x80000000u | right, w);           // (nw, right): side R, other = right
            }
            s[o++] = (int32_t)nw; r += 2;
        } else {
            s[o++] = s[r++];
        }
    }
    while (r < len) s[o++] = s[r++];
    M.W.len[w] = o;
}

__global__ void __launch_bounds__(256) k_insert_initial_pairs(MergeState M, const u64 *__restrict__ dense) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < 65536 && dense[p]) pair_add(M, p >> 8, p & 255u, (i64)dense[p]);
}

__device__ __forceinline__ int bytes_cmp_dev(const uint8_t *x, u32 nx, const uint8_t *y, u32 ny) {
    u32 m = nx < ny ? nx : ny;
    for (u32 i = 0; i < m; i++) { if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1; }
    return nx < ny ? -1 : (nx > ny ? 1 : 0);
}

__global__ void __launch_bounds__(MG_NT) k_merge_loop(MergeState M) {
    cg::grid_group grid = cg::this_grid();
    __shared__ Best s_best[MG_NT / 32];
    __shared__ Best s_win;
    __shared__ u32 s_red[MG_NT / 32 + 1];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u32 G = gridDim.x;
    const u32 warps_per_cta = MG_NT / 32;
    const u32 gwarp = blockIdx.x * warps_per_cta + warp, total_warps = G * warps_per_cta;
    const bool token_cta = blockIdx.x == G - 1;
    const u32 apply_ctas = G - 1;
    u64 prev_key = PAIR_EMPTY;                   // winner of the previous step: popped lazily during the rescan
    u32 n_tok = 256;

    for (int step = 0; step < M.n_merges; step++) {
        if (*((volatile u64 *)&M.ctr[3])) break;  // a table overflowed in the previous step (uniform: read after grid.sync)
        // ---- phase 1: refresh cached block maxima, reduce to one candidate per CTA ----------------
        if (blockIdx.x == 0 && tid == 0) M.log_begin[step] = M.ctr[0];
        Best mine; mine.cnt = CNT_DEAD; mine.key = PAIR_EMPTY;
        for (u32 base = gwarp * 32; base < M.n_blocks; base += total_warps * 32) {
            u32 blk = base + lane;
            bool d = blk < M.n_blocks && M.dirty[blk];
            u32 dmask = __ballot_sync(0xffffffffu, d);
            while (dmask) {                      // warp-cooperative rescan of a dirty block
                u32 l = __ffs(dmask) - 1; dmask &= dmask - 1;
                u32 bb = base + l;
                Best bst; bst.cnt = CNT_DEAD; bst.key = PAIR_EMPTY;
                u64 sbase = (u64)bb * PB;
#pragma unroll
                for (u32 k = 0; k < PB / 32; k++) {
                    u64 s = sbase + k * 32 + lane;
                    u64 key = M.pkey[s];
                    if (key == PAIR_EMPTY) continue;
                    if (key == prev_key) { M.pcnt[s] = CNT_DEAD; continue; }
                    Best c; c.cnt = M.pcnt[s]; c.key = key;
                    if (c.cnt != CNT_DEAD && best_greater(c, bst, M.lexrank)) bst = c;
                }
                bst = warp_best(bst, M.lexrank);
                if (lane == 0) { M.bmax[bb] = bst; M.dirty[bb] = 0; }
            }
            __syncwarp();
            if (blk < M.n_blocks) {
                Best c = M.bmax[blk];
                if (c.cnt != CNT_DEAD && best_greater(c, mine, M.lexrank)) mine = c;
            }
        }
        mine = warp_best(mine, M.lexrank);
        if (lane == 0) s_best[warp] = mine;
        __syncthreads();
        if (warp == 0) {
            Best c = lane < warps_per_cta ? s_best[lane] : Best{CNT_DEAD, PAIR_EMPTY};
            c = warp_best(c, M.lexrank);
            if (lane == 0) M.cta_best[blockIdx.x] = c;
        }
        grid.sync();
        // ---- phase 2: every CTA derives the same winner ----------------------------------------
        if (warp == 0) {
            Best c; c.cnt = CNT_DEAD; c.key = PAIR_EMPTY;
            for (u32 i = lane; i < G; i += 32) { Best o = M.cta_best[i]; if (o.cnt != CNT_DEAD && best_greater(o, c, M.lexrank)) c = o; }
            c = warp_best(c, M.lexrank);
            if (lane == 0) s_win = c;
        }
        __syncthreads();
        const Best win = s_win;
        if (win.cnt == CNT_DEAD) break;          // len(byte_pair_frequencies) == 0 (train.py:184-185)
        const u32 a = (u32)(win.key >> 32), b = (u32)win.key;
        const u32 nw = n_tok;                    // symbol id of new_byte = a + b (train.py:190)

        if (token_cta) {
            // ---- token bookkeeping: bytes of the new token, its lexicographic rank, outputs -----
            u32 la = M.tok_len[a], lb = M.tok_len[b];
            u64 cur = M.ctr[4];
            __syncthreads();
            if (cur + la + lb > M.tok_bytes_cap) { if (tid == 0) M.ctr[3] = 4; }
            else {
                uint8_t *dst = M.tok_bytes + cur;
                const uint8_t *pa = M.tok_bytes + M.tok_off[a], *pb = M.tok_bytes + M.tok_off[b];
                for (u32 i = tid; i < la + lb; i += MG_NT) dst[i] = i < la ? pa[i] : pb[i - la];
                __syncthreads();
                u32 less = 0, dup = 0;
                for (u32 t = tid; t < n_tok; t += MG_NT) {
                    int c = bytes_cmp_dev(M.tok_bytes + M.tok_off[t], M.tok_len[t], dst, la + lb);
                    less += c <= 0 ? 1u : 0u;    // equal bytes (never expected): older token ranks first
                    dup += c == 0 ? 1u : 0u;
                }
                for (int d = 16; d; d >>= 1) { less += __shfl_down_sync(0xffffffffu, less, d); dup += __shfl_down_sync(0xffffffffu, dup, d); }
                if (lane == 0) { s_red[warp] = less | (dup ? 0x80000000u : 0u); }
                __syncthreads();
                if (tid == 0) {
                    u32 tot = 0, anydup = 0;
                    for (u32 i = 0; i < warps_per_cta; i++) { tot += s_red[i] & 0x7FFFFFFFu; anydup |= s_red[i] >> 31; }
                    s_red[warps_per_cta] = tot;
                    if (anydup && win.cnt > 0) atomicAdd(&M.ctr[5], 1ull);
                    M.tok_off[nw] = (u32)cur; M.tok_len[nw] = la + lb; M.ctr[4] = cur + la + lb;
                    M.merges_out[2 * step] = (int32_t)a; M.merges_out[2 * step + 1] = (int32_t)b;
                    M.ctr[1] = (u64)(step + 1);
                }
                __syncthreads();
                const int32_t r = (int32_t)s_red[warps_per_cta];
                for (u32 t = tid; t < n_tok; t += MG_NT) if (M.lexrank[t] >= r) M.lexrank[t]++;
                if (tid == 0) M.lexrank[nw] = r;
            }
            // make sure the winner's block is rescanned so the key gets popped
            if (tid == 0) {
                u64 mask = M.pcap - 1, s = mix64(win.key) & mask;
                while (M.pkey[s] != win.key) s = (s + 1) & mask;
                M.dirty[s / PB] = 1;
            }
        } else {
            // ---- apply the merge to every word indexed under (a,b) ------------------------------
            const u32 T = a > b ? a : b;
            const u64 gthread = (u64)blockIdx.x * MG_NT + tid, gstride = (u64)apply_ctas * MG_NT;
            if (T < 256) {
                u32 p = (a << 8) | b;
                u32 lo = M.csr_off[p], hi = M.csr_off[p + 1];
                for (u64 i = lo + gthread; i < hi; i += gstride) {
                    u32 w = M.csr_words[i];
                    if (atomicExch(&M.W.stamp[w], (u32)step + 1) != (u32)step + 1) apply_merge_to_word(M, w, a, b, nw);
                }
            } else {
                u32 t = T - 256;
                u64 lo = M.log_begin[t], hi = M.log_begin[t + 1];
                u32 want = b >= a ? a : (0x80000000u | b);
                for (u64 i = lo + gthread; i < hi; i += gstride) {
                    uint2 rec = M.log[i];
                    if (rec.x == want && atomicExch(&M.W.stamp[rec.y], (u32)step + 1) != (u32)step + 1)
                        apply_merge_to_word(M, rec.y, a, b, nw);
                }
            }
        }
        prev_key = win.key;
        n_tok++;
        grid.sync();
    }
}

// =============================================================================================
// host orchestration
// =============================================================================================
struct TrainBufs {
    DevBuf sym, off, len, cnt, stamp, wctr;
    DevBuf dense, hist, csr_off, csr_words;
    DevBuf pkey, pcnt, bmax, dirty, log, log_begin, tok_off, tok_len, tok_bytes, lexrank, cta_best, merges, ctr;
    void free_all() {
        for (DevBuf *b : {&sym, &off, &len, &cnt, &stamp, &wctr, &dense, &hist, &csr_off, &csr_words, &pkey, &pcnt, &bmax,
                          &dirty, &log, &log_begin, &tok_off, &tok_len, &tok_bytes, &lexrank, &cta_best, &merges, &ctr})
            bpe_buf_free(*b);
    }
};

static CountTables count_tables(bpe_ctx *ctx) {
    CountState *cs = ctx->count;
    CountTables t;
    t.skey = (u64 *)cs->skey.p; t.scnt = (u64 *)cs->scnt.p; t.scap = cs->scap;
    t.lmeta = (u64 *)cs->lmeta.p; t.lhash = (u64 *)cs->lhash.p; t.lcnt = (u64 *)cs->lcnt.p; t.lcap = cs->lcap;
    t.text = ctx->text.p ? (const uint8_t *)ctx->text.p + BPE_PAD : nullptr;
    t.pool = (const uint8_t *)cs->pool.p;
    t.counters = (u64 *)cs->counters.p;
    return t;
}

void count_state_free(bpe_ctx *ctx) {
    if (!ctx->count) return;
    CountState *cs = ctx->count;
    for (DevBuf *b : {&cs->skey, &cs->scnt, &cs->lmeta, &cs->lhash, &cs->lcnt, &cs->pool, &cs->counters}) bpe_buf_free(*b);
    delete cs;
    ctx->count = nullptr;
}

static int alloc_exact(bpe_ctx *ctx, DevBuf &b, size_t bytes) {
    if (b.cap >= bytes && b.cap <= bytes * 2 + (1 << 20)) return BPE_OK;
    bpe_buf_free(b);
    cudaError_t e = cudaMalloc(&b.p, bytes ? bytes : 256);
    if (e != cudaSuccess) { cudaGetLastError(); b.p = nullptr; return bpe_set_error(ctx, BPE_ERR_OOM, "cudaMalloc(%zu) failed", bytes); }
    b.cap = bytes ? bytes : 256;
    return BPE_OK;
}

static int count_tables_alloc(bpe_ctx *ctx, u64 scap, u64 lcap) {
    CountState *cs = ctx->count;
    BPE_TRY(alloc_exact(ctx, cs->skey, scap * 8)); BPE_TRY(alloc_exact(ctx, cs->scnt, scap * 8));
    BPE_TRY(alloc_exact(ctx, cs->lmeta, lcap * 8)); BPE_TRY(alloc_exact(ctx, cs->lhash, lcap * 8));
    BPE_TRY(alloc_exact(ctx, cs->lcnt, lcap * 8));
    cudaStream_t st = ctx->stream;
    CUDA_TRY(ctx, cudaMemsetAsync(cs->skey.p, 0, scap * 8, st)); CUDA_TRY(ctx, cudaMemsetAsync(cs->scnt.p, 0, scap * 8, st));
    CUDA_TRY(ctx, cudaMemsetAsync(cs->lmeta.p, 0xFF, lcap * 8, st)); CUDA_TRY(ctx, cudaMemsetAsync(cs->lhash.p, 0, lcap * 8, st));
    CUDA_TRY(ctx, cudaMemsetAsync(cs->lcnt.p, 0, lcap * 8, st));
    cs->scap = scap; cs->lcap = lcap;
    return BPE_OK;
}

BPE_API int bpe_count_begin(bpe_ctx *ctx) {
    if (!ctx) return BPE_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->count) ctx->count = new CountState();
    CountState *cs = ctx->count;
    BPE_TRY(bpe_buf_reserve(ctx, cs->counters, 64 * sizeof(u64)));
    CUDA_TRY(ctx, cudaMemsetAsync(cs->counters.p, 0, 64 * sizeof(u64), ctx->stream));
    BPE_TRY(count_tables_alloc(ctx, 1 << 16, 1 << 14));
    cs->pool_used = 0; cs->active = true; cs->n_pretokens = 0;
    return BPE_OK;
}

static int read_counters(bpe_ctx *ctx, u64 *out, int k) {
    u64 *host = (u64 *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(host, ctx->count->counters.p, k * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < k; i++) out[i] = host[i];
    return BPE_OK;
}

// Make sure the tables can absorb `new_short` / `new_long` more unique words without exceeding 7/8 load.
static int count_ensure_capacity(bpe_ctx *ctx, u64 n_short, u64 n_long, u64 new_short, u64 new_long) {
    CountState *cs = ctx->count;
    u64 need_s = next_pow2(std::max<u64>(1 << 16, (n_short + new_short) * 8 / 7 + 64));
    u64 need_l = next_pow2(std::max<u64>(1 << 14, (n_long + new_long) * 8 / 7 + 64));
    if (need_s <= cs->scap && need_l <= cs->lcap) return BPE_OK;
    need_s = std::max(need_s, cs->scap); need_l = std::max(need_l, cs->lcap);
    // grow: move old tables aside, allocate, re-insert
    CountState old_view = *cs;                   // shallow copy of the DevBufs
    cs->skey = DevBuf(); cs->scnt = DevBuf(); cs->lmeta = DevBuf(); cs->lhash = DevBuf(); cs->lcnt = DevBuf();
    int rc = count_tables_alloc(ctx, need_s, need_l);
    if (rc != BPE_OK) return rc;
    // the short rehash re-counts its uniques through short_add: reset [0]; [1],[2] (long) are untouched
    CUDA_TRY(ctx, cudaMemsetAsync(cs->counters.p, 0, sizeof(u64), ctx->stream));
    CountTables t = count_tables(ctx);
    unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (old_view.scap + 255) / 256);
    k_rehash_short<<<grid, 256, 0, ctx->stream>>>((const u64 *)old_view.skey.p, (const u64 *)old_view.scnt.p, old_view.scap, t);
    grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (old_view.lcap + 255) / 256);
    k_rehash_long<<<grid, 256, 0, ctx->stream>>>((const u64 *)old_view.lmeta.p, (const u64 *)old_view.lhash.p,
                                                 (const u64 *)old_view.lcnt.p, old_view.lcap, t);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (DevBuf *b : {&old_view.skey, &old_view.scnt, &old_view.lmeta, &old_view.lhash, &old_view.lcnt}) bpe_buf_free(*b);
    return BPE_OK;
}

#define COUNT_BATCH_BYTES (256ull << 20)

// Count the pretokens whose start lies in [own_begin, own_end) of the text currently in the arena
// (flags already computed).
static int count_current_text(bpe_ctx *ctx, u64 n, u64 own_begin, u64 own_end) {
    CountState *cs = ctx->count;
    cudaStream_t st = ctx->stream;
    if (own_end > n) own_end = n;
    if (own_begin >= own_end) return BPE_OK;
    u64 w_lo = own_begin / 32, w_hi = (own_end + 31) / 32;
    u64 words_per_batch = COUNT_BATCH_BYTES / 32;
    u64 n_batches = (w_hi - w_lo + words_per_batch - 1) / words_per_batch;
    // per-batch upper bound of new uniques = number of start bits
    BPE_TRY(bpe_buf_reserve(ctx, ctx->tmp0, n_batches * sizeof(u64)));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->tmp0.p, 0, n_batches * sizeof(u64), st));
    {
        dim3 grid((unsigned)std::min<u64>(ctx->sm_count * 4, (words_per_batch + 255) / 256), (unsigned)n_batches);
        k_popc_ranges<<<grid, 256, 0, st>>>((const u32 *)ctx->flags.p + w_lo, w_hi - w_lo, words_per_batch, (u64 *)ctx->tmp0.p);
    }
    std::vector<u64> bound(n_batches);
    CUDA_TRY(ctx, cudaMemcpyAsync(bound.data(), ctx->tmp0.p, n_batches * sizeof(u64), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    u64 c[6];
    BPE_TRY(read_counters(ctx, c, 6));
    for (u64 bi = 0; bi < n_batches; bi++) {
        u64 b_lo = w_lo + bi * words_per_batch, b_hi = std::min(w_hi, b_lo + words_per_batch);
        u64 bytes = (b_hi - b_lo) * 32;
        BPE_TRY(count_ensure_capacity(ctx, c[0], c[1], bound[bi], std::min(bound[bi], bytes / (SHORT_MAX + 1) + 1)));
        CountTables t = count_tables(ctx);
        unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (b_hi - b_lo + 255) / 256);
        k_count_pretokens<<<grid, 256, 0, st>>>(t, (const u32 *)ctx->flags.p, n, b_lo, b_hi, own_begin, own_end);
        CUDA_TRY(ctx, cudaGetLastError());
        if (bi + 1 < n_batches) BPE_TRY(read_counters(ctx, c, 6));
    }
    BPE_TRY(read_counters(ctx, c, 6));
    if (c[3]) return bpe_set_error(ctx, BPE_ERR_CAPACITY, "pretoken table overflow");
    if (c[5]) return bpe_set_error(ctx, BPE_ERR_UNSUPPORTED, "a pretoken longer than %u bytes", MAX_TOKEN_LEN);
    cs->n_pretokens = c[4];
    return BPE_OK;
}

// Copy long-word representatives out of the (transient) text arena into the persistent pool.
static int count_rehome(bpe_ctx *ctx) {
    CountState *cs = ctx->count;
    cudaStream_t st = ctx->stream;
    CountTables t = count_tables(ctx);
    u64 *scr = (u64 *)ctx->scratch.p + 8;
    CUDA_TRY(ctx, cudaMemsetAsync(scr, 0, 16, st));
    unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (cs->lcap + 255) / 256);
    k_rehome_sizes<<<grid, 256, 0, st>>>(t, scr);
    u64 *host = (u64 *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(host, scr, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    u64 need = host[0];
    if (!need) return BPE_OK;
    if (cs->pool_used + need > cs->pool.cap) {
        DevBuf nb;
        BPE_TRY(bpe_buf_reserve(ctx, nb, (cs->pool_used + need) * 2 + (1 << 20)));
        if (cs->pool_used) CUDA_TRY(ctx, cudaMemcpyAsync(nb.p, cs->pool.p, cs->pool_used, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        bpe_buf_free(cs->pool);
        cs->pool = nb;
        t = count_tables(ctx);
    }
    host[0] = cs->pool_used;
    CUDA_TRY(ctx, cudaMemcpyAsync(scr + 1, host, 8, cudaMemcpyHostToDevice, st));
    k_rehome_copy<<<grid, 256, 0, st>>>(t, (uint8_t *)cs->pool.p, scr + 1);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    cs->pool_used += need;
    return BPE_OK;
}

BPE_API int bpe_count_add_shard(bpe_ctx *ctx, const uint8_t *text_host, uint64_t n, uint64_t own_begin, uint64_t own_end,
                                int at_file_start, int at_file_end) {
    (void)at_file_start; (void)at_file_end;      // the halo bytes carry all the context the stencil needs
    if (!ctx || !ctx->count || !ctx->count->active || (!text_host && n) || own_begin > own_end || own_end > n) return BPE_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    BPE_TRY(ctx_load_text(ctx, text_host, n, false));
    u64 nn = n;
    BPE_TRY(ctx_run_flags(ctx, &nn, false, nullptr, nullptr, 0, 0));
    BPE_TRY(count_current_text(ctx, n, own_begin, own_end));
    return count_rehome(ctx);
}

BPE_API int bpe_count_export_size(bpe_ctx *ctx, uint64_t *n_words, uint64_t *blob_bytes) {
    if (!ctx || !ctx->count || !n_words || !blob_bytes) return BPE_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CountState *cs = ctx->count;
    cudaStream_t st = ctx->stream;
    CountTables t = count_tables(ctx);
    u64 total = cs->scap + cs->lcap;
    size_t lens_b = round_up(total * 4, 256), pres_b = lens_b, boff_b = round_up((total + 1) * 8, 256), widx_b = boff_b;
    size_t tmp_b = scan_tmp_elems_host(total) * 8;
    BPE_TRY(bpe_buf_reserve(ctx, ctx->tmp1, lens_b + pres_b + boff_b + widx_b + tmp_b));
    u32 *lens = (u32 *)ctx->tmp1.p; u32 *pres = (u32 *)((uint8_t *)lens + lens_b);
    u64 *boff = (u64 *)((uint8_t *)pres + pres_b); u64 *widx = (u64 *)((uint8_t *)boff + boff_b);
    u64 *tmp = (u64 *)((uint8_t *)widx + widx_b);
    unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (total + 255) / 256);
    k_export_lens<<<grid, 256, 0, st>>>(t, lens);
    k_export_flags<<<grid, 256, 0, st>>>(lens, total, pres);
    launch_scan_u32(lens, total, boff, tmp, st);
    launch_scan_u32(pres, total, widx, tmp, st);
    u64 *host = (u64 *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(host, boff + total, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(host + 1, widx + total, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    *blob_bytes = host[0]; *n_words = host[1];
    return BPE_OK;
}

BPE_API int bpe_count_export(bpe_ctx *ctx, uint8_t *blob, uint64_t *offs, int64_t *counts) {
    if (!ctx || !ctx->count || !offs) return BPE_ERR_ARG;
    uint64_t nw, nb;
    BPE_TRY(bpe_count_export_size(ctx, &nw, &nb));   // recomputes the scans in tmp1
    CountState *cs = ctx->count;
    cudaStream_t st = ctx->stream;
    CountTables t = count_tables(ctx);
    u64 total = cs->scap + cs->lcap;
    size_t lens_b = round_up(total * 4, 256), pres_b = lens_b, boff_b = round_up((total + 1) * 8, 256);
    u32 *lens = (u32 *)ctx->tmp1.p; u32 *pres = (u32 *)((uint8_t *)lens + lens_b);
    u64 *boff = (u64 *)((uint8_t *)pres + pres_b); u64 *widx = (u64 *)((uint8_t *)boff + boff_b);
    size_t ob = round_up(nb + 1, 256), oo = round_up((nw + 1) * 8, 256), oc = round_up((nw + 1) * 8, 256);
    BPE_TRY(bpe_buf_reserve(ctx, ctx->tmp0, ob + oo + oc));
    uint8_t *dblob = (uint8_t *)ctx->tmp0.p; u64 *doffs = (u64 *)(dblob + ob); i64 *dcnt = (i64 *)(dblob + ob + oo);
    unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (total + 255) / 256);
    k_export_write<<<grid, 256, 0, st>>>(t, lens, boff, widx, dblob, doffs, dcnt);
    CUDA_TRY(ctx, cudaGetLastError());
    if (nb && blob) CUDA_TRY(ctx, cudaMemcpyAsync(blob, dblob, nb, cudaMemcpyDeviceToHost, st));
    if (nw) CUDA_TRY(ctx, cudaMemcpyAsync(offs, doffs, nw * 8, cudaMemcpyDeviceToHost, st));
    if (nw && counts) CUDA_TRY(ctx, cudaMemcpyAsync(counts, dcnt, nw * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    offs[nw] = nb;
    return BPE_OK;
}

BPE_API int bpe_count_import(bpe_ctx *ctx, const uint8_t *blob, const uint64_t *offs, const int64_t *counts, uint64_t n_words) {
    if (!ctx || !ctx->count || !ctx->count->active || (n_words && (!offs || !counts))) return BPE_ERR_ARG;
    if (!n_words) return BPE_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CountState *cs = ctx->count;
    cudaStream_t st = ctx->stream;
    u64 nb = offs[n_words];
    // blob goes straight into the pool (long words keep pointing at it)
    if (cs->pool_used + nb > cs->pool.cap) {
        DevBuf nbuf;
        BPE_TRY(bpe_buf_reserve(ctx, nbuf, (cs->pool_used + nb) * 2 + (1 << 20)));
        if (cs->pool_used) CUDA_TRY(ctx, cudaMemcpyAsync(nbuf.p, cs->pool.p, cs->pool_used, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        bpe_buf_free(cs->pool);
        cs->pool = nbuf;
    }
    u64 base = cs->pool_used;
    if (nb) CUDA_TRY(ctx, cudaMemcpyAsync((uint8_t *)cs->pool.p + base, blob, nb, cudaMemcpyHostToDevice, st));
    cs->pool_used += nb;
    size_t ob = round_up((n_words + 1) * 8, 256);
    BPE_TRY(bpe_buf_reserve(ctx, ctx->tmp0, ob * 2));
    u64 *doffs = (u64 *)ctx->tmp0.p; i64 *dcnt = (i64 *)((uint8_t *)ctx->tmp0.p + ob);
    CUDA_TRY(ctx, cudaMemcpyAsync(doffs, offs, (n_words + 1) * 8, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(dcnt, counts, n_words * 8, cudaMemcpyHostToDevice, st));
    u64 c[6];
    BPE_TRY(read_counters(ctx, c, 6));
    BPE_TRY(count_ensure_capacity(ctx, c[0], c[1], n_words, n_words));
    CountTables t = count_tables(ctx);
    unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (n_words + 255) / 256);
    k_import_words<<<grid, 256, 0, st>>>(t, (const uint8_t *)cs->pool.p + base, base, doffs, dcnt, n_words);
    CUDA_TRY(ctx, cudaGetLastError());
    BPE_TRY(read_counters(ctx, c, 6));
    if (c[3]) return bpe_set_error(ctx, BPE_ERR_CAPACITY, "pretoken table overflow on import");
    return BPE_OK;
}

// ---- merge phase ---------------------------------------------------------------------------------
static int run_merges(bpe_ctx *ctx, const uint8_t *sp_blob, const u32 *sp_offs, int n_sp, int n_merges,
                      int32_t *merge_pairs_out, int *n_done, bpe_train_stats *stats, EvTimer &tm, int ev_start) {
    cudaStream_t st = ctx->stream;
    CountState *cs = ctx->count;
    TrainBufs B;
    struct Guard { TrainBufs &b; ~Guard() { b.free_all(); } } guard{B};
    u64 c[6];
    BPE_TRY(read_counters(ctx, c, 6));
    u64 n_short = c[0], n_long = c[1], long_bytes = c[2];
    u64 max_words = n_short + n_long, max_syms = n_short * SHORT_MAX + long_bytes;
    if (max_syms >= (1ull << 32)) return bpe_set_error(ctx, BPE_ERR_UNSUPPORTED, "more than 2^32 symbols in unique words");
    const uint8_t *spb; const u32 *spo; u32 spmax;
    BPE_TRY(ctx_upload_specials(ctx, sp_blob, sp_offs, n_sp, &spb, &spo, &spmax));

    int ev_build0 = tm.mark();
    BPE_TRY(alloc_exact(ctx, B.sym, (max_syms + 1) * 4)); BPE_TRY(alloc_exact(ctx, B.off, (max_words + 1) * 4));
    BPE_TRY(alloc_exact(ctx, B.len, (max_words + 1) * 4)); BPE_TRY(alloc_exact(ctx, B.cnt, (max_words + 1) * 8));
    BPE_TRY(alloc_exact(ctx, B.stamp, (max_words + 1) * 4)); BPE_TRY(alloc_exact(ctx, B.wctr, 64));
    CUDA_TRY(ctx, cudaMemsetAsync(B.wctr.p, 0, 64, st));
    Words W{(int32_t *)B.sym.p, (u32 *)B.off.p, (u32 *)B.len.p, (i64 *)B.cnt.p, (u32 *)B.stamp.p, (u64 *)B.wctr.p};
    CountTables t = count_tables(ctx);
    {
        u64 total = cs->scap + cs->lcap;
        unsigned grid = (unsigned)std::min<u64>((u64)ctx->sm_count * 8, (total + 255) / 256);
        k_build_words<<<grid, 256, 0, st>>>(t, W, spb, spo, n_sp);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    u64 *host = (u64 *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(host, B.wctr.p, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    u64 n_words = host[0], n_syms = host[1], max_len = host[2];

    BPE_TRY(alloc_exact(ctx, B.dense, 65536 * 8)); BPE_TRY(alloc_exact(ctx, B.hist, 65536 * 4));
    BPE_TRY(alloc_exact(ctx, B.csr_off, 65537 * 4)); BPE_TRY(alloc_exact(ctx, B.csr_words, (n_syms + 1) * 4));
    CUDA_TRY(ctx, cudaMemsetAsync(B.dense.p, 0, 65536 * 8, st)); CUDA_TRY(ctx, cudaMemsetAsync(B.hist.p, 0, 65536 * 4, st));
    unsigned wgrid = (unsigned)std::max<u64>(1, std::min<u64>((u64)ctx->sm_count * 8, (n_words + 255) / 256));
    k_init_pair_counts<<<wgrid, 256, 0, st>>>(W, n_words, (u64 *)B.dense.p, (u32 *)B.hist.p);
    k_csr_scan<<<1, 1024, 0, st>>>((u32 *)B.hist.p, (u32 *)B.csr_off.p);
    k_csr_fill<<<wgrid, 256, 0, st>>>(W, n_words, (const u32 *)B.csr_off.p, (u32 *)B.hist.p, (u32 *)B.csr_words.p);
    CUDA_TRY(ctx, cudaGetLastError());
    std::vector<u64> dense(65536);
    CUDA_TRY(ctx, cudaMemcpyAsync(dense.data(), B.dense.p, 65536 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));

    // pair table capacity: every key ever created = initial pairs + <= 2 per merge site (<= n_syms sites)
    u64 n_pairs0 = 0;
    for (u64 v : dense) n_pairs0 += v != 0;
    u64 pcap = next_pow2(std::max<u64>(1 << 12, 2 * (n_pairs0 + 2 * n_syms)));
    if (pcap > (1ull << 28)) pcap = 1ull << 28;
    MergeState M;
    memset(&M, 0, sizeof(M));
    M.W = W; M.n_words = (u32)n_words; M.pcap = pcap; M.n_blocks = (u32)(pcap / PB);
    BPE_TRY(alloc_exact(ctx, B.pkey, pcap * 8)); BPE_TRY(alloc_exact(ctx, B.pcnt, pcap * 8));
    BPE_TRY(alloc_exact(ctx, B.bmax, (u64)M.n_blocks * sizeof(Best))); BPE_TRY(alloc_exact(ctx, B.dirty, M.n_blocks));
    u64 log_cap = 2 * n_syms + 16;
    BPE_TRY(alloc_exact(ctx, B.log, log_cap * 8)); BPE_TRY(alloc_exact(ctx, B.log_begin, ((u64)n_merges + 2) * 8));
    u64 n_tok_max = 256 + (u64)n_merges;
    u64 tok_bytes_cap = 256 + (u64)n_merges * 2 * std::max<u64>(max_len, 1);
    if (tok_bytes_cap > (4ull << 30)) tok_bytes_cap = 4ull << 30;
    BPE_TRY(alloc_exact(ctx, B.tok_off, n_tok_max * 4)); BPE_TRY(alloc_exact(ctx, B.tok_len, n_tok_max * 4));
    BPE_TRY(alloc_exact(ctx, B.tok_bytes, tok_bytes_cap)); BPE_TRY(alloc_exact(ctx, B.lexrank, n_tok_max * 4));
    BPE_TRY(alloc_exact(ctx, B.merges, ((u64)n_merges + 1) * 8)); BPE_TRY(alloc_exact(ctx, B.ctr, 64));
    CUDA_TRY(ctx, cudaMemsetAsync(B.pkey.p, 0xFF, pcap * 8, st)); CUDA_TRY(ctx, cudaMemsetAsync(B.pcnt.p, 0, pcap * 8, st));
    CUDA_TRY(ctx, cudaMemsetAsync(B.dirty.p, 1, M.n_blocks, st));
    CUDA_TRY(ctx, cudaMemsetAsync(B.ctr.p, 0, 64, st));
    CUDA_TRY(ctx, cudaMemsetAsync(B.log_begin.p, 0, ((u64)n_merges + 2) * 8, st));
    {   // byte tokens: bytes(i) = [i], rank(i) = i
        std::vector<u32> toff(256), tlen(256, 1); std::vector<int32_t> rk(256); std::vector<uint8_t> tb(256);
        for (int i = 0; i < 256; i++) { toff[i] = i; rk[i] = i; tb[i] = (uint8_t)i; }
        CUDA_TRY(ctx, cudaMemcpyAsync(B.tok_off.p, toff.data(), 1024, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(B.tok_len.p, tlen.data(), 1024, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(B.lexrank.p, rk.data(), 1024, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpyAsync(B.tok_bytes.p, tb.data(), 256, cudaMemcpyHostToDevice, st));
        host[0] = 256;
        CUDA_TRY(ctx, cudaMemcpyAsync((u64 *)B.ctr.p + 4, host, 8, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
    }
    M.pkey = (u64 *)B.pkey.p; M.pcnt = (i64 *)B.pcnt.p; M.bmax = (Best *)B.bmax.p; M.dirty = (uint8_t *)B.dirty.p;
    M.csr_off = (const u32 *)B.csr_off.p; M.csr_words = (const u32 *)B.csr_words.p;
    M.log = (uint2 *)B.log.p; M.log_begin = (u64 *)B.log_begin.p; M.log_cap = log_cap;
    M.tok_off = (u32 *)B.tok_off.p; M.tok_len = (u32 *)B.tok_len.p; M.tok_bytes = (uint8_t *)B.tok_bytes.p;
    M.lexrank = (int32_t *)B.lexrank.p; M.tok_bytes_cap = tok_bytes_cap;
    M.merges_out = (int32_t *)B.merges.p; M.n_merges = n_merges; M.ctr = (u64 *)B.ctr.p;

    // initial pair table from the dense 256x256 counts (same device-side insert as the merge loop uses)
    k_insert_initial_pairs<<<256, 256, 0, st>>>(M, (const u64 *)B.dense.p);
    CUDA_TRY(ctx, cudaGetLastError());
    int ev_build1 = tm.mark();

    int G = ctx->sm_count;
    {
        int per_sm = 0;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_merge_loop, MG_NT, 0));
        if (per_sm < 1) return bpe_set_error(ctx, BPE_ERR_CUDA, "merge kernel does not fit on an SM");
        int want = (int)std::min<u64>((u64)ctx->sm_count, (u64)M.n_blocks / 64 + 2);
        G = std::max(2, want);
    }
    BPE_TRY(alloc_exact(ctx, B.cta_best, (u64)G * sizeof(Best)));
    M.cta_best = (Best *)B.cta_best.p;
    if (n_merges > 0) {
        void *args[] = {&M};
        CUDA_TRY(ctx, cudaLaunchCooperativeKernel((void *)k_merge_loop, dim3(G), dim3(MG_NT), args, 0, st));
    }
    int ev_merge1 = tm.mark();
    CUDA_TRY(ctx, cudaMemcpyAsync(host, B.ctr.p, 64, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    u64 ctr[8];
    for (int i = 0; i < 8; i++) ctr[i] = host[i];
    if (ctr[3]) return bpe_set_error(ctx, BPE_ERR_CAPACITY, "merge loop table overflow (code %llu)", (unsigned long long)ctr[3]);
    int done = (int)ctr[1];
    if (done > 0) CUDA_TRY(ctx, cudaMemcpyAsync(merge_pairs_out, B.merges.p, (size_t)done * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    *n_done = done;
    int ev_end = tm.mark();
    if (stats) {
        stats->n_unique = n_words; stats->n_symbols = n_syms; stats->n_pairs_initial = n_pairs0;
        stats->n_pairs_final = ctr[2]; stats->log_records = ctr[0]; stats->duplicate_tokens = ctr[5];
        stats->n_pretokens = cs->n_pretokens;
        stats->ms_build = tm.ms(ev_build0, ev_build1); stats->ms_merge = tm.ms(ev_build1, ev_merge1);
        stats->ms_total = tm.ms(ev_start, ev_end);
    }
    return BPE_OK;
}

static int train_impl(bpe_ctx *ctx, const uint8_t *text, u64 n, bool text_is_device, const uint8_t *sp_blob, const u32 *sp_offs,
                      int n_sp, int n_merges, int32_t *merge_pairs_out, int *n_done, bpe_train_stats *stats) {
    if (!ctx || (!text && n) || !n_done || (n_merges > 0 && !merge_pairs_out) || (n_sp > 0 && (!sp_blob || !sp_offs))) return BPE_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    *n_done = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_merges < 0) n_merges = 0;
    EvTimer tm(ctx);
    int e0 = tm.mark();
    BPE_TRY(ctx_load_text(ctx, text, n, text_is_device));
    int e1 = tm.mark();
    u64 nn = n;
    BPE_TRY(ctx_run_flags(ctx, &nn, true, nullptr, nullptr, 0, 0));
    int e2 = tm.mark();
    BPE_TRY(bpe_count_begin(ctx));
    BPE_TRY(count_current_text(ctx, nn, 0, nn));
    int e3 = tm.mark();
    int rc = run_merges(ctx, sp_blob, sp_offs, n_sp, n_merges, merge_pairs_out, n_done, stats, tm, e0);
    if (stats) {
        stats->n_bytes = nn;
        stats->ms_h2d = tm.ms(e0, e1); stats->ms_pretok = tm.ms(e1, e2); stats->ms_count = tm.ms(e2, e3);
    }
    ctx->count->active = false;
    return rc;
}

BPE_API int bpe_train(bpe_ctx *ctx, const uint8_t *text_host, uint64_t n, const uint8_t *specials_blob,
                      const uint32_t *special_offs, int n_specials, int n_merges, int32_t *merge_pairs_out, int *n_done,
                      bpe_train_stats *stats) {
    return train_impl(ctx, text_host, n, false, specials_blob, special_offs, n_specials, n_merges, merge_pairs_out, n_done, stats);
}
BPE_API int bpe_train_dev(bpe_ctx *ctx, const uint8_t *text_dev, uint64_t n, const uint8_t *specials_blob,
                          const uint32_t *special_offs, int n_specials, int n_merges, int32_t *merge_pairs_out, int *n_done,
                          bpe_train_stats *stats) {
    return train_impl(ctx, text_dev, n, true, specials_blob, special_offs, n_specials, n_merges, merge_pairs_out, n_done, stats);
}
BPE_API int bpe_train_from_counts(bpe_ctx *ctx, const uint8_t *specials_blob, const uint32_t *special_offs, int n_specials,
                                  int n_merges, int32_t *merge_pairs_out, int *n_done, bpe_train_stats *stats) {
    if (!ctx || !ctx->count || !n_done || (n_merges > 0 && !merge_pairs_out)) return BPE_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    *n_done = 0;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n_merges < 0) n_merges = 0;
    EvTimer tm(ctx);
    int e0 = tm.mark();
    int rc = run_merges(ctx, specials_blob, special_offs, n_specials, n_merges, merge_pairs_out, n_done, stats, tm, e0);
    ctx->count->active = false;
    return rc;
}
