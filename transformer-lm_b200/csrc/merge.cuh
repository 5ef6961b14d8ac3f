// merge.cuh -- the BPE merge loop as one persistent cooperative kernel.
//
// Reference: the loop at models/tokenizer/train.py:183-228:
//     best = max(byte_pair_frequencies, key=lambda x: (byte_pair_frequencies[x], x))        187-189
//     for every word indexed under best: left-to-right rewrite + 4 count updates             192-224
//     pop best from both dicts, merges.append(best)                                          226-228
//
// Device data structures
//   pair table   open addressing, key = a<<32|b, 64-bit count.  A key lives from its first touch
//                (defaultdict semantics: counts may be 0) until it is merged (count = CNT_DEAD).
//   argmax       every PB slots form a block with a cached maximum; an update marks its block dirty,
//                so a step rescans only the blocks that changed.  The cached maxima of the blocks a CTA owns
//                live in its shared memory for the whole launch (written through to bmax for the next one).
//   tie-break    (count, (bytes_a, bytes_b)) with python's bytes ordering.  Each token keeps its
//                first 8 bytes as a big-endian integer (tok_key): comparing the integers decides
//                almost every tie; equal prefixes fall back to lengths / a byte loop.
//   words        symbols never move: a merge writes the new token over its left operand and a tombstone over the
//                right one, so an occurrence is addressed by the position of its left symbol for ever.
//   token_indices  one record (neighbour, position, word count) per pair OCCURRENCE.  Pairs of two initial bytes:
//                CSR built once; any other pair (p,q): every occurrence is created in the step that created the
//                younger of p,q, so it is found in that step's slice of an append-only log (bucket-sorted by
//                neighbour when large).  Stale records are harmless (the symbols are re-checked, like the
//                reference re-checks its stale index entries, train.py:196-200).  One warp per occurrence in most
//                steps: no per-word serialisation, no divergence between sites.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include "common.cuh"

namespace cg = cooperative_groups;

// slot hash of the pair table: 32-bit multiplies only (the apply path computes four of these per site)
__device__ __forceinline__ u32 pair_hash(u64 key) {
    u32 h = ((u32)(key >> 32) * 0x9E3779B1u) ^ ((u32)key * 0x85EBCA77u);
    h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15;
    return h;
}

#define PAIR_EMPTY 0xFFFFFFFFFFFFFFFFull
#define CNT_DEAD ((i64)0x8000000000000000ll)     // key was popped from the dict (train.py:226)
#define CNT_DEAD_LIMIT ((i64)0xC000000000000000ll) // counts below this are "popped (+ later deltas)"
#define PB 64u                                   // pair-table slots per block (one rescan = two slots per lane)
#define MG_NT 512
#define MG_NEED_GROW 8ull                        // ctr[3] code: pair table more than half full, host must grow it

struct __align__(16) WordMeta {
    u32 off, len;        // word w occupies sym[off, off+len) at build time
    i64 cnt;             // word frequency
};
// Symbol array: every word is preceded by one SYM_SEP; the array starts and ends with SYM_PAD separators, so the
// neighbour scans of apply_site never need a bounds check.  Values: >= 0 live symbol, SYM_SEP word boundary,
// <= -2 tombstone written by merge step (-v - 2).
#define SYM_SEP (-1)
#define SYM_PAD 32u
struct Words {
    int32_t *sym;
    WordMeta *meta;
    u64 *counters;       // [0]=n_words [1]=sym slots used (symbols + one separator per word) [2]=max_len
};
// index record: one pair occurrence.  x = neighbour token (bit 31: the neighbour is on the right), pos = position of
// the occurrence's left symbol, cnt = frequency of the word it lies in
struct __align__(16) Rec { u32 x, pos; i64 cnt; };
__device__ __forceinline__ Rec load_rec(const Rec *p) {
    uint4 v = *reinterpret_cast<const uint4 *>(p);
    Rec r; r.x = v.x; r.pos = v.y; r.cnt = (i64)(((u64)v.w << 32) | v.z); return r;
}
__device__ __forceinline__ void store_rec(Rec *p, u32 x, u32 pos, i64 cnt) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(x, pos, (u32)cnt, (u32)((u64)cnt >> 32));
}

// candidate of the argmax: count, pair key, and the 8-byte prefix keys of both tokens (so that almost every
// tie is broken in registers, without dependent loads)
struct __align__(32) Best { i64 cnt; u64 key; u64 ka; u64 kb; };
#define BEST_NONE Best{CNT_DEAD, PAIR_EMPTY, 0, 0}
// whole-struct 2 x 128-bit accesses (otherwise the compiler loads .cnt first and the rest behind a branch:
// two dependent round trips instead of one)
__device__ __forceinline__ Best load_best(const Best *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = q[0], hi = q[1];
    Best b;
    b.cnt = (i64)(((u64)lo.y << 32) | lo.x); b.key = ((u64)lo.w << 32) | lo.z;
    b.ka = ((u64)hi.y << 32) | hi.x; b.kb = ((u64)hi.w << 32) | hi.z;
    return b;
}
__device__ __forceinline__ void store_best(Best *p, const Best &b) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4((u32)b.cnt, (u32)((u64)b.cnt >> 32), (u32)b.key, (u32)(b.key >> 32));
    q[1] = make_uint4((u32)b.ka, (u32)(b.ka >> 32), (u32)b.kb, (u32)(b.kb >> 32));
}

struct MergeState {
    Words W; u32 n_words;
    u64 *pkey; i64 *pcnt; u64 pcap;
    Best *bmax; uint8_t *dirty; u32 n_blocks;
    const u32 *csr_off; const Rec *csr_rec;
    Rec *log; u64 *log_begin; u64 log_cap;
    // big log slices are bucket-sorted by hash(neighbour) into log2 right after the step that wrote them:
    // bk_lg[step] = log2(buckets) (0 = slice left unsorted), bk_start[step] = first of its buckets+1 offsets in bk_off
    Rec *log2; u32 *bk_lg; u64 *bk_start; u32 *bk_off; u64 bk_off_cap; u32 *bk_scratch;   // bk_scratch: 2 x (hist, cursor) x SORT_MAX_BK
    u32 *tok_off; u32 *tok_len; u64 *tok_key; uint8_t *tok_bytes; u64 tok_bytes_cap;
    struct BarSlot *bar; u32 *bar_ctr;            // gather slots / the two barrier counters (128 B apart) of k_merge_loop, zeroed before every launch
    // tail kernel (one thread-block cluster): per 64-block superblock a 64-bit mask of dirty blocks
    u64 *sdirty_mask; u32 n_super; int tail_mode;
    int stop_at;                                 // this launch runs steps [ctr[1], stop_at)
    int32_t *merges_out; i64 *merge_cnt_out; int n_merges;
    // [0]=log cursor [1]=n_done (next step) [2]=pair keys created [3]=status flags [4]=tok bytes cursor
    // [5]=keys popped since the table was last rebuilt [6]=previous winner key still to be popped (PAIR_EMPTY if none)
    // [8]=words rewritten (all steps) [9]=bk_off pool cursor [10]=slices sorted
    // [7]=sum over steps of the live pair-table keys (the reference's max() scans that many dict entries, train.py:187-189)
    u64 *ctr;
    // profile (ns / counts): [0]=phase1 [1]=sync1 [2]=apply [3]=sync2 on CTA 0; [4]=token CTA work; [5]=index records scanned;
    // [6]=words rewritten; [7]=steps; [9]=dirty blocks rescanned
    u64 *prof;
    u32 min_rec;      // smallest number of records dealt to a warp per pass of the apply phase (BPE_MERGE_MINREC)
    u64 *cta_prof;    // optional (profile builds): per step and CTA {start, arrive1, exit1, arrive2}
    u32 *step_prof;   // optional per-step trace: 4 x u32 per step (phase1+sync1 ns, apply+sync2 ns, records scanned, words rewritten so far)
};

// The loop state lives in constant memory: helpers are real (non-inlined) functions so that the persistent
// kernel stays small enough for the instruction caches -- every step runs each code path only once, so a
// large inlined kernel is instruction-fetch bound.
__constant__ MergeState cM;

#ifdef BPE_MERGE_PROFILE
__device__ __forceinline__ u64 gtime_ns() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// timer read that cannot be scheduled before `dep` is available
__device__ __forceinline__ u64 gtime_after(u64 dep) { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) : "l"(dep) : "memory"); return t; }
#define PROF_ADD(i, v) atomicAdd(&cM.prof[i], (u64)(v))
#else
__device__ __forceinline__ u64 gtime_ns() { return 0; }
__device__ __forceinline__ u64 gtime_after(u64) { return 0; }
#define PROF_ADD(i, v) do { } while (0)
#endif

__device__ __forceinline__ int bytes_cmp_dev(const uint8_t *x, u32 nx, const uint8_t *y, u32 ny) {
    u32 m = nx < ny ? nx : ny;
    for (u32 i = 0; i < m; i++) { if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1; }
    return nx < ny ? -1 : (nx > ny ? 1 : 0);
}
// python: bytes(p) > bytes(q) for two tokens whose 8-byte prefix keys are equal (rare slow path)
__device__ __noinline__ bool tok_greater_slow(u32 p, u32 q) {
    PROF_ADD(8, 1);
    u32 lp = cM.tok_len[p], lq = cM.tok_len[q];
    if (lp <= 8 && lq <= 8) return lp > lq;      // equal zero-padded prefixes: the shorter one is a prefix of the longer
    return bytes_cmp_dev(cM.tok_bytes + cM.tok_off[p], lp, cM.tok_bytes + cM.tok_off[q], lq) > 0;
}
// Tie-break of two candidates with equal counts: python's ((bytes_a, bytes_b)) tuple order.  A real function
// (not inlined): the persistent kernel must stay small, see the note at cM.
__device__ __forceinline__ bool best_tie_greater(u64 xkey, u64 xka, u64 xkb, u64 ykey, u64 yka, u64 ykb) {
    u32 xa = (u32)(xkey >> 32), ya = (u32)(ykey >> 32);
    if (xa != ya) return xka != yka ? xka > yka : tok_greater_slow(xa, ya);
    u32 xb = (u32)xkey, yb = (u32)ykey;
    if (xb == yb) return false;
    return xkb != ykb ? xkb > ykb : tok_greater_slow(xb, yb);
}
// (count, (bytes_a, bytes_b)) ordering of train.py:187-189
__device__ __forceinline__ bool best_greater(const Best &x, const Best &y) {
    if (x.cnt != y.cnt) return x.cnt > y.cnt;
    if (x.cnt == CNT_DEAD) return false;
    return best_tie_greater(x.key, x.ka, x.kb, y.key, y.ka, y.kb);
}
// butterfly reduction with the full comparison (handles every tie): the slow path of warp_best
__device__ __noinline__ Best warp_best_butterfly(Best b) {
#pragma unroll 1
    for (int d = 16; d; d >>= 1) {
        Best o;
        o.cnt = __shfl_xor_sync(0xffffffffu, b.cnt, d);
        o.key = __shfl_xor_sync(0xffffffffu, b.key, d);
        o.ka = __shfl_xor_sync(0xffffffffu, b.ka, d);
        o.kb = __shfl_xor_sync(0xffffffffu, b.kb, d);
        if (best_greater(o, b)) b = o;
    }
    return b;
}
// warp-wide maximum of v over the lanes in `active` (0 elsewhere), two redux.sync steps
__device__ __forceinline__ u64 warp_max_u64(u64 v, bool active) {
    const u32 hi = active ? (u32)(v >> 32) : 0u;
    const u32 mh = __reduce_max_sync(0xffffffffu, hi);
    const u32 lo = (active && hi == mh) ? (u32)v : 0u;
    const u32 ml = __reduce_max_sync(0xffffffffu, lo);
    return ((u64)mh << 32) | ml;
}
// Warp arg-max under the (count, (bytes_a, bytes_b)) order; every lane returns the winner.  The common cases -- a
// unique maximal count, or ties decided by the 8-byte prefix keys -- take a few redux / ballot steps instead of a
// five-round butterfly of 256-bit shuffles; anything subtler (equal prefixes of different tokens) falls back to it.
__device__ __forceinline__ Best warp_best(Best b) {
    // counts: CNT_DEAD (most negative) marks "no candidate"; bias to unsigned order
    const u64 bc = (u64)b.cnt ^ 0x8000000000000000ull;
    const u64 mc = warp_max_u64(bc, true);
    if (mc == 0) return BEST_NONE;               // every lane is empty
    u32 t = __ballot_sync(0xffffffffu, bc == mc);
    if (t & (t - 1)) {                           // tie on the count: first tokens by prefix key
        const bool in1 = (t >> lane_id()) & 1u;
        const u64 ma = warp_max_u64(b.ka, in1);
        const u32 t2 = __ballot_sync(0xffffffffu, in1 && b.ka == ma);
        if (t2 & (t2 - 1)) {
            const bool in2 = (t2 >> lane_id()) & 1u;
            const u32 a0 = __shfl_sync(0xffffffffu, (u32)(b.key >> 32), __ffs(t2) - 1);
            if (__ballot_sync(0xffffffffu, in2 && (u32)(b.key >> 32) != a0)) return warp_best_butterfly(b);   // equal prefix, different tokens
            const u64 mb = warp_max_u64(b.kb, in2);
            const u32 t3 = __ballot_sync(0xffffffffu, in2 && b.kb == mb);
            if (t3 & (t3 - 1)) {
                const u32 b0 = __shfl_sync(0xffffffffu, (u32)b.key, __ffs(t3) - 1);
                const bool in3 = (t3 >> lane_id()) & 1u;
                if (__ballot_sync(0xffffffffu, in3 && (u32)b.key != b0)) return warp_best_butterfly(b);
            }
            t = t3;
        } else t = t2;
    }
    const int src = __ffs(t) - 1;
    Best r;
    r.cnt = __shfl_sync(0xffffffffu, b.cnt, src); r.key = __shfl_sync(0xffffffffu, b.key, src);
    r.ka = __shfl_sync(0xffffffffu, b.ka, src); r.kb = __shfl_sync(0xffffffffu, b.kb, src);
    return r;
}

// A pair-table slot changed: its block must be rescanned before the next arg-max.
__device__ __forceinline__ void mark_dirty(u64 slot) {
    const u32 blk = (u32)(slot / PB);
    if (cM.tail_mode) atomicOr(&cM.sdirty_mask[blk >> 6], 1ull << (blk & 63u));   // fire-and-forget reduction
    else cM.dirty[blk] = 1;
}

// frequencies[key] += delta with defaultdict semantics (train.py:36,65-78): a missing key is created.
// The count update is a fire-and-forget reduction; a popped key that is touched again is repaired by
// the next rescan of its block (see rescan_block).  (s, k) = first probe slot and the key read there.
// kClaimed: the probe at slot s was the claiming CAS itself and k is what that CAS returned (only for keys that cannot
// have been in the table before this step); otherwise k is the key read at s.
template <bool kClaimed>
__device__ __forceinline__ void pair_add_probe(u64 key, i64 delta, u64 s, u64 k) {
    const u64 mask = cM.pcap - 1;
    for (u64 probes = 0; probes < cM.pcap; probes++) {
        if (!kClaimed && k == PAIR_EMPTY) k = atomicCAS(&cM.pkey[s], PAIR_EMPTY, key);
        if (k == PAIR_EMPTY) { atomicAdd(&cM.ctr[2], 1ull); k = key; }       // our CAS claimed the slot
        if (k == key) {
            atomicAdd((u64 *)&cM.pcnt[s], (u64)delta);
            mark_dirty(s);
            return;
        }
        s = (s + 1) & mask;
        k = kClaimed ? atomicCAS(&cM.pkey[s], PAIR_EMPTY, key) : cM.pkey[s];
    }
    cM.ctr[3] = 1;                                // table full
}
__device__ __noinline__ void pair_add_from(u64 key, i64 delta, u64 s, u64 k) { pair_add_probe<false>(key, delta, s, k); }
__device__ __forceinline__ void pair_add(u32 a, u32 b, i64 delta) {
    u64 key = ((u64)a << 32) | b;
    u64 s = pair_hash(key) & (cM.pcap - 1);
    pair_add_from(key, delta, s, cM.pkey[s]);
}

// Apply merge (a,b)->nw at the occurrence whose left symbol sits at position p (word frequency c): one merge site of the
// left-to-right scan of train.py:196-224 with update_frequencies_after_merge (52-78), merge_subwords (132-139) and
// create_new_token_indices (107-129).
//
// All sites of a step run concurrently, so every thread reasons about the state AT THE START OF THE STEP, which it can
// always reconstruct: a symbol equal to nw was `a`, a tombstone of this step was `b`.  The reference's sequential
// semantics in those terms:
//   * a record is a site iff its symbols are (a,b) and -- for a == b, inside a run of a's -- it sits at an even offset
//     from the start of the run (non-overlapping, left to right: aaaa -> [aa, aa], aaa -> [aa, a]);
//   * the left neighbour is the already merged token when the occurrence directly to the left is a site as well
//     (abab: the second site sees (nw, a), decrements it and creates (nw, nw)); the right neighbour is always the
//     unmerged symbol (the scan has not reached it yet);
//   * both new pairs are indexed, even if the right one is merged away later in the same step (stale, harmless).
__device__ __noinline__ void apply_site(u32 p, i64 c, u32 a, u32 b, u32 nw, int step) {
    int32_t *s = cM.W.sym;
    const int32_t dead_now = -(step + 2);
    const int32_t ia = (int32_t)a, ib = (int32_t)b, inw = (int32_t)nw;
#define ORIG(e) ((e) == inw ? ia : ((e) == dead_now ? ib : (e)))            /* value at the start of the step */
#define WAS_LIVE(e) ((e) >= 0 || (e) == dead_now)                          /* live at the start of the step */
#define OLD_TOMB(e) ((e) < SYM_SEP && (e) != dead_now)                     /* tombstone of an earlier step */
    // the four symbols around p are loaded together (one round trip in the common case of no old tombstones)
    int32_t vl = s[p - 1];
    const int32_t e0 = s[p];
    int32_t vb = s[p + 1], vr = s[p + 2];
    if (!WAS_LIVE(e0) || ORIG(e0) != ia) return;                           // stale record
    u32 pb = p + 1;
    while (OLD_TOMB(vb)) { pb++; vb = s[pb]; }
    if (!WAS_LIVE(vb) || ORIG(vb) != ib) return;                           // stale record
    // ---- left context ----
    bool has_l = false; u32 left = 0, pos_l = 0;
    u32 q = p - 1;
    while (OLD_TOMB(vl)) { q--; vl = s[q]; }
    if (a == b) {
        // run of a's ending just before p: its length decides whether p starts a site
        u32 n_left = 0, second = p;                                        // position of the second nearest a on the left
        while (WAS_LIVE(vl) && ORIG(vl) == ia) {
            n_left++;
            if (n_left == 2) second = q;
            q--; vl = s[q];
            while (OLD_TOMB(vl)) { q--; vl = s[q]; }
        }
        if (n_left & 1u) return;                                           // the a at p is the right half of the previous site
        if (n_left) { has_l = true; left = nw; pos_l = second; }
        else if (vl != SYM_SEP) { has_l = true; left = (u32)ORIG(vl); pos_l = q; }
    } else if (vl != SYM_SEP) {
        const int32_t o1 = ORIG(vl);
        has_l = true; left = (u32)o1; pos_l = q;
        if (o1 == ib) {                                                    // (.., a, b, [a, b]): the left occurrence is a site too
            u32 q2 = q - 1;
            int32_t v2 = s[q2];
            while (OLD_TOMB(v2)) { q2--; v2 = s[q2]; }
            if (WAS_LIVE(v2) && ORIG(v2) == ia) { left = nw; pos_l = q2; }
        }
    }
    // ---- right context: the symbol as it was (it is merged later, if at all) ----
    u32 qr = pb + 1;
    if (pb != p + 1) vr = s[qr];
    while (OLD_TOMB(vr)) { qr++; vr = s[qr]; }
    const bool has_r = vr != SYM_SEP;
    const u32 right = has_r ? (u32)ORIG(vr) : 0;
    // ---- the four dict updates of update_frequencies_after_merge.  (left, nw) and (nw, right) contain the token made in
    // this step, so they are not in the table unless another site of this step put them there: their first probe is the
    // claiming CAS itself, issued together with the read probes of the two old keys and the log allocation ----
    const u64 mask = cM.pcap - 1;
    const u64 k0 = ((u64)left << 32) | a, k1 = ((u64)left << 32) | nw, k2 = ((u64)b << 32) | right, k3 = ((u64)nw << 32) | right;
    const u64 s0 = pair_hash(k0) & mask, s1 = pair_hash(k1) & mask, s2 = pair_hash(k2) & mask, s3 = pair_hash(k3) & mask;
    u64 v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    if (has_l) { v0 = cM.pkey[s0]; v1 = atomicCAS(&cM.pkey[s1], PAIR_EMPTY, k1); }
    if (has_r) { v2 = cM.pkey[s2]; v3 = atomicCAS(&cM.pkey[s3], PAIR_EMPTY, k3); }
    // index records of the two new pairs: one atomic per group of threads that arrive here together
    const u32 n_rec = (u32)has_l + (u32)has_r;
    u64 li = 0;
    {
        const u32 act = __activemask();
        const u32 ml = __ballot_sync(act, has_l), mr = __ballot_sync(act, has_r);
        const u32 lt = (1u << lane_id()) - 1u;
        const u32 pre = __popc(ml & lt) + __popc(mr & lt), tot = __popc(ml) + __popc(mr);
        const int leader = __ffs(act) - 1;
        u64 base = 0;
        if ((int)lane_id() == leader && tot) base = atomicAdd(&cM.ctr[0], (u64)tot);
        li = __shfl_sync(act, base, leader) + pre;
    }
    const bool log_ok = li + n_rec <= cM.log_cap;
    if (!log_ok) cM.ctr[3] = 2;
    s[p] = inw; s[pb] = dead_now;                                          // merge_subwords: positions never move
    if (log_ok) {
        if (has_l) { store_rec(&cM.log[li], left, pos_l, c); li++; }        // (left, nw) at the position of `left`
        if (has_r) store_rec(&cM.log[li], 0x80000000u | right, p, c);       // (nw, right) at p
    }
    // The four probe sequences advance together: one loop whose trip count is the longest of the four (in the warp), with
    // the loads / CASes of an iteration in flight at the same time -- not four loops one after the other.
    // vj = what the probe of key j at slot sj returned; casj: that probe was a claiming CAS (PAIR_EMPTY = claimed by us).
    {
        u64 kk[4] = {k0, k1, k2, k3}, ss[4] = {s0, s1, s2, s3}, vv[4] = {v0, v1, v2, v3};
        const i64 dd[4] = {-c, c, -c, c};
        u32 cas = 0xAu;                                                    // keys 1 and 3 start with a claiming CAS
        u32 pend = (has_l ? 3u : 0u) | (has_r ? 12u : 0u);
        for (u64 probes = 0; pend && probes <= cM.pcap; probes++) {
#pragma unroll
            for (u32 j = 0; j < 4; j++) {
                if (!((pend >> j) & 1u)) continue;
                u64 k = vv[j];
                if (k == PAIR_EMPTY) {
                    if (!((cas >> j) & 1u)) {                              // empty slot seen by a read probe: claim it
                        vv[j] = atomicCAS(&cM.pkey[ss[j]], PAIR_EMPTY, kk[j]);
                        cas |= 1u << j;
                        continue;
                    }
                    atomicAdd(&cM.ctr[2], 1ull);                           // our CAS claimed the slot
                    k = kk[j];
                }
                if (k == kk[j]) {
                    atomicAdd((u64 *)&cM.pcnt[ss[j]], (u64)dd[j]);
                    mark_dirty(ss[j]);
                    pend &= ~(1u << j);
                    continue;
                }
                ss[j] = (ss[j] + 1) & mask;
                vv[j] = ((cas >> j) & 1u) ? atomicCAS(&cM.pkey[ss[j]], PAIR_EMPTY, kk[j]) : cM.pkey[ss[j]];
            }
        }
        if (pend) cM.ctr[3] = 1;                                           // table full
    }
    PROF_ADD(6, 1);
#undef ORIG
#undef WAS_LIVE
#undef OLD_TOMB
}

// Rescan the PB slots of one block with a full warp.  Pops `prev_key` when it meets it and repairs
// popped keys that were touched again (count = CNT_DEAD + deltas  ->  deltas).
__device__ __forceinline__ Best rescan_block(u32 blk, u64 prev_key) {
    Best bst = BEST_NONE;
    const u64 sbase = (u64)blk * PB;
    u64 keys[PB / 32]; i64 cnts[PB / 32]; u64 kas[PB / 32], kbs[PB / 32];
#pragma unroll
    for (u32 k = 0; k < PB / 32; k++) { keys[k] = cM.pkey[sbase + k * 32 + lane_id()]; cnts[k] = cM.pcnt[sbase + k * 32 + lane_id()]; }
#pragma unroll
    for (u32 k = 0; k < PB / 32; k++) {          // second round trip: prefix keys of both tokens of every live slot
        bool live = keys[k] != PAIR_EMPTY;
        kas[k] = live ? cM.tok_key[(u32)(keys[k] >> 32)] : 0;
        kbs[k] = live ? cM.tok_key[(u32)keys[k]] : 0;
    }
#pragma unroll
    for (u32 k = 0; k < PB / 32; k++) {
        Best c;
        c.key = keys[k]; c.cnt = cnts[k]; c.ka = kas[k]; c.kb = kbs[k];
        if (c.key == PAIR_EMPTY) continue;
        if (c.key == prev_key) { cM.pcnt[sbase + k * 32 + lane_id()] = CNT_DEAD; continue; }
        if (c.cnt == CNT_DEAD) continue;
        if (c.cnt < CNT_DEAD_LIMIT) { c.cnt = (i64)((u64)c.cnt - (u64)CNT_DEAD); cM.pcnt[sbase + k * 32 + lane_id()] = c.cnt; }
        if (best_greater(c, bst)) bst = c;
    }
    return warp_best(bst);
}


// ---- token bookkeeping: bytes and prefix key of the new token, outputs (one CTA, all its threads) -------------
__device__ __forceinline__ void token_bookkeeping(int step, const Best &win, u32 a, u32 b, u32 nw, bool prof_thread, u64 t2) {
    const u32 tid = threadIdx.x;
    u64 tA = prof_thread ? gtime_ns() : 0;
    const u32 la = cM.tok_len[a], lb = cM.tok_len[b], ln = la + lb;
    const u32 oa = cM.tok_off[a], ob = cM.tok_off[b];
    const u64 cur = cM.ctr[4];
    const u64 live = cM.ctr[2];
    __syncthreads();
    u64 tB = prof_thread ? gtime_ns() : 0;
    if (cur + ln > cM.tok_bytes_cap) { if (tid == 0) cM.ctr[3] = 4; }
    else {
        uint8_t *dst = cM.tok_bytes + cur;
        const uint8_t *pa = cM.tok_bytes + oa, *pb = cM.tok_bytes + ob;
        for (u32 i = tid; i < ln; i += MG_NT) dst[i] = i < la ? pa[i] : pb[i - la];
        if (tid == 0) {
            // first 8 bytes of a+b from the operands' zero-padded big-endian prefix keys
            u64 nkey = cM.tok_key[a];
            if (la < 8) nkey |= cM.tok_key[b] >> (8 * la);
            cM.tok_off[nw] = (u32)cur; cM.tok_len[nw] = ln; cM.tok_key[nw] = nkey; cM.ctr[4] = cur + ln;
            cM.merges_out[2 * step] = (int32_t)a; cM.merges_out[2 * step + 1] = (int32_t)b;
            cM.merge_cnt_out[step] = win.cnt;
            cM.ctr[1] = (u64)(step + 1);
            cM.ctr[7] += live - cM.ctr[5]; cM.ctr[5] += 1;
            if (prof_thread) { cM.prof[10] += tA - t2; cM.prof[11] += tB - tA; cM.prof[12] += gtime_ns() - tB; }
        }
    }
    // make sure the winner's block is rescanned so the key gets popped
    if (tid == 32) {
        u64 mask = cM.pcap - 1, s = pair_hash(win.key) & mask;
        while (cM.pkey[s] != win.key) s = (s + 1) & mask;
        mark_dirty(s);
    }
}

// source array of an index range: 0 = log, 1 = bucket-sorted copy, 2 = CSR of the initial byte pairs
#define T_SRC(code) ((code) == 2 ? cM.csr_rec : ((code) == 1 ? (const Rec *)cM.log2 : (const Rec *)cM.log))
#define SORT_MIN 2048u                           // slices with fewer records are scanned whole
#define SORT_MAX_LG 12u                          // at most 4096 buckets
#define SORT_MAX_BK (1u << SORT_MAX_LG)
__device__ __forceinline__ u32 bucket_of(u32 x, u32 lg) { return (x * 0x9E3779B1u) >> (32u - lg); }

// Index range of the winner (a,b): token_indices[best_pair] of train.py:192.  out[0..1] = record range, out[2] = 1 when
// the range lies in the bucket-sorted copy (log2).  One thread per CTA.
__device__ __forceinline__ void winner_range(u64 key, u64 *out) {
    const u32 wa = (u32)(key >> 32), wb = (u32)key, wT = wa > wb ? wa : wb;
    if (wT < 256) { u32 pp = (wa << 8) | wb; out[0] = cM.csr_off[pp]; out[1] = cM.csr_off[pp + 1]; out[2] = 2; return; }
    const u32 t = wT - 256;
    const u64 lo = cM.log_begin[t], hi = cM.log_begin[t + 1];
    const u32 lg = cM.bk_lg[t];
    const u64 st0 = cM.bk_start[t];              // (garbage when the slice is unsorted; loaded alongside, not after)
    if (!lg) { out[0] = lo; out[1] = hi; out[2] = 0; return; }
    const u32 want = wb >= wa ? wa : (0x80000000u | wb);
    const u64 st = st0 + bucket_of(want, lg);
    out[0] = lo + cM.bk_off[st]; out[1] = lo + cM.bk_off[st + 1]; out[2] = 1;
}

// Bucket-sort the slice [lo, lo + n) of the log by hash(record.x) into log2 (same offsets) and publish the bucket
// offsets.  Called by every thread of every CTA between two steps; `sync` is the grid / cluster barrier.
// s_scan: SORT_MAX_BK u32 of shared memory.  parity alternates so that the scratch of the previous sort can be cleared here.
template <typename SyncFn>
__device__ __forceinline__ void sort_slice(int step, u64 lo, u32 n, u32 parity, bool leader_cta, u64 gthread, u64 gstride,
                                           u32 *s_scan, SyncFn sync) {
    u32 lg = 4;
    while ((64u << lg) < n && lg < SORT_MAX_LG) lg++;
    const u32 nbk = 1u << lg;
    u32 *hist = cM.bk_scratch + (size_t)parity * 2 * SORT_MAX_BK, *cur = hist + SORT_MAX_BK;
    u32 *other = cM.bk_scratch + (size_t)(parity ^ 1u) * 2 * SORT_MAX_BK;
    for (u64 i = gthread; i < n; i += gstride) atomicAdd(&hist[bucket_of(cM.log[lo + i].x, lg)], 1u);
    for (u64 i = gthread; i < 2 * SORT_MAX_BK; i += gstride) other[i] = 0;       // last used two barriers ago
    sync();
    // exclusive scan of the histogram, redundantly in every CTA
    const u32 tid = threadIdx.x;
    const u32 per = (nbk + MG_NT - 1) / MG_NT;   // <= 8
    u32 v[8], sum = 0;
#pragma unroll
    for (u32 k = 0; k < 8; k++) { u32 i = tid * per + k; v[k] = (k < per && i < nbk) ? hist[i] : 0; sum += v[k]; }
    __shared__ u32 s_wsum[MG_NT / 32 + 1];
    u32 inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, inc, d); if (lane_id() >= (u32)d) inc += t; }
    if (lane_id() == 31) s_wsum[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        u32 x = tid < MG_NT / 32 ? s_wsum[tid] : 0, xi = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, xi, d); if (tid >= (u32)d) xi += t; }
        s_wsum[tid] = xi - x;
    }
    __syncthreads();
    u32 base = inc - sum + s_wsum[tid >> 5];
    u64 pool = 0;
    if (leader_cta) pool = cM.ctr[9];
#pragma unroll
    for (u32 k = 0; k < 8; k++) {
        u32 i = tid * per + k;
        if (k < per && i < nbk) { s_scan[i] = base; if (leader_cta) cM.bk_off[pool + i] = base; base += v[k]; }
    }
    if (leader_cta && tid == 0) {
        cM.bk_off[pool + nbk] = n; cM.bk_start[step] = pool; cM.bk_lg[step] = lg; cM.ctr[9] = pool + nbk + 1; cM.ctr[10] += 1;
    }
    __syncthreads();
    for (u64 i = gthread; i < n; i += gstride) {
        const uint4 r = *reinterpret_cast<const uint4 *>(&cM.log[lo + i]);
        const u32 bk = bucket_of(r.x, lg);
        *reinterpret_cast<uint4 *>(&cM.log2[lo + s_scan[bk] + atomicAdd(&cur[bk], 1u)]) = r;
    }
    sync();
}

// ---- apply the merge at every occurrence indexed under (a,b): index slice [lo, hi) spread over `gstride` threads ----
__device__ __forceinline__ void apply_winner(int step, u32 a, u32 b, u32 nw, u64 lo, u64 hi, const Rec *__restrict__ src, u64 gthread, u64 gstride) {
    const u32 T = a > b ? a : b;
    if (gthread == 0) PROF_ADD(5, hi - lo);
    // pairs of two initial bytes: the CSR slice holds exactly the occurrences of (a,b); otherwise filter by neighbour
    const bool filter = T >= 256;
    const u32 want = b >= a ? a : (0x80000000u | b);
    for (u64 i = lo + gthread; i < hi; i += 4 * gstride) {               // 4 records in flight per thread
        Rec r0 = load_rec(&src[i]), r1, r2, r3;
        const bool h1 = i + gstride < hi, h2 = i + 2 * gstride < hi, h3 = i + 3 * gstride < hi;
        if (h1) r1 = load_rec(&src[i + gstride]);
        if (h2) r2 = load_rec(&src[i + 2 * gstride]);
        if (h3) r3 = load_rec(&src[i + 3 * gstride]);
        if (!filter || r0.x == want) apply_site(r0.pos, r0.cnt, a, b, nw, step);
        if (h1 && (!filter || r1.x == want)) apply_site(r1.pos, r1.cnt, a, b, nw, step);
        if (h2 && (!filter || r2.x == want)) apply_site(r2.pos, r2.cnt, a, b, nw, step);
        if (h3 && (!filter || r3.x == want)) apply_site(r3.pos, r3.cnt, a, b, nw, step);
    }
}

// ---- grid-wide synchronisation of k_merge_loop (all CTAs are co-resident: cooperative launch) ------------------------
// Counter barrier: every CTA adds 1 to one global counter with release semantics and ONE lane polls that word with acquire
// loads until it reaches (barriers so far) x (CTAs).  Measured against the alternatives at 11 GB (same box): per-CTA flags
// polled by a warp 514 ms (five acquire loads per lane per poll, 148 pollers), cooperative_groups::grid.sync 581 ms,
// fence + relaxed stores / polls 580 ms, this 481 ms.
//  * grid_barrier: the barrier alone.
//  * grid_gather:  barrier + all-gather of the per-CTA arg-max candidates: a CTA stores its candidate in its slot before it
//    arrives; after the barrier warp 0 loads all slots and reduces them, so every CTA derives the same winner.  (An LL-style
//    slot -- every 64-bit word tagged with the epoch, no barrier-then-data round trip -- was slower: 649 against 541 ms.)
// Slot of a CTA: w[0..3] candidate, w[6..7] index range of that candidate (epoch-tagged words, written during the gather).
struct __align__(64) BarSlot { u64 w[8]; };
#define MG_MAX_CTAS 160u
__device__ __forceinline__ void st_relaxed_v2(u64 *p, u64 a, u64 b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ void ld_relaxed_v2(const u64 *p, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p) { u32 v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_add_u32(u32 *p, u32 v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// warp 0 only; the CTA barriers around it extend the ordering to the other warps
__device__ __forceinline__ void counter_arrive_wait(u32 *counter, u32 target) {
    if (threadIdx.x == 0) {
        red_release_add_u32(counter, 1u);
        while ((int)(ld_acquire_u32(counter) - target) < 0) { }
    }
    __syncwarp();
}
// Called by all threads of the CTA.
__device__ __forceinline__ void grid_barrier(u32 *counter, u32 G, u32 epoch, u64 *tp = nullptr) {
    __syncthreads();
#ifdef BPE_MERGE_PROFILE
    if (tp && threadIdx.x == 0) tp[0] = gtime_ns();
#endif
    if (threadIdx.x < 32) counter_arrive_wait(counter, epoch * G);
    __syncthreads();
}

// While warp 0 waits in the gather, warp 1 looks up the index range of this CTA's OWN candidate and publishes it in the
// spare words of its slot, tagged with the epoch (each 64-bit word carries a tag, so a torn or stale read is detected and
// the reader falls back to winner_range); it also pulls the first records of that range and the symbols they point at
// into L2.  Whichever candidate wins, the apply phase then starts one to three dependent round trips later in the chain.
#define RANGE_TAG0(e) ((u64)((e) & 0xFFFFFFu) << 40)
#define RANGE_TAG1(e) ((u64)((e) & 0x3FFFFFu) << 42)
__device__ __forceinline__ void gather_side_work(BarSlot *slots, u32 epoch, const Best &cand, int step) {
    if (cand.cnt == CNT_DEAD || epoch >= (1u << 22)) return;   // (the tags are 22 bits: no publishing beyond 4 M steps of one launch)
    const u32 lane = lane_id();
    const u32 a = (u32)(cand.key >> 32), b = (u32)cand.key, T = a > b ? a : b;
    if (T >= 256 && (int)(T - 256) + 1 >= step) return;      // log_begin[step] is being written by CTA 0 right now: not visible yet
    u64 r[3] = {0, 0, 0};
    if (lane == 0) {
        winner_range(cand.key, r);
        if (r[0] < (1ull << 40) && r[1] < (1ull << 40))
            st_relaxed_v2(&slots[blockIdx.x].w[6], r[0] | RANGE_TAG0(epoch), r[1] | (r[2] << 40) | RANGE_TAG1(epoch));
    }
    const u64 lo = __shfl_sync(0xffffffffu, r[0], 0), hi = __shfl_sync(0xffffffffu, r[1], 0);
    const u32 code = (u32)__shfl_sync(0xffffffffu, r[2], 0);
    const Rec *src = T_SRC(code);
    const bool filter = T >= 256;
    const u32 want = b >= a ? a : (0x80000000u | b);
    const u64 i = lo + lane;
    if (i < hi) {
        const Rec rec = load_rec(&src[i]);
        if (!filter || rec.x == want) asm volatile("prefetch.global.L2 [%0];" ::"l"(&cM.W.sym[rec.pos]));
    }
}

// Called by all threads of the CTA; `mine` is read from lane 0 of warp 0, *s_cand is the same candidate in shared memory
// (for warp 1).  Returns the maximum of all CTAs' candidates in every lane of warp 0 (other warps: BEST_NONE) and, in `sup`,
// the index of the CTA that supplied it.
__device__ __forceinline__ Best grid_gather(BarSlot *slots, u32 *counter, u32 G, u32 epoch, const Best &mine, u32 &sup,
                                            const Best *s_cand, int step, u64 *tp = nullptr) {
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) gather_side_work(slots, epoch, *s_cand, step);
#ifdef BPE_MERGE_PROFILE
    if (tp && threadIdx.x == 0) tp[0] = gtime_ns();
#endif
    Best c = BEST_NONE;
    if (threadIdx.x < 32) {
        const u32 lane = threadIdx.x;
        if (lane == 0) store_best(reinterpret_cast<Best *>(slots[blockIdx.x].w), mine);
        counter_arrive_wait(counter, epoch * G);
        Best o5[MG_MAX_CTAS / 32];
#pragma unroll
        for (u32 k = 0; k < MG_MAX_CTAS / 32; k++) { const u32 i = lane + 32 * k; o5[k] = i < G ? load_best(reinterpret_cast<const Best *>(slots[i].w)) : BEST_NONE; }
#pragma unroll
        for (u32 k = 0; k < MG_MAX_CTAS / 32; k++) if (o5[k].cnt != CNT_DEAD && best_greater(o5[k], c)) c = o5[k];
        c = warp_best(c);
        u32 who = 0xFFFFFFFFu;                   // a key lives in one block, so exactly one CTA supplied the winner
#pragma unroll
        for (u32 k = 0; k < MG_MAX_CTAS / 32; k++) if (lane + 32 * k < G && o5[k].cnt != CNT_DEAD && o5[k].key == c.key) who = lane + 32 * k;
        const u32 m = __ballot_sync(0xffffffffu, who != 0xFFFFFFFFu);
        sup = m ? __shfl_sync(0xffffffffu, who, __ffs(m) - 1) : 0xFFFFFFFFu;
    }
    __syncthreads();
#ifdef BPE_MERGE_PROFILE
    if (tp && threadIdx.x == 0) tp[1] = gtime_ns();
#endif
    return c;
}

// ---- shared-memory copy of the cached block maxima a CTA owns (structure of arrays: conflict-free 8-byte accesses) ----
#define MG_CACHE_ITERS 5u                        // chunks (of 64 blocks) per warp kept in shared memory; more are read from bmax
#define MG_CACHE_N (MG_CACHE_ITERS * (MG_NT / 32) * 64u)
#define MG_LIST_CAP 4096u                        // dirty blocks a CTA lists per step; the overflow is rescanned by the owning warp
#define MG_DYN_SMEM ((size_t)MG_CACHE_N * 32 + (size_t)MG_LIST_CAP * 4)
struct BmaxCache { i64 *cnt; u64 *key, *ka, *kb; };
__device__ __forceinline__ BmaxCache bmax_cache(unsigned char *base) {
    BmaxCache c;
    c.cnt = reinterpret_cast<i64 *>(base); c.key = reinterpret_cast<u64 *>(base) + MG_CACHE_N;
    c.ka = c.key + MG_CACHE_N; c.kb = c.ka + MG_CACHE_N;
    return c;
}
__device__ __forceinline__ void cache_store(const BmaxCache &c, u32 i, const Best &b) { c.cnt[i] = b.cnt; c.key[i] = b.key; c.ka[i] = b.ka; c.kb[i] = b.kb; }
__device__ __forceinline__ void cache_consider(const BmaxCache &c, u32 i, Best &mine) {
    const i64 n = c.cnt[i];
    if (n == CNT_DEAD || n < mine.cnt) return;
    Best o; o.cnt = n; o.key = c.key[i]; o.ka = c.ka[i]; o.kb = c.kb[i];
    if (best_greater(o, mine)) mine = o;
}

__global__ void __launch_bounds__(MG_NT) k_merge_loop() {
    __shared__ Best s_best[MG_NT / 32];
    __shared__ Best s_win, s_cand;
    __shared__ u64 s_status[2];
    __shared__ u64 s_range[4];
    __shared__ u32 s_scan[SORT_MAX_BK];
    __shared__ u32 s_nd;                         // dirty blocks listed in this step
    extern __shared__ __align__(16) unsigned char mg_smem[];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u32 G = gridDim.x;
    u32 n_sorts = 0;
    u32 epoch = 0, gepoch = 0;                   // barriers / gathers passed in this launch (the slots start at 0)
    const u32 warps_per_cta = MG_NT / 32;
    const u32 gwarp = blockIdx.x * warps_per_cta + warp, total_warps = G * warps_per_cta;
    const bool token_cta = blockIdx.x == G - 1;
    const u32 apply_ctas = G - 1;
    u64 prev_key = cM.ctr[6];                     // winner of the previous step: popped lazily during the rescan
    const int first_step = (int)cM.ctr[1];
    u32 n_tok = 256 + (u32)first_step;
    // The cached block maxima of the chunks this CTA owns stay in shared memory for the whole launch (chunk (it, warp) =
    // 64 blocks starting at (it * total_warps + gwarp) * 64); bmax in global memory is written through for the next launch.
    const BmaxCache sc = bmax_cache(mg_smem);
    if (tid == 0) s_nd = 0;
    u32 *s_list = reinterpret_cast<u32 *>(mg_smem + (size_t)MG_CACHE_N * 32);
    const u32 n_iter = (cM.n_blocks + total_warps * 64 - 1) / (total_warps * 64);
    for (u32 it = 0; it < n_iter && it < MG_CACHE_ITERS; it++) {
        const u32 base = (it * total_warps + gwarp) * 64, i0 = (it * warps_per_cta + warp) * 64 + lane;
        cache_store(sc, i0, base + lane < cM.n_blocks ? load_best(&cM.bmax[base + lane]) : BEST_NONE);
        cache_store(sc, i0 + 32, base + 32 + lane < cM.n_blocks ? load_best(&cM.bmax[base + 32 + lane]) : BEST_NONE);
    }
    __syncthreads();

    for (int step = first_step; step < cM.stop_at; step++) {
        // status flags are written before the barrier that ends a step and read here (uniform across the grid); the loads
        // overlap with the dirty-flag loads of phase 1A and are checked behind its CTA barrier
        u64 st_err = 0, st_keys = 0;
        if (tid == 0) { st_err = *((volatile u64 *)&cM.ctr[3]); st_keys = *((volatile u64 *)&cM.ctr[2]); }
        // ---- phase 1: rescan dirty blocks, reduce cached block maxima to one candidate per CTA ----
        if (blockIdx.x == 0 && tid == 0) cM.log_begin[step] = cM.ctr[0];
#ifdef BPE_MERGE_PROFILE
        const bool prof_thread = tid == 0 && (blockIdx.x == 0 || token_cta);
#else
        const bool prof_thread = false;
#endif
        u64 t0 = prof_thread ? gtime_ns() : 0;
#ifdef BPE_MERGE_PROFILE
        u64 *ctp = cM.cta_prof ? cM.cta_prof + ((size_t)step * G + blockIdx.x) * 4 : nullptr;
        if (ctp && tid == 0) ctp[0] = gtime_ns();
#else
        u64 *ctp = nullptr;
#endif
        Best mine = BEST_NONE;
        // A: collect the dirty blocks of this CTA's chunks in a shared list (any warp of the CTA may rescan them)
        for (u32 it = 0; it < n_iter; it++) {
            const u32 base = (it * total_warps + gwarp) * 64;
            const u32 b0 = base + lane, b1 = base + 32 + lane;
            const bool d0 = b0 < cM.n_blocks && cM.dirty[b0], d1 = b1 < cM.n_blocks && cM.dirty[b1];
            const u32 dlo = __ballot_sync(0xffffffffu, d0), dhi = __ballot_sync(0xffffffffu, d1);
            if (dlo | dhi) {
                const u32 nlo = __popc(dlo), n = nlo + __popc(dhi);
                u32 pos = 0;
                if (lane == 0) pos = atomicAdd(&s_nd, n);
                pos = __shfl_sync(0xffffffffu, pos, 0);
                const u32 lt = (1u << lane) - 1u;
                const u32 e0 = pos + __popc(dlo & lt), e1 = pos + nlo + __popc(dhi & lt);
                if (d0 && e0 < MG_LIST_CAP) s_list[e0] = b0;
                if (d1 && e1 < MG_LIST_CAP) s_list[e1] = b1;
                if (pos + n > MG_LIST_CAP) {     // list full (huge tables, first step): this warp rescans the rest itself
                    u64 dm = (u64)__ballot_sync(0xffffffffu, d0 && e0 >= MG_LIST_CAP) | ((u64)__ballot_sync(0xffffffffu, d1 && e1 >= MG_LIST_CAP) << 32);
                    while (dm) {
                        const u32 l2 = __ffsll((long long)dm) - 1; dm &= dm - 1;
                        const Best bst = rescan_block(base + l2, prev_key);
                        if (lane == 0) { store_best(&cM.bmax[base + l2], bst); cM.dirty[base + l2] = 0; if (it < MG_CACHE_ITERS) cache_store(sc, (it * warps_per_cta + warp) * 64 + l2, bst); }
                    }
                }
            }
        }
        if (tid == 0) { s_status[0] = st_err; s_status[1] = st_keys; }
        __syncthreads();
        if (s_status[0]) break;
        if (s_status[1] * 2 > cM.pcap) {         // table over half full: hand back to the host to grow it
            grid_barrier(cM.bar_ctr, G, ++epoch);
            if (blockIdx.x == 0 && tid == 0) { cM.ctr[6] = prev_key; cM.ctr[3] = MG_NEED_GROW; }
            return;
        }
        // B: the CTA's warps share the rescans evenly
        {
            const u32 nd = s_nd < MG_LIST_CAP ? s_nd : MG_LIST_CAP;
            for (u32 e = warp; e < nd; e += warps_per_cta) {
                const u32 blk = s_list[e];
                const Best bst = rescan_block(blk, prev_key);
                if (lane == 0) {
                    store_best(&cM.bmax[blk], bst); cM.dirty[blk] = 0; PROF_ADD(9, 1);
                    const u32 chunk = blk >> 6, it = chunk / total_warps, w = chunk - it * total_warps - blockIdx.x * warps_per_cta;
                    if (it < MG_CACHE_ITERS) cache_store(sc, (it * warps_per_cta + w) * 64 + (blk & 63u), bst);
                }
            }
        }
        __syncthreads();
        if (tid == 0) s_nd = 0;                  // (next written in phase A of the next step)
        // C: maximum of the cached block maxima this warp owns (shared memory; global for chunks beyond the cache)
        for (u32 it = 0; it < n_iter; it++) {
            if (it < MG_CACHE_ITERS) {
                const u32 i0 = (it * warps_per_cta + warp) * 64 + lane;
                cache_consider(sc, i0, mine);
                cache_consider(sc, i0 + 32, mine);
            } else {
                const u32 base = (it * total_warps + gwarp) * 64;
                const u32 b0 = base + lane, b1 = base + 32 + lane;
                Best m0 = BEST_NONE, m1 = BEST_NONE;
                if (b0 < cM.n_blocks) m0 = load_best(&cM.bmax[b0]);
                if (b1 < cM.n_blocks) m1 = load_best(&cM.bmax[b1]);
                if (m0.cnt != CNT_DEAD && best_greater(m0, mine)) mine = m0;
                if (m1.cnt != CNT_DEAD && best_greater(m1, mine)) mine = m1;
            }
        }
        mine = warp_best(mine);
        if (lane == 0) s_best[warp] = mine;
        __syncthreads();
        Best cta_cand = BEST_NONE;
        if (warp == 0) {
            Best c = lane < warps_per_cta ? s_best[lane] : BEST_NONE;
            cta_cand = warp_best(c);
        }
        u64 t1 = prof_thread ? gtime_ns() : 0;
        // ---- barrier + all-gather of the CTA candidates: every CTA derives the same winner ----
        u32 sup = 0xFFFFFFFFu;
        if (warp == 0 && lane == 0) s_cand = cta_cand;
        ++gepoch;
        Best gw = grid_gather(cM.bar, cM.bar_ctr + 32, G, gepoch, cta_cand, sup, &s_cand, step, ctp ? ctp + 1 : nullptr);
        u64 t2 = prof_thread ? gtime_ns() : 0;
        if (warp == 0 && lane == 0) {
            s_win = gw;
            if (gw.cnt != CNT_DEAD) {            // index range of the winner, read once per CTA: published by its supplier, or looked up
                bool have = false;
                if (sup != 0xFFFFFFFFu && gepoch < (1u << 22)) {
                    u64 w6, w7;
                    ld_relaxed_v2(&cM.bar[sup].w[6], w6, w7);
                    if ((w6 & (0xFFFFFFull << 40)) == RANGE_TAG0(gepoch) && (w7 & (0x3FFFFFull << 42)) == RANGE_TAG1(gepoch)) {
                        s_range[0] = w6 & ((1ull << 40) - 1); s_range[1] = w7 & ((1ull << 40) - 1); s_range[2] = (w7 >> 40) & 3u;
                        have = true;
                    }
                }
                if (!have) winner_range(gw.key, s_range);
            }
        }
        __syncthreads();
        const Best win = s_win;
        if (win.cnt == CNT_DEAD) break;          // len(byte_pair_frequencies) == 0 (train.py:184-185)
        const u32 a = (u32)(win.key >> 32), b = (u32)win.key;
        const u32 nw = n_tok;                    // symbol id of new_byte = a + b (train.py:190)
        const u64 r_lo = s_range[0], r_hi = s_range[1];

        if (token_cta) token_bookkeeping(step, win, a, b, nw, prof_thread, t2);
        // Records are dealt to the warps of all apply CTAs R at a time, R as small as one pass over the slice allows (but not
        // below cM.min_rec): a step with a few hundred occurrences runs a few lanes on every SM instead of sixteen full warps on
        // one SM, with less divergence between the sites that share a warp.
        else {
            const u64 n_rec = r_hi - r_lo, aw = (u64)apply_ctas * warps_per_cta;
            u32 R = 32;
            while (R > cM.min_rec && n_rec <= aw * (R >> 1)) R >>= 1;
            if (lane < R) apply_winner(step, a, b, nw, r_lo, r_hi, T_SRC(s_range[2]), ((u64)warp * apply_ctas + blockIdx.x) * R + lane, aw * R);
        }
        prev_key = win.key;
        n_tok++;
        u64 t3 = prof_thread ? gtime_ns() : 0;
        u64 tp2[2];
        grid_barrier(cM.bar_ctr, G, ++epoch, ctp ? tp2 : nullptr);
        if (ctp && tid == 0) ctp[3] = tp2[0];
        if ((r_hi - r_lo) * 2 >= SORT_MIN) {     // heuristic: about two records per index entry visited
            const u64 lb = *((volatile u64 *)&cM.log_begin[step]), cur = *((volatile u64 *)&cM.ctr[0]);
            const u64 pool = *((volatile u64 *)&cM.ctr[9]);
            if (cur - lb >= SORT_MIN && cur <= cM.log_cap && pool + SORT_MAX_BK + 1 <= cM.bk_off_cap) {
                sort_slice(step, lb, (u32)(cur - lb), n_sorts & 1u, blockIdx.x == 0, (u64)blockIdx.x * MG_NT + tid, (u64)G * MG_NT, s_scan,
                           [&]() { grid_barrier(cM.bar_ctr, G, ++epoch); });
                n_sorts++;
            }
        }
        if (prof_thread) {
            u64 t4 = gtime_ns();
            if (blockIdx.x == 0) {
                cM.prof[0] += t1 - t0; cM.prof[1] += t2 - t1; cM.prof[2] += t3 - t2; cM.prof[3] += t4 - t3; cM.prof[7] += 1;
                if (cM.step_prof) {
                    cM.step_prof[4 * step] = (u32)(t2 - t0); cM.step_prof[4 * step + 1] = (u32)(t4 - t2);
                    cM.step_prof[4 * step + 2] = (u32)cM.prof[5]; cM.step_prof[4 * step + 3] = (u32)cM.prof[6];
                }
            }
            if (token_cta) cM.prof[4] += t3 - t2;
        }
    }
    if (blockIdx.x == 0 && tid == 0) cM.ctr[6] = prev_key;
}


// =============================================================================================
// Tail kernel: the same loop run by ONE thread-block cluster.
//
// After the first ~thousand merges a step touches a handful of words, and its cost in k_merge_loop is the latency
// of two grid-wide barriers plus the L2 round trips of the cross-CTA arg-max.  Inside one cluster the barrier is the
// hardware cluster barrier (0.27 us for 16 CTAs against 1.2 us for the grid barrier, tools/bench_gridsync.cu), the
// per-superblock maxima live in shared memory and the per-CTA candidates travel through distributed shared memory.
//   level 0  pair-table slots
//   level 1  bmax[block]          global, PB slots per block
//   level 2  s_smax[superblock]   shared memory of the CTA that owns the superblock (64 blocks), persistent across steps
// Updates set a bit in sdirty_mask[superblock] (fire-and-forget atomicOr); a step rescans only those blocks.
// =============================================================================================
#define TAIL_MAX_CTAS 16u
#define MG_STOP_ERROR 1u
#define MG_STOP_GROW 2u

__global__ void __launch_bounds__(256) k_tail_prepare() {
    // dirty[] bytes of k_merge_loop -> per-superblock bit masks
    for (u32 sb = blockIdx.x * blockDim.x + threadIdx.x; sb < cM.n_super; sb += gridDim.x * blockDim.x) {
        u64 m = 0;
        for (u32 j = 0; j < 64; j++) { u32 b = sb * 64 + j; if (b < cM.n_blocks && cM.dirty[b]) { m |= 1ull << j; cM.dirty[b] = 0; } }
        cM.sdirty_mask[sb] = m;
    }
}

__device__ __forceinline__ Best load_best_cg(const Best *p) {      // from L2: the entry may just have been written by another warp
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    uint4 lo = __ldcg(q), hi = __ldcg(q + 1);
    Best b;
    b.cnt = (i64)(((u64)lo.y << 32) | lo.x); b.key = ((u64)lo.w << 32) | lo.z;
    b.ka = ((u64)hi.y << 32) | hi.x; b.kb = ((u64)hi.w << 32) | hi.z;
    return b;
}
__device__ __forceinline__ u32 nth_set_bit64(u64 m, u32 n) {       // position of the n-th (0-based) set bit
    for (u32 i = 0; i < n; i++) m &= m - 1;
    return __ffsll((long long)m) - 1;
}
// maximum of superblock sb from its 64 cached block maxima
__device__ __forceinline__ Best superblock_max(u32 sb, u32 lane) {
    const u32 b0 = sb * 64 + lane, b1 = b0 + 32;
    Best c = BEST_NONE;
    if (b0 < cM.n_blocks) { Best m = load_best_cg(&cM.bmax[b0]); if (m.cnt != CNT_DEAD) c = m; }
    if (b1 < cM.n_blocks) { Best m = load_best_cg(&cM.bmax[b1]); if (m.cnt != CNT_DEAD && best_greater(m, c)) c = m; }
    return warp_best(c);
}

#define TAIL_LIST_MAX 1024u

__global__ void __launch_bounds__(MG_NT) k_merge_tail() {
    cg::cluster_group cl = cg::this_cluster();
    extern __shared__ Best s_smax[];             // maxima of the superblocks this CTA owns: entry k * 16 + warp; then u32 s_sblist[]
    __shared__ Best s_best[MG_NT / 32];
    __shared__ Best s_cand[TAIL_MAX_CTAS];       // candidate of every CTA of the cluster (written through DSMEM)
    __shared__ Best s_win;
    __shared__ u64 s_range[4];
    __shared__ u32 s_stop, s_nlist, s_nsb;
    __shared__ u32 s_list[TAIL_LIST_MAX];        // dirty blocks of this CTA's superblocks (work list of the step)
    __shared__ u32 s_scan[SORT_MAX_BK];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u32 C = cl.num_blocks(), rank = cl.block_rank();
    const u32 warps_per_cta = MG_NT / 32;
    const u32 gwarp = rank * warps_per_cta + warp, total_warps = C * warps_per_cta;
    const u32 KS = (cM.n_super + total_warps - 1) / total_warps;   // superblocks per warp
    u32 *s_sblist = reinterpret_cast<u32 *>(s_smax + (size_t)KS * warps_per_cta);
    const bool token_cta = rank == C - 1;
    const u32 apply_ctas = C - 1;
    u64 prev_key = cM.ctr[6];
    const int first_step = (int)cM.ctr[1];
    u32 n_tok = 256 + (u32)first_step;
    u32 n_sorts = 0;
    if (tid == 0) {
        u64 st = *((volatile u64 *)&cM.ctr[3]), keys = *((volatile u64 *)&cM.ctr[2]);
        s_stop = st ? MG_STOP_ERROR : (keys * 2 > cM.pcap ? MG_STOP_GROW : 0u);
        s_nlist = 0; s_nsb = 0;
    }
    __syncthreads();
    bool first = true;

    for (int step = first_step; step < cM.stop_at; step++) {
        const u32 stop = s_stop;                 // uniform across the cluster (broadcast before the closing barrier)
        if (stop == MG_STOP_ERROR) break;
        if (stop == MG_STOP_GROW) {
            if (rank == 0 && tid == 0) { cM.ctr[6] = prev_key; cM.ctr[3] = MG_NEED_GROW; }
            break;
        }
        if (rank == 0 && tid == MG_NT - 1) cM.log_begin[step] = *((volatile u64 *)&cM.ctr[0]);
#ifdef BPE_MERGE_PROFILE
        const bool prof_thread = tid == 0 && (rank == 0 || token_cta);
#else
        const bool prof_thread = false;
#endif
        u64 t0 = prof_thread ? gtime_ns() : 0;
        // ---- phase 1a: collect the dirty blocks of the superblocks this CTA owns into a CTA-wide work list -----
        for (u32 k0 = 0; k0 < KS; k0 += 32) {
            const u32 kk = k0 + lane;
            const u32 sbl = gwarp + kk * total_warps;
            u64 my = (kk < KS && sbl < cM.n_super) ? cM.sdirty_mask[sbl] : 0;      // one round trip for 32 masks
            const u32 kend = KS - k0 < 32 ? KS - k0 : 32;
            for (u32 q = 0; q < kend; q++) {
                const u32 k = k0 + q, sb = gwarp + k * total_warps;
                u64 dm = __shfl_sync(0xffffffffu, my, q);
                if (sb >= cM.n_super) { if (first && lane == 0) s_smax[k * warps_per_cta + warp] = BEST_NONE; continue; }
                if (!first && dm == 0) continue; // (at kernel start every superblock maximum has to be built)
                const u32 nb = cM.n_blocks - sb * 64;
                if (nb < 64) dm &= (1ull << nb) - 1;
                const u32 cnt = __popcll(dm);
                u32 base = 0;
                if (lane == 0 && cnt) base = atomicAdd(&s_nlist, cnt);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (lane == 0 && dm) cM.sdirty_mask[sb] = 0;
                if (cnt && base + cnt > TAIL_LIST_MAX) {
                    // work list full (table just rebuilt: everything is dirty): this warp does the superblock on its own
                    for (u32 r = lane; r < cnt && base + r < TAIL_LIST_MAX; r += 32) s_list[base + r] = 0xFFFFFFFFu;
                    const u32 b0 = sb * 64 + lane, b1 = b0 + 32;
                    Best m0 = BEST_NONE, m1 = BEST_NONE;
                    if (b0 < cM.n_blocks) m0 = load_best(&cM.bmax[b0]);
                    if (b1 < cM.n_blocks) m1 = load_best(&cM.bmax[b1]);
                    while (dm) {
                        u32 j = __ffsll((long long)dm) - 1; dm &= dm - 1;
                        Best bst = rescan_block(sb * 64 + j, prev_key);
                        if (lane == (j & 31)) { if (j < 32) m0 = bst; else m1 = bst; store_best(&cM.bmax[sb * 64 + j], bst); }
                    }
                    Best c = BEST_NONE;
                    if (m0.cnt != CNT_DEAD) c = m0;
                    if (m1.cnt != CNT_DEAD && best_greater(m1, c)) c = m1;
                    c = warp_best(c);
                    if (lane == 0) s_smax[k * warps_per_cta + warp] = c;
                    continue;
                }
                for (u32 r = lane; r < cnt; r += 32) s_list[base + r] = sb * 64 + nth_set_bit64(dm, r);
                if (lane == 0) s_sblist[atomicAdd(&s_nsb, 1u)] = k * warps_per_cta + warp;
            }
        }
        first = false;
        __syncthreads();
        u64 ta = prof_thread ? gtime_ns() : 0;
        if (prof_thread && rank == 0) { cM.prof[20] += s_nlist; cM.prof[21] += s_nsb; }
        // ---- phase 1b: rescan the listed blocks, spread evenly over the warps ---------------------------------
        {
            const u32 nl = s_nlist < TAIL_LIST_MAX ? s_nlist : TAIL_LIST_MAX;
            for (u32 i = warp; i < nl; i += warps_per_cta) {
                const u32 blk = s_list[i];
                if (blk == 0xFFFFFFFFu) continue;
                Best bst = rescan_block(blk, prev_key);
                if (lane == 0) { store_best(&cM.bmax[blk], bst); PROF_ADD(9, 1); }
            }
        }
        __threadfence_block();
        __syncthreads();
        u64 tb = prof_thread ? gtime_ns() : 0;
        // ---- phase 1c: maxima of the superblocks that had dirty blocks ------------------------------------------
        {
            const u32 ns = s_nsb;
            for (u32 i = warp; i < ns; i += warps_per_cta) {
                const u32 local = s_sblist[i];
                const u32 sb = rank * warps_per_cta + (local % warps_per_cta) + (local / warps_per_cta) * total_warps;
                Best c = superblock_max(sb, lane);
                if (lane == 0) s_smax[local] = c;
            }
        }
        __syncthreads();
        u64 tc = prof_thread ? gtime_ns() : 0;
        if (tid == 0) { s_nlist = 0; s_nsb = 0; }
        // ---- phase 2: CTA candidate from its superblock maxima, exchanged through distributed shared memory ----
        {
            Best mine = BEST_NONE;
            for (u32 i = tid; i < KS * warps_per_cta; i += MG_NT) { Best e = s_smax[i]; if (e.cnt != CNT_DEAD && best_greater(e, mine)) mine = e; }
            mine = warp_best(mine);
            if (lane == 0) s_best[warp] = mine;
            __syncthreads();
            if (warp == 0) {
                Best c = lane < warps_per_cta ? s_best[lane] : BEST_NONE;
                c = warp_best(c);
                if (lane < C) store_best(cl.map_shared_rank(&s_cand[rank], lane), c);
            }
        }
        u64 t1 = prof_thread ? gtime_ns() : 0;
        cl.sync();
        u64 t2 = prof_thread ? gtime_ns() : 0;
        if (warp == 0) {
            Best c = lane < C ? load_best(&s_cand[lane]) : BEST_NONE;
            c = warp_best(c);
            if (lane == 0) {
                s_win = c;
                if (c.cnt != CNT_DEAD) winner_range(c.key, s_range);   // index range of the winner, read once per CTA
            }
        }
        __syncthreads();
        const Best win = s_win;
        if (win.cnt == CNT_DEAD) break;          // len(byte_pair_frequencies) == 0 (train.py:184-185): uniform across the cluster
        const u32 a = (u32)(win.key >> 32), b = (u32)win.key;
        const u32 nw = n_tok;
        const u64 r_lo = s_range[0], r_hi = s_range[1];
        if (token_cta) {
            token_bookkeeping(step, win, a, b, nw, prof_thread, t2);
            if (tid == 0) {                      // stop code for the next step, broadcast to every CTA
                u64 st = *((volatile u64 *)&cM.ctr[3]), keys = *((volatile u64 *)&cM.ctr[2]);
                u32 code = st ? MG_STOP_ERROR : (keys * 2 > cM.pcap ? MG_STOP_GROW : 0u);
                for (u32 r = 0; r < C; r++) *cl.map_shared_rank(&s_stop, r) = code;
            }
        } else {
            apply_winner(step, a, b, nw, r_lo, r_hi, T_SRC(s_range[2]), (u64)rank * MG_NT + tid, (u64)apply_ctas * MG_NT);
        }
        prev_key = win.key;
        n_tok++;
        u64 t3 = prof_thread ? gtime_ns() : 0;
        cl.sync();
        if (prof_thread) {
            u64 t4 = gtime_ns();
            if (rank == 0) {
                cM.prof[0] += t1 - t0; cM.prof[1] += t2 - t1; cM.prof[2] += t3 - t2; cM.prof[3] += t4 - t3; cM.prof[7] += 1;
                cM.prof[16] += ta - t0; cM.prof[17] += tb - ta; cM.prof[18] += tc - tb; cM.prof[19] += t1 - tc; cM.prof[22] += r_hi - r_lo;
            }
            if (token_cta) cM.prof[4] += t3 - t2;
        }
        if ((r_hi - r_lo) * 2 >= SORT_MIN) {     // heuristic: about two records per index entry visited
            u64 ts = prof_thread ? gtime_ns() : 0;
            const u64 lb = *((volatile u64 *)&cM.log_begin[step]), cur = *((volatile u64 *)&cM.ctr[0]);
            const u64 pool = *((volatile u64 *)&cM.ctr[9]);
            if (cur - lb >= SORT_MIN && cur <= cM.log_cap && pool + SORT_MAX_BK + 1 <= cM.bk_off_cap) {
                sort_slice(step, lb, (u32)(cur - lb), n_sorts & 1u, rank == 0, (u64)rank * MG_NT + tid, (u64)C * MG_NT, s_scan,
                           [&]() { cl.sync(); });
                n_sorts++;
            }
            if (prof_thread && rank == 0) { cM.prof[23] += gtime_ns() - ts; cM.prof[24] += 1; }
        }
    }
    if (rank == 0 && tid == 0) cM.ctr[6] = prev_key;
    cl.sync();                                   // nobody leaves while a peer may still touch its shared memory
}

__global__ void __launch_bounds__(256) k_insert_initial_pairs(const u64 *__restrict__ dense) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < 65536 && dense[p]) pair_add(p >> 8, p & 255u, (i64)dense[p]);
}

// Table growth: re-insert every live key of the old table (popped keys and the pending pop are dropped).
__global__ void __launch_bounds__(256) k_pairs_rehash(const u64 *__restrict__ okey, const i64 *__restrict__ ocnt, u64 ocap, u64 pending_pop) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < ocap; i += (u64)gridDim.x * blockDim.x) {
        u64 k = okey[i];
        if (k == PAIR_EMPTY || k == pending_pop) continue;
        i64 c = ocnt[i];
        if (c == CNT_DEAD) continue;
        if (c < CNT_DEAD_LIMIT) c = (i64)((u64)c - (u64)CNT_DEAD);     // popped, then touched again
        u64 mask = cM.pcap - 1, s = pair_hash(k) & mask;
        for (;;) {
            if (cM.pkey[s] == PAIR_EMPTY && atomicCAS(&cM.pkey[s], PAIR_EMPTY, k) == PAIR_EMPTY) { cM.pcnt[s] = c; break; }
            s = (s + 1) & mask;
        }
        atomicAdd(&cM.ctr[2], 1ull);
    }
}
