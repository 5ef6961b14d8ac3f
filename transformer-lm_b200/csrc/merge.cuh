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
//                so a step rescans only the blocks that changed and otherwise streams the (small)
//                array of cached maxima.
//   tie-break    (count, (bytes_a, bytes_b)) with python's bytes ordering.  Each token keeps its
//                first 8 bytes as a big-endian integer (tok_key): comparing the integers decides
//                almost every tie; equal prefixes fall back to lengths / a byte loop.
//   token_indices  for pairs of two initial bytes: CSR built once from the initial words; for any
//                other pair (p,q) every occurrence is created in the step that created the younger
//                of p,q, so it is found by filtering that step's slice of an append-only log.
//                Stale entries are harmless (the rewrite re-checks symbols, like the reference).
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#include "common.cuh"

namespace cg = cooperative_groups;

#define PAIR_EMPTY 0xFFFFFFFFFFFFFFFFull
#define CNT_DEAD ((i64)0x8000000000000000ll)     // key was popped from the dict (train.py:226)
#define CNT_DEAD_LIMIT ((i64)0xC000000000000000ll) // counts below this are "popped (+ later deltas)"
#define PB 256u                                  // pair-table slots per block
#define MG_NT 512
#define MG_NEED_GROW 8ull
#define CTA_BEST_STRIDE 32u                       // Best entries (1 KiB) between per-CTA candidates: spreads the all-read-all
                                                 // exchange over many L2 slices instead of hammering a handful of lines                        // ctr[3] code: pair table more than half full, host must grow it

// one 32-byte sector per word: the claim (atomicExch on stamp) pulls in everything the rewrite needs
struct __align__(32) WordMeta {
    u32 off, len;        // word w occupies sym[off, off+len)
    u32 stamp;           // last merge step (+1) that processed the word
    u32 pad0;
    i64 cnt;             // word frequency
    i64 pad1;
};
struct Words {
    int32_t *sym;        // symbols of all words
    WordMeta *meta;
    u64 *counters;       // [0]=n_words [1]=n_syms [2]=max_len
};

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
    const u32 *csr_off; const u32 *csr_words;
    uint2 *log; u64 *log_begin; u64 log_cap;
    u32 *tok_off; u32 *tok_len; u64 *tok_key; uint8_t *tok_bytes; u64 tok_bytes_cap;
    Best *cta_best;
    int32_t *merges_out; i64 *merge_cnt_out; int n_merges;
    // [0]=log cursor [1]=n_done (next step) [2]=pair keys created [3]=status flags [4]=tok bytes cursor
    // [5]=keys popped since the table was last rebuilt [6]=previous winner key still to be popped (PAIR_EMPTY if none)
    // [7]=sum over steps of the live pair-table keys (the reference's max() scans that many dict entries, train.py:187-189)
    u64 *ctr;
    // profile (ns / counts): [0]=phase1 [1]=sync1 [2]=apply [3]=sync2 on CTA 0; [4]=token CTA work; [5]=index records scanned;
    // [6]=words rewritten; [7]=steps; [9]=dirty blocks rescanned
    u64 *prof;
    u32 *step_prof;   // optional per-step trace: 4 x u32 per step (phase1+sync1 ns, apply+sync2 ns, records scanned, words rewritten so far)
};

// The loop state lives in constant memory: helpers are real (non-inlined) functions so that the persistent
// kernel stays small enough for the instruction caches -- every step runs each code path only once, so a
// large inlined kernel is instruction-fetch bound.
__constant__ MergeState cM;

#ifdef BPE_MERGE_PROFILE
__device__ __forceinline__ u64 gtime_ns() { u64 t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PROF_ADD(i, v) atomicAdd(&cM.prof[i], (u64)(v))
#else
__device__ __forceinline__ u64 gtime_ns() { return 0; }
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
__device__ __noinline__ bool best_tie_greater(u64 xkey, u64 xka, u64 xkb, u64 ykey, u64 yka, u64 ykb) {
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
__device__ __forceinline__ Best warp_best(Best b) {
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

// frequencies[key] += delta with defaultdict semantics (train.py:36,65-78): a missing key is created.
// The count update is a fire-and-forget reduction; a popped key that is touched again is repaired by
// the next rescan of its block (see rescan_block).  (s, k) = first probe slot and the key read there.
__device__ __noinline__ void pair_add_from(u64 key, i64 delta, u64 s, u64 k) {
    const u64 mask = cM.pcap - 1;
    for (u64 probes = 0; probes < cM.pcap; probes++) {
        if (k == PAIR_EMPTY) {
            u64 old = atomicCAS(&cM.pkey[s], PAIR_EMPTY, key);
            if (old == PAIR_EMPTY) { atomicAdd(&cM.ctr[2], 1ull); k = key; }
            else k = old;
        }
        if (k == key) {
            atomicAdd((u64 *)&cM.pcnt[s], (u64)delta);
            cM.dirty[s / PB] = 1;
            return;
        }
        s = (s + 1) & mask;
        k = cM.pkey[s];
    }
    cM.ctr[3] = 1;                                // table full
}
__device__ __forceinline__ void pair_add(u32 a, u32 b, i64 delta) {
    u64 key = ((u64)a << 32) | b;
    u64 s = mix64(key) & (cM.pcap - 1);
    pair_add_from(key, delta, s, cM.pkey[s]);
}

// Apply merge (a,b)->nw to word w: the left-to-right scan of train.py:196-224 with
// update_frequencies_after_merge (52-78), merge_subwords (132-139) and create_new_token_indices (107-129).
__device__ __noinline__ void apply_merge_to_word(u32 w, u32 a, u32 b, u32 nw) {
    WordMeta *wm = &cM.W.meta[w];
    int32_t *s = cM.W.sym + wm->off;
    const u32 len = wm->len;
    const i64 c = wm->cnt;
    // pass 1: how many index records will this word append (one per neighbour of every merge site)?
    u32 n_rec = 0, n_site = 0;
    {
        u32 r = 0, o = 0;
        while (r + 1 < len) {
            if ((u32)s[r] == a && (u32)s[r + 1] == b) { n_rec += (o > 0) + (r + 2 < len); n_site++; r += 2; }
            else r++;
            o++;
        }
    }
    if (n_site == 0) return;                     // stale index entry: the pair no longer occurs here
    u64 li = 0;
    if (n_rec) {                                 // one atomic per group of threads that arrive here together
        cg::coalesced_group cgp = cg::coalesced_threads();
        u32 pre = cg::exclusive_scan(cgp, n_rec);
        u64 base = 0;
        if (cgp.thread_rank() == cgp.size() - 1) base = atomicAdd(&cM.ctr[0], (u64)(pre + n_rec));
        li = cgp.shfl(base, cgp.size() - 1) + pre;
    }
    const bool log_ok = li + n_rec <= cM.log_cap;
    if (!log_ok) cM.ctr[3] = 2;
    const u64 mask = cM.pcap - 1;
    u32 o = 0, r = 0;
    while (r + 1 < len) {
        if ((u32)s[r] == a && (u32)s[r + 1] == b) {
            const bool has_l = o > 0, has_r = r + 2 < len;
            const u32 left = has_l ? (u32)s[o - 1] : 0, right = has_r ? (u32)s[r + 2] : 0;
            // the four dict updates of update_frequencies_after_merge: first probes issued together
            u64 k0 = ((u64)left << 32) | a, k1 = ((u64)left << 32) | nw, k2 = ((u64)b << 32) | right, k3 = ((u64)nw << 32) | right;
            u64 s0 = mix64(k0) & mask, s1 = mix64(k1) & mask, s2 = mix64(k2) & mask, s3 = mix64(k3) & mask;
            u64 v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (has_l) { v0 = cM.pkey[s0]; v1 = cM.pkey[s1]; }
            if (has_r) { v2 = cM.pkey[s2]; v3 = cM.pkey[s3]; }
            if (has_l) {
                pair_add_from(k0, -c, s0, v0);
                pair_add_from(k1, c, s1, v1);
                if (log_ok) cM.log[li] = make_uint2(left, w);                       // (left, nw): side L, other = left
                li++;
            }
            if (has_r) {
                pair_add_from(k2, -c, s2, v2);
                pair_add_from(k3, c, s3, v3);
                if (log_ok) cM.log[li] = make_uint2(0x80000000u | right, w);        // (nw, right): side R, other = right
                li++;
            }
            s[o++] = (int32_t)nw; r += 2;
        } else {
            s[o++] = s[r++];
        }
    }
#pragma unroll 1
    while (r < len) s[o++] = s[r++];
    wm->len = o;
    PROF_ADD(6, 1);
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
#pragma unroll 1
    for (u32 k = 0; k < PB / 32; k++) {          // rolled on purpose (code size); operands picked with a select chain
        Best c;
        c.key = keys[0]; c.cnt = cnts[0]; c.ka = kas[0]; c.kb = kbs[0];
#pragma unroll
        for (u32 j = 1; j < PB / 32; j++) if (k == j) { c.key = keys[j]; c.cnt = cnts[j]; c.ka = kas[j]; c.kb = kbs[j]; }
        if (c.key == PAIR_EMPTY) continue;
        if (c.key == prev_key) { cM.pcnt[sbase + k * 32 + lane_id()] = CNT_DEAD; continue; }
        if (c.cnt == CNT_DEAD) continue;
        if (c.cnt < CNT_DEAD_LIMIT) { c.cnt = (i64)((u64)c.cnt - (u64)CNT_DEAD); cM.pcnt[sbase + k * 32 + lane_id()] = c.cnt; }
        if (best_greater(c, bst)) bst = c;
    }
    return warp_best(bst);
}

__global__ void __launch_bounds__(MG_NT) k_merge_loop() {
    cg::grid_group grid = cg::this_grid();
    __shared__ Best s_best[MG_NT / 32];
    __shared__ Best s_win;
    __shared__ u64 s_status[2];
    __shared__ u64 s_range[2];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u32 G = gridDim.x;
    const u32 warps_per_cta = MG_NT / 32;
    const u32 gwarp = blockIdx.x * warps_per_cta + warp, total_warps = G * warps_per_cta;
    const bool token_cta = blockIdx.x == G - 1;
    const u32 apply_ctas = G - 1;
    u64 prev_key = cM.ctr[6];                     // winner of the previous step: popped lazily during the rescan
    const int first_step = (int)cM.ctr[1];
    u32 n_tok = 256 + (u32)first_step;

    for (int step = first_step; step < cM.n_merges; step++) {
        // status flags are written before the grid.sync that ends a step and read here: uniform across the grid
        if (tid == 0) { s_status[0] = *((volatile u64 *)&cM.ctr[3]); s_status[1] = *((volatile u64 *)&cM.ctr[2]); }
        __syncthreads();
        if (s_status[0]) break;
        if (s_status[1] * 2 > cM.pcap) {         // table over half full: hand back to the host to grow it
            grid.sync();
            if (blockIdx.x == 0 && tid == 0) { cM.ctr[6] = prev_key; cM.ctr[3] = MG_NEED_GROW; }
            return;
        }
        // ---- phase 1: rescan dirty blocks, reduce cached block maxima to one candidate per CTA ----
        if (blockIdx.x == 0 && tid == 0) cM.log_begin[step] = cM.ctr[0];
#ifdef BPE_MERGE_PROFILE
        const bool prof_thread = tid == 0 && (blockIdx.x == 0 || token_cta);
#else
        const bool prof_thread = false;
#endif
        u64 t0 = prof_thread ? gtime_ns() : 0;
        Best mine = BEST_NONE;
        for (u32 base = gwarp * 64; base < cM.n_blocks; base += total_warps * 64) {
            // 64 consecutive blocks per warp iteration: flags and cached maxima are loaded unconditionally
            u32 b0 = base + lane, b1 = base + 32 + lane;
            bool d0 = b0 < cM.n_blocks && cM.dirty[b0], d1 = b1 < cM.n_blocks && cM.dirty[b1];
            Best m0 = BEST_NONE, m1 = BEST_NONE;
            if (b0 < cM.n_blocks) m0 = load_best(&cM.bmax[b0]);
            if (b1 < cM.n_blocks) m1 = load_best(&cM.bmax[b1]);
            u64 dm = (u64)__ballot_sync(0xffffffffu, d0) | ((u64)__ballot_sync(0xffffffffu, d1) << 32);
            while (dm) {                         // warp-cooperative rescan of every dirty block (single inlined copy)
                u32 l2 = __ffsll((long long)dm) - 1; dm &= dm - 1;
                Best bst = rescan_block(base + l2, prev_key);
                if (lane == (l2 & 31)) {
                    if (l2 < 32) m0 = bst; else m1 = bst;
                    store_best(&cM.bmax[base + l2], bst); cM.dirty[base + l2] = 0; PROF_ADD(9, 1);
                }
            }
            if (m0.cnt != CNT_DEAD && best_greater(m0, mine)) mine = m0;
            if (m1.cnt != CNT_DEAD && best_greater(m1, mine)) mine = m1;
        }
        mine = warp_best(mine);
        if (lane == 0) s_best[warp] = mine;
        __syncthreads();
        if (warp == 0) {
            Best c = lane < warps_per_cta ? s_best[lane] : BEST_NONE;
            c = warp_best(c);
            if (lane == 0) store_best(&cM.cta_best[blockIdx.x * CTA_BEST_STRIDE], c);
        }
        u64 t1 = prof_thread ? gtime_ns() : 0;
        grid.sync();
        u64 t2 = prof_thread ? gtime_ns() : 0;
        // ---- phase 2: every CTA derives the same winner ----------------------------------------
        if (warp == 0) {
            Best c = BEST_NONE;
            Best o5[5];                          // G <= 160: every lane issues its (up to) 5 loads before using any
#pragma unroll
            for (u32 k = 0; k < 5; k++) { u32 i = lane + 32 * k; o5[k] = i < G ? load_best(&cM.cta_best[i * CTA_BEST_STRIDE]) : BEST_NONE; }
#pragma unroll
            for (u32 k = 0; k < 5; k++) if (o5[k].cnt != CNT_DEAD && best_greater(o5[k], c)) c = o5[k];
            c = warp_best(c);
            if (lane == 0) {
                s_win = c;
                if (c.cnt != CNT_DEAD && !token_cta) {           // index range of the winner, read once per CTA
                    u32 wa = (u32)(c.key >> 32), wb = (u32)c.key, wT = wa > wb ? wa : wb;
                    if (wT < 256) { u32 pp = (wa << 8) | wb; s_range[0] = cM.csr_off[pp]; s_range[1] = cM.csr_off[pp + 1]; }
                    else { s_range[0] = cM.log_begin[wT - 256]; s_range[1] = cM.log_begin[wT - 256 + 1]; }
                }
            }
        }
        __syncthreads();
        const Best win = s_win;
        if (win.cnt == CNT_DEAD) break;          // len(byte_pair_frequencies) == 0 (train.py:184-185)
        const u32 a = (u32)(win.key >> 32), b = (u32)win.key;
        const u32 nw = n_tok;                    // symbol id of new_byte = a + b (train.py:190)

        if (token_cta) {
            // ---- token bookkeeping: bytes and prefix key of the new token, outputs -------------------
            u64 tA = prof_thread ? gtime_ns() : 0;
            const u32 la = cM.tok_len[a], lb = cM.tok_len[b], ln = la + lb;
            const u32 oa = cM.tok_off[a], ob = cM.tok_off[b];
            const u64 cur = cM.ctr[4];
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
                    cM.ctr[7] += s_status[1] - cM.ctr[5]; cM.ctr[5] += 1;
                    if (prof_thread) { cM.prof[10] += tA - t2; cM.prof[11] += tB - tA; cM.prof[12] += gtime_ns() - tB; }
                }
            }
            // make sure the winner's block is rescanned so the key gets popped
            if (tid == 32) {
                u64 mask = cM.pcap - 1, s = mix64(win.key) & mask;
                while (cM.pkey[s] != win.key) s = (s + 1) & mask;
                cM.dirty[s / PB] = 1;
            }
        } else {
            // ---- apply the merge to every word indexed under (a,b) ------------------------------
            const u32 T = a > b ? a : b;
            const u64 gthread = (u64)blockIdx.x * MG_NT + tid, gstride = (u64)apply_ctas * MG_NT;
            if (T < 256) {
                const u64 lo = s_range[0], hi = s_range[1];
                if (gthread == 0) PROF_ADD(5, hi - lo);
                for (u64 i = lo + gthread; i < hi; i += 4 * gstride) {       // 4 candidates in flight per thread
                    const u32 none = 0xFFFFFFFFu;
                    u32 w0 = cM.csr_words[i];
                    u32 w1 = i + gstride < hi ? cM.csr_words[i + gstride] : none;
                    u32 w2 = i + 2 * gstride < hi ? cM.csr_words[i + 2 * gstride] : none;
                    u32 w3 = i + 3 * gstride < hi ? cM.csr_words[i + 3 * gstride] : none;
#pragma unroll 1
                    for (int k = 0; k < 4; k++) {
                        u32 w = k == 0 ? w0 : k == 1 ? w1 : k == 2 ? w2 : w3;
                        if (w != none && atomicExch(&cM.W.meta[w].stamp, (u32)step + 1) != (u32)step + 1) apply_merge_to_word(w, a, b, nw);
                    }
                }
            } else {
                const u64 lo = s_range[0], hi = s_range[1];
                u32 want = b >= a ? a : (0x80000000u | b);
                if (gthread == 0) PROF_ADD(5, hi - lo);
                for (u64 i = lo + gthread; i < hi; i += 4 * gstride) {       // 4 records in flight per thread
                    const uint2 none = make_uint2(0xFFFFFFFFu, 0);
                    uint2 r0 = cM.log[i];
                    uint2 r1 = i + gstride < hi ? cM.log[i + gstride] : none;
                    uint2 r2 = i + 2 * gstride < hi ? cM.log[i + 2 * gstride] : none;
                    uint2 r3 = i + 3 * gstride < hi ? cM.log[i + 3 * gstride] : none;
#pragma unroll 1
                    for (int k = 0; k < 4; k++) {
                        uint2 r = k == 0 ? r0 : k == 1 ? r1 : k == 2 ? r2 : r3;
                        if (r.x == want && atomicExch(&cM.W.meta[r.y].stamp, (u32)step + 1) != (u32)step + 1) apply_merge_to_word(r.y, a, b, nw);
                    }
                }
            }
        }
        prev_key = win.key;
        n_tok++;
        u64 t3 = prof_thread ? gtime_ns() : 0;
        grid.sync();
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
        u64 mask = cM.pcap - 1, s = mix64(k) & mask;
        for (;;) {
            if (cM.pkey[s] == PAIR_EMPTY && atomicCAS(&cM.pkey[s], PAIR_EMPTY, k) == PAIR_EMPTY) { cM.pcnt[s] = c; break; }
            s = (s + 1) & mask;
        }
        atomicAdd(&cM.ctr[2], 1ull);
    }
}
