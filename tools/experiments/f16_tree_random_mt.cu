// ExtraTrees, THROUGHPUT variant: one CTA grows MT_TW trees at once.
//   node_split_random      sklearn/tree/_splitter.pyx:507-736
//   builder loop           sklearn/tree/_tree.pyx:139-336
// (reference call site: ExtraTreesClassifier(random_state=0).fit, experiment.py:96,469)
//
// Why: the nodes of a tree are sequential (its xorshift stream is consumed depth-first), so a tree
// alternates between two kinds of work: wide sweeps over big nodes (thousands of rows: worth a whole
// CTA) and long chains of tiny nodes (a few dozen rows: worth one warp).  With one tree per CTA
// (f16_tree_random.cu) seven of the eight warps sit at a barrier while warp 0 grows the small
// subtrees - 35 % of all warp-time under saturation (profiles/r2_sat_et_stalls_by_line.txt) - and
// their registers and shared memory cap the SM at four trees.  Here the roles are split:
//   * MT_TW TREE WARPS, one tree each, never meet a block barrier.  A tree warp runs the scalar
//     feature draw, grows every subtree of <= S rows out of its own shared-memory region
//     (subtree_warp_v2), sweeps nodes of up to MT_WIDE rows itself (gathers from the L2-resident
//     row matrix), and POSTS the three sweeps of wider nodes to
//   * MT_NH HELPER WARPS, which serve the CTA's trees in turn: min/max of all features, left
//     counts of <= 4 candidate thresholds, stable partition - the same sweeps as the one-tree
//     kernel, synchronised among themselves with a named barrier.
// Requests and results travel through a mailbox per tree in shared memory (volatile state word,
// __threadfence_block on both sides); nothing crosses a CTA, so there is no forward-progress
// assumption between CTAs.  Three trees + five helper warps per CTA, four CTAs per SM: 12 trees
// resident per SM instead of 4.  The trees are bit-identical to the one-tree kernel's (same
// arithmetic, same stream), whatever the schedule.  Compile with -fmad=false.
#include "f16_tree_dev.cuh"
#include "f16_tree_random_sub.cuh"

#ifndef MT_TW
#define MT_TW 3                 // tree warps per CTA
#endif
#define MT_NH (NW - MT_TW)      // helper warps
#define MT_HT (MT_NH * 32)      // helper threads
#ifndef MT_S16
#define MT_S16 240              // rows of the shared-memory subtree region, d > 8   (3 x 17.9 KB + scratch: 4 CTAs / SM)
#endif
#ifndef MT_S8
#define MT_S8 480
#endif
#ifndef MT_WIDE
#define MT_WIDE 1024            // nodes with more rows are swept by the helper warps
#endif
#ifndef MT_U
#define MT_U 4                  // row loads a tree warp keeps in flight per lane in its own sweeps
#endif
#ifndef MT_PU
#define MT_PU 8                 // 32-row tiles per round of the tree warp's own partition
#endif
#ifndef MT_MINB
#define MT_MINB 4
#endif
#define MT_SSTK 24              // stack records cached in shared memory per tree
#define MT_SPIN_LIMIT (1 << 25)

enum { MT_OP_MINMAX = 1, MT_OP_COUNT = 2, MT_OP_PART = 3 };

struct MtReq {
    volatile int state;         // 0 free, 1 posted by the tree warp, 2 served by the helpers
    int op, start, nn, ncand, n_left;
    const uint32_t* src;
    uint32_t* dst;
    int cand_f[4];
    float cand_tf[4];           // thresholds rounded down to float32: x <= thr (f64)  <=>  x <= tf for float32 x
    int best_f;
    float best_tf;
    float mn[F16_MAX_D], mx[F16_MAX_D];
    unsigned long long cnt[4];  // left count | left positives << 32
};

template <int DP, int S>
struct MtTree {
    float s_col[DP * (S + 1)];
    uint16_t s_idx[2][S];
    uint8_t s_y[S];
    F16StackRec s_stack[MT_SSTK];
    Ctl c;
    DrawState ds;
    MtReq req;
};

struct MtScratch {
    float wmin[MT_NH][F16_MAX_D], wmax[MT_NH][F16_MAX_D];
    unsigned long long part[MT_NH][4];
    int wcnt[2][PU][MT_NH];
    volatile int pick;
    volatile int done;          // tree warps that have finished
};

struct MtStack {                // same record layout as TreeStack, smaller shared-memory cache
    F16StackRec* smem;
    F16StackRec* gmem;
    __device__ __forceinline__ F16StackRec get(int i) const { return i < MT_SSTK ? smem[i] : gmem[i]; }
    __device__ __forceinline__ void put(int i, const F16StackRec& r) const { if (i < MT_SSTK) smem[i] = r; else gmem[i] = r; }
};

__device__ __forceinline__ void mt_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MT_HT) : "memory"); }

// ------------------------------------------------------------------ helper side: the three sweeps
template <int DP>
__device__ __forceinline__ void mt_serve_minmax(MtReq& r, const float* __restrict__ X, MtScratch& hs, int htid) {
    constexpr int Q = DP / 4, SPI = MT_HT / Q;
    const int lane = htid & 31, hw = htid >> 5;
    const int q = htid % Q, sl = htid / Q;
    const int start = r.start, nn = r.nn;
    const uint32_t* src = r.src;
    float mn[4], mx[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
    for (int i0 = sl; i0 < nn; i0 += SPI * 4) {
        uint32_t id[4]; float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { int i = i0 + u * SPI; id[u] = f16_id(src[start + (i < nn ? i : i0)]); }
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)id[u] * DP) + q);
#pragma unroll
        for (int u = 0; u < 4; u++) {     // tail slots re-read row i0: harmless for min/max
            mn[0] = fminf(mn[0], v[u].x); mx[0] = fmaxf(mx[0], v[u].x);
            mn[1] = fminf(mn[1], v[u].y); mx[1] = fmaxf(mx[1], v[u].y);
            mn[2] = fminf(mn[2], v[u].z); mx[2] = fmaxf(mx[2], v[u].z);
            mn[3] = fminf(mn[3], v[u].w); mx[3] = fmaxf(mx[3], v[u].w);
        }
    }
#pragma unroll
    for (int off = Q; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            mn[j] = fminf(mn[j], __shfl_xor_sync(F16_FULL, mn[j], off));
            mx[j] = fmaxf(mx[j], __shfl_xor_sync(F16_FULL, mx[j], off));
        }
    }
    if (lane < Q) {
#pragma unroll
        for (int j = 0; j < 4; j++) { hs.wmin[hw][lane * 4 + j] = mn[j]; hs.wmax[hw][lane * 4 + j] = mx[j]; }
    }
    mt_bar();
    if (htid < DP) {
        float a = hs.wmin[0][htid], b = hs.wmax[0][htid];
#pragma unroll
        for (int w = 1; w < MT_NH; w++) { a = fminf(a, hs.wmin[w][htid]); b = fmaxf(b, hs.wmax[w][htid]); }
        r.mn[htid] = a; r.mx[htid] = b;
    }
}

template <int DP>
__device__ __forceinline__ void mt_serve_count(MtReq& r, const float* __restrict__ X, MtScratch& hs, int htid) {
    // four threads share a row: thread t loads the 16-byte quarter that holds candidate t's feature
    const int lane = htid & 31, hw = htid >> 5;
    const int start = r.start, nn = r.nn, ncand = r.ncand;
    const uint32_t* src = r.src;
    const int t4 = htid & 3;
    const int k = (t4 < ncand) ? t4 : 0;
    const int fk = r.cand_f[k];
    const float tk = r.cand_tf[k];
    const int qk = fk >> 2, ck = fk & 3;
    unsigned long long acc = 0;
    constexpr int RPS = MT_HT / 4;
    for (int i0 = htid >> 2; i0 < nn; i0 += RPS * 4) {
        uint32_t e[4]; float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { int i = i0 + u * RPS; e[u] = (i < nn) ? src[start + i] : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < 4; u++)
            v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)f16_id(e[u] == 0xffffffffu ? 0u : e[u]) * DP) + qk);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float x = ck == 0 ? v[u].x : ck == 1 ? v[u].y : ck == 2 ? v[u].z : v[u].w;
            if (e[u] != 0xffffffffu && x <= tk) acc += 1ull | ((unsigned long long)f16_y(e[u]) << 32);
        }
    }
    acc += __shfl_xor_sync(F16_FULL, acc, 4);
    acc += __shfl_xor_sync(F16_FULL, acc, 8);
    acc += __shfl_xor_sync(F16_FULL, acc, 16);
    if (lane < 4) hs.part[hw][lane] = acc;
    mt_bar();
    if (htid < 4) {
        unsigned long long s = 0;
#pragma unroll
        for (int w = 0; w < MT_NH; w++) s += hs.part[w][htid];
        r.cnt[htid] = s;
    }
}

// stable out-of-place partition by the helper threads (block_partition of f16_tree_dev.cuh for MT_HT threads)
template <int DP>
__device__ __forceinline__ void mt_serve_partition(MtReq& r, const float* __restrict__ X, MtScratch& hs, int htid) {
    const int lane = htid & 31, hw = htid >> 5;
    const int start = r.start, n = r.nn, n_left = r.n_left, bf = r.best_f;
    const float bthr = r.best_tf;
    const uint32_t* src = r.src;
    uint32_t* dst = r.dst;
    int run_l = 0, buf = 0;
    for (int base = 0; base < n; base += MT_HT * PU, buf ^= 1) {
        uint32_t e[PU]; bool valid[PU], left[PU]; unsigned bal[PU];
        const int nj = min(PU, (n - base + MT_HT - 1) / MT_HT);
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) {
                int p = base + j * MT_HT + htid;
                valid[j] = p < n;
                e[j] = valid[j] ? src[start + p] : 0u;
            }
        }
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) left[j] = valid[j] && (__ldg(X + (size_t)f16_id(e[j]) * DP + bf) <= bthr);
        }
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) {
                bal[j] = __ballot_sync(F16_FULL, left[j]);
                if (lane == 0) hs.wcnt[buf][j][hw] = __popc(bal[j]);
            }
        }
        mt_bar();
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) {
                int before = 0, tot = 0;
#pragma unroll
                for (int w = 0; w < MT_NH; w++) { int cq = hs.wcnt[buf][j][w]; if (w < hw) before += cq; tot += cq; }
                int lrank = before + __popc(bal[j] & ((1u << lane) - 1u));
                if (valid[j]) {
                    if (left[j]) dst[start + run_l + lrank] = e[j];
                    else dst[start + n_left + (base + j * MT_HT - run_l) + (htid - lrank)] = e[j];
                }
                run_l += tot;
            }
        }
    }
}

template <int DP, int S>
__device__ void mt_helper_loop(const F16FitParams& P, MtTree<DP, S>* trees, MtScratch& hs, int htid) {
    int rr = 0;
    while (true) {
        if (htid == 0) {
            int pick = -1;
            for (int spin = 0; pick == -1; spin++) {
#pragma unroll
                for (int k = 0; k < MT_TW; k++) {
                    const int s = (rr + k) % MT_TW;
                    if (pick == -1 && trees[s].req.state == 1) pick = s;
                }
                if (pick == -1) {
                    if (hs.done >= MT_TW) pick = -2;
                    else if (spin > MT_SPIN_LIMIT) { atomicExch(P.err, F16_ERR_CUDA); pick = -2; }
                    else __nanosleep(100);
                }
            }
            hs.pick = pick;
        }
        mt_bar();
        const int pick = hs.pick;
        if (pick < 0) break;
        rr = pick + 1;
        __threadfence_block();                        // the request's fields were written before its state
        MtReq& r = trees[pick].req;
        if (r.op == MT_OP_MINMAX) mt_serve_minmax<DP>(r, P.X, hs, htid);
        else if (r.op == MT_OP_COUNT) mt_serve_count<DP>(r, P.X, hs, htid);
        else mt_serve_partition<DP>(r, P.X, hs, htid);
        __threadfence_block();                        // every helper's results (shared and global) ...
        mt_bar();                                     // ... are complete before the state word says so
        if (htid == 0) r.state = 2;
    }
}

// ------------------------------------------------------------------ tree-warp side
__device__ __forceinline__ bool mt_call(MtReq& r, int lane, int32_t* err) {
    // the request fields were written by lane 0 before this; every lane publishes its own earlier
    // global writes (the root's sample list) before lane 0 releases the request
    __threadfence_block();
    __syncwarp();
    bool ok = true;
    if (lane == 0) {
        __threadfence_block();
        r.state = 1;
        int spin = 0;
        while (r.state != 2) {
            __nanosleep(60);
            if (++spin > MT_SPIN_LIMIT) { atomicExch(err, F16_ERR_CUDA); ok = false; break; }
        }
        __threadfence_block();
    }
    ok = __shfl_sync(F16_FULL, ok ? 1 : 0, 0) != 0;
    __syncwarp();
    return ok;
}
__device__ __forceinline__ void mt_release(MtReq& r, int lane) {
    __syncwarp();
    if (lane == 0) r.state = 0;
}

template <int DP, int S>
__device__ void mt_tree_warp(const F16FitParams& P, const int t, MtTree<DP, S>& T, const int lane) {
    constexpr int Q = DP / 4;          // float4 quads per row
    constexpr int RPI = 32 / Q;        // rows per warp step in the min/max sweep
    constexpr int SP = S + 1;
    const unsigned lt = (1u << lane) - 1u;
    const int n = P.n, d = P.d, max_features = P.max_features;
    const float* __restrict__ X = P.X;
    uint32_t* buf0 = P.buf + (size_t)t * 2 * n;
    uint32_t* buf1 = buf0 + n;
    F16Node* nodes = P.nodes + (size_t)t * P.node_cap;
    Ctl& c = T.c;
    DrawState& ds = T.ds;
    MtReq& rq = T.req;
    MtStack stk;
    stk.smem = T.s_stack;
    stk.gmem = P.stack + (size_t)t * P.stack_cap;
    const double W_total = (double)n;

    // ---- root: identity sample list with packed labels; class counts
    int c1 = 0;
    for (int i = lane; i < n; i += 32) {
        uint32_t y = P.y[i];
        buf0[i] = f16_pack((uint32_t)i, 1u, y);
        c1 += (int)y;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c1 += __shfl_xor_sync(F16_FULL, c1, off);
    if (lane == 0) {
        for (int f = 0; f < F16_MAX_D; f++) { ds.features[f] = f; ds.const_feats[f] = 0; }
        ds.rng = P.rand_r_state[t];
        c.done = 0; c.abort = 0;
    }
    __syncwarp();
    // registers mirror the scalar state; every lane holds the same values
    int sp = 0, node_count = 0;
    uint32_t rng = P.rand_r_state[t];
    {
        F16StackRec r;
        r.start = 0; r.end = n; r.parent = -1; r.c0 = n - c1; r.c1 = c1;
        r.const_mask = 0; r.n_const = 0; r.is_left = 0; r.pad = 0; r.depth = 0;
        if (lane == 0) stk.put(0, r);
        sp = 1;
    }
    __syncwarp();

    while (sp > 0) {
        F16StackRec r = stk.get(sp - 1);
        if (r.pad) {
            // ---- SHARED regime: the shared-memory subtree routine keeps its scalars in `c` / `ds`
            if (lane == 0) { c.sp = sp; c.node_count = node_count; ds.rng = rng; }
            __syncwarp();
            subtree_warp_v2<DP, S>(c, ds, stk, P, nodes, T.s_col, T.s_idx, T.s_y);
            __syncwarp();
            sp = c.sp; node_count = c.node_count; rng = ds.rng;
            if (c.abort) break;
            continue;
        }
        sp--;
        const int start = r.start, nn = r.end - r.start;
        const uint32_t* src = (r.depth & 1) ? buf1 : buf0;
        uint32_t* dst = (r.depth & 1) ? buf0 : buf1;
        const int t0 = r.c0, t1 = r.c1;
        const bool leaf = leaf_pretest(nn, t0, t1);

        if (!leaf && nn <= S) {
            // ---- relocate the node into this tree's shared-memory region, then the SHARED regime
            const int q = lane % Q;
            for (int i = lane / Q; i < nn; i += RPI) {
                uint32_t e = src[start + i];
                float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)f16_id(e) * DP) + q);
                T.s_col[(q * 4 + 0) * SP + i] = v.x;
                T.s_col[(q * 4 + 1) * SP + i] = v.y;
                T.s_col[(q * 4 + 2) * SP + i] = v.z;
                T.s_col[(q * 4 + 3) * SP + i] = v.w;
                if (q == 0) { T.s_idx[r.depth & 1][i] = (uint16_t)i; T.s_y[i] = (uint8_t)f16_y(e); }
            }
            if (lane == 0) {
                F16StackRec m = r;
                m.start = 0; m.end = nn; m.pad = 1;
                stk.put(sp, m);
            }
            sp++;
            __syncwarp();
            continue;
        }

        bool split = false;
        int best_f = -2, n_left = 0, bl0 = 0, bl1 = 0;
        double best_thr = -2.0;
        int n_total = r.n_const;
        uint32_t cmask = r.const_mask;
        const bool wide = nn > MT_WIDE;                // the helper warps sweep this node

        if (!leaf) {
            // ---- sweep 1: min / max of every feature
            float mn[4], mx[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
            if (wide) {
                if (lane == 0) { rq.op = MT_OP_MINMAX; rq.start = start; rq.nn = nn; rq.src = src; rq.dst = dst; }
                if (!mt_call(rq, lane, P.err)) break;
                if (lane < Q) {
#pragma unroll
                    for (int j = 0; j < 4; j++) { mn[j] = rq.mn[lane * 4 + j]; mx[j] = rq.mx[lane * 4 + j]; }
                }
                mt_release(rq, lane);
            } else {
                const int q = lane % Q, rl = lane / Q;
                for (int i0 = rl; i0 < nn; i0 += RPI * MT_U) {
                    uint32_t id[MT_U]; float4 v[MT_U];
#pragma unroll
                    for (int u = 0; u < MT_U; u++) { int i = i0 + u * RPI; id[u] = f16_id(src[start + (i < nn ? i : i0)]); }
#pragma unroll
                    for (int u = 0; u < MT_U; u++) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)id[u] * DP) + q);
#pragma unroll
                    for (int u = 0; u < MT_U; u++) {
                        mn[0] = fminf(mn[0], v[u].x); mx[0] = fmaxf(mx[0], v[u].x);
                        mn[1] = fminf(mn[1], v[u].y); mx[1] = fmaxf(mx[1], v[u].y);
                        mn[2] = fminf(mn[2], v[u].z); mx[2] = fmaxf(mx[2], v[u].z);
                        mn[3] = fminf(mn[3], v[u].w); mx[3] = fmaxf(mx[3], v[u].w);
                    }
                }
#pragma unroll
                for (int off = Q; off < 32; off <<= 1) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        mn[j] = fminf(mn[j], __shfl_xor_sync(F16_FULL, mn[j], off));
                        mx[j] = fmaxf(mx[j], __shfl_xor_sync(F16_FULL, mx[j], off));
                    }
                }
            }
            // lane q (< Q) now holds min/max of features 4q .. 4q+3
            auto feat_minmax = [&](int f, float& fmn, float& fmx) {
                const int fq = f >> 2, fj = f & 3;
                float a = (fj == 0) ? mn[0] : (fj == 1) ? mn[1] : (fj == 2) ? mn[2] : mn[3];
                float b = (fj == 0) ? mx[0] : (fj == 1) ? mx[1] : (fj == 2) ? mx[2] : mx[3];
                fmn = __shfl_sync(F16_FULL, a, fq);
                fmx = __shfl_sync(F16_FULL, b, fq);
            };

            // ---- feature draw (all lanes run the scalar loop; lane 0 applies the swaps)
            int f_i = d, n_visited = 0, n_found = 0, n_drawn = 0, ncand = 0;
            const int n_known = r.n_const;
            int cf[4] = {0, 0, 0, 0};
            double ct[4] = {0.0, 0.0, 0.0, 0.0};
            double best = -INFINITY;
            bool failed = false;

            auto eval_chunk = [&](int cnt) {
                // ---- sweep 2: left counts of up to 4 candidates
                int nl[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0};
                if (wide) {
                    if (lane == 0) {
                        rq.op = MT_OP_COUNT; rq.start = start; rq.nn = nn; rq.src = src; rq.ncand = cnt;
                        for (int k = 0; k < 4; k++) { rq.cand_f[k] = cf[k]; rq.cand_tf[k] = __double2float_rd(ct[k]); }
                    }
                    if (!mt_call(rq, lane, P.err)) { failed = true; return; }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k < cnt) { const unsigned long long s = rq.cnt[k]; nl[k] = (int)(uint32_t)s; l1[k] = (int)(s >> 32); }
                    }
                    mt_release(rq, lane);
                } else {
                    // four lanes share a row: lane t loads the quarter that holds candidate t's feature
                    const int t4 = lane & 3;
                    const int k = (t4 < cnt) ? t4 : 0;
                    int fk = cf[0]; float tk = __double2float_rd(ct[0]);
#pragma unroll
                    for (int j = 1; j < 4; j++) if (k == j) { fk = cf[j]; tk = __double2float_rd(ct[j]); }
                    const int qk = fk >> 2, ck = fk & 3;
                    unsigned long long acc = 0;
                    for (int i0 = lane >> 2; i0 < nn; i0 += 8 * MT_U) {
                        uint32_t e[MT_U]; float4 v[MT_U];
#pragma unroll
                        for (int u = 0; u < MT_U; u++) { int i = i0 + u * 8; e[u] = (i < nn) ? src[start + i] : 0xffffffffu; }
#pragma unroll
                        for (int u = 0; u < MT_U; u++)
                            v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)f16_id(e[u] == 0xffffffffu ? 0u : e[u]) * DP) + qk);
#pragma unroll
                        for (int u = 0; u < MT_U; u++) {
                            const float x = ck == 0 ? v[u].x : ck == 1 ? v[u].y : ck == 2 ? v[u].z : v[u].w;
                            if (e[u] != 0xffffffffu && x <= tk) acc += 1ull | ((unsigned long long)f16_y(e[u]) << 32);
                        }
                    }
                    acc += __shfl_xor_sync(F16_FULL, acc, 4);
                    acc += __shfl_xor_sync(F16_FULL, acc, 8);
                    acc += __shfl_xor_sync(F16_FULL, acc, 16);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const unsigned long long s = __shfl_sync(F16_FULL, acc, j);
                        if (j < cnt) { nl[j] = (int)(uint32_t)s; l1[j] = (int)(s >> 32); }
                    }
                }
                const int k = (lane >> 1) & 3;
                int mnl = nl[0], ml1 = l1[0];
#pragma unroll
                for (int j = 1; j < 4; j++) if (k == j) { mnl = nl[j]; ml1 = l1[j]; }
                double a, b, part;
                if (lane & 1) { a = (double)(mnl - ml1); b = (double)ml1; double w = a + b; part = w * gini_of(a, b, w); }
                else { a = (double)(t0 - (mnl - ml1)); b = (double)(t1 - ml1); double w = a + b; part = (-w) * gini_of(a, b, w); }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j < cnt) {
                        double proxy = __shfl_sync(F16_FULL, part, 2 * j) - __shfl_sync(F16_FULL, part, 2 * j + 1);
                        if (proxy > best) {
                            best = proxy; best_f = cf[j]; best_thr = ct[j]; n_left = nl[j]; bl1 = l1[j]; bl0 = nl[j] - l1[j];
                        }
                    }
                }
            };

            while (!failed && f_i > n_total && (n_visited < max_features || n_visited <= n_found + n_drawn)) {
                n_visited++;
                int f_j = f16_rand_int(n_drawn, f_i - n_found, &rng);
                if (f_j < n_known) {
                    int a = ds.features[n_drawn], b = ds.features[f_j];
                    __syncwarp();
                    if (lane == 0) { ds.features[n_drawn] = b; ds.features[f_j] = a; }
                    __syncwarp();
                    n_drawn++;
                    continue;
                }
                f_j += n_found;
                const int f = ds.features[f_j];
                float fmn, fmx;
                feat_minmax(f, fmn, fmx);
                if (fmx <= __fadd_rn(fmn, 1e-7f)) {
                    int b = ds.features[n_total];
                    __syncwarp();
                    if (lane == 0) { ds.features[f_j] = b; ds.features[n_total] = f; }
                    __syncwarp();
                    n_found++; n_total++;
                    continue;
                }
                f_i--;
                {
                    int b = ds.features[f_i];
                    __syncwarp();
                    if (lane == 0) { ds.features[f_i] = f; ds.features[f_j] = b; }
                    __syncwarp();
                }
                double thr = f16_rand_uniform((double)fmn, (double)fmx, &rng);
                if (thr == (double)fmx) thr = (double)fmn;
#pragma unroll
                for (int j = 0; j < 4; j++) if (j == ncand) { cf[j] = f; ct[j] = thr; }
                ncand++;
                if (ncand == 4) { eval_chunk(4); ncand = 0; }
            }
            if (!failed && ncand > 0) eval_chunk(ncand);
            if (failed) break;
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < n_known; i++) ds.features[i] = ds.const_feats[i];
                for (int i = n_known; i < n_total; i++) ds.const_feats[i] = ds.features[i];
            }
            __syncwarp();
            for (int i = n_known; i < n_total; i++) cmask |= 1u << ds.const_feats[i];
            if (best_f >= 0) split = improvement_certain(t0, t1, n) || improvement_ok(bl0, bl1, t0, t1, W_total);
        }

        // ---- node record
        const int id = node_count++;
        if (id >= P.node_cap || sp + 2 > P.stack_cap) { if (lane == 0) atomicExch(P.err, F16_ERR_OVERFLOW); break; }
        if (lane == 0) {
            F16Node nd;
            nd.thr = split ? best_thr : -2.0;
            nd.feature = split ? best_f : -2;
            nd.right = -1;
            nd.c0 = t0; nd.c1 = t1; nd.n = nn; nd.depth = r.depth;
            nodes[id] = nd;
            if (r.parent >= 0 && !r.is_left) nodes[r.parent].right = id;
        }
        if (split) {
            // ---- sweep 3: stable partition
            const float best_tf = __double2float_rd(best_thr);
            if (wide) {
                if (lane == 0) {
                    rq.op = MT_OP_PART; rq.start = start; rq.nn = nn; rq.src = src; rq.dst = dst;
                    rq.best_f = best_f; rq.best_tf = best_tf; rq.n_left = n_left;
                }
                if (!mt_call(rq, lane, P.err)) break;
                mt_release(rq, lane);
            } else {
                int run_l = 0;
                for (int base = 0; base < nn; base += 32 * MT_PU) {
                    const int nj = min(MT_PU, (nn - base + 31) >> 5);
                    uint32_t e[MT_PU]; bool valid[MT_PU], left[MT_PU];
#pragma unroll
                    for (int j = 0; j < MT_PU; j++) {
                        if (j < nj) { int p = base + j * 32 + lane; valid[j] = p < nn; e[j] = valid[j] ? src[start + p] : 0u; }
                    }
#pragma unroll
                    for (int j = 0; j < MT_PU; j++) {
                        if (j < nj) left[j] = valid[j] && (__ldg(X + (size_t)f16_id(e[j]) * DP + best_f) <= best_tf);
                    }
#pragma unroll
                    for (int j = 0; j < MT_PU; j++) {
                        if (j < nj) {
                            unsigned bal = __ballot_sync(F16_FULL, left[j]);
                            int lrank = __popc(bal & lt);
                            if (valid[j])
                                dst[start + (left[j] ? (run_l + lrank) : (n_left + (base + j * 32 - run_l) + (lane - lrank)))] = e[j];
                            run_l += __popc(bal);
                        }
                    }
                }
            }
            if (lane == 0) {
                F16StackRec q;
                q.parent = id; q.depth = r.depth + 1; q.n_const = (int16_t)n_total; q.const_mask = cmask; q.pad = 0;
                q.start = start + n_left; q.end = r.end; q.c0 = t0 - bl0; q.c1 = t1 - bl1; q.is_left = 0;
                stk.put(sp, q);
                q.start = start; q.end = start + n_left; q.c0 = bl0; q.c1 = bl1; q.is_left = 1;
                stk.put(sp + 1, q);
            }
            sp += 2;
        }
        __syncwarp();
    }
    if (lane == 0) { const int nc = min(node_count, P.node_cap); P.node_count[t] = nc; atomicMax(P.err + 1, nc); }
}

template <int DP>
__global__ void __launch_bounds__(NT, MT_MINB) k_build_random_mt(F16FitParams P) {
    constexpr int S = (DP == 16) ? MT_S16 : MT_S8;
    extern __shared__ __align__(16) unsigned char mt_raw[];
    using Tree = MtTree<DP, S>;
    Tree* trees = reinterpret_cast<Tree*>(mt_raw);
    MtScratch& hs = *reinterpret_cast<MtScratch*>(mt_raw + sizeof(Tree) * MT_TW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int first = blockIdx.x * MT_TW;
    if (tid == 0) { hs.done = 0; hs.pick = -1; }
    if (tid < MT_TW) trees[tid].req.state = 0;
    __syncthreads();
    if (warp < MT_TW) {
        const int t = first + warp;
        if (t < P.n_trees) mt_tree_warp<DP, S>(P, t, trees[warp], lane);
        __syncwarp();
        if (lane == 0) { __threadfence_block(); atomicAdd(const_cast<int*>(&hs.done), 1); }
    } else {
        mt_helper_loop<DP, S>(P, trees, hs, tid - MT_TW * 32);
    }
}

template <int DP> static size_t mt_smem_bytes() {
    constexpr int S = (DP == 16) ? MT_S16 : MT_S8;
    return sizeof(MtTree<DP, S>) * MT_TW + sizeof(MtScratch);
}

int f16_launch_build_random_mt(const F16FitParams& P, cudaStream_t st) {
    const int grid = (P.n_trees + MT_TW - 1) / MT_TW;
    cudaError_t e;
    if (P.dp == 8) {
        const size_t smem = mt_smem_bytes<8>();
        e = cudaFuncSetAttribute(k_build_random_mt<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return F16_ERR_CUDA;
        k_build_random_mt<8><<<grid, NT, smem, st>>>(P);
    } else {
        const size_t smem = mt_smem_bytes<16>();
        e = cudaFuncSetAttribute(k_build_random_mt<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return F16_ERR_CUDA;
        k_build_random_mt<16><<<grid, NT, smem, st>>>(P);
    }
    f16_count_launch(1);
    return cudaGetLastError() == cudaSuccess ? F16_OK : F16_ERR_CUDA;
}
