// ExtraTrees: DepthFirstTreeBuilder + RandomSplitter + Gini, one CTA per tree.
//   node_split_random      sklearn/tree/_splitter.pyx:507-736
//   find_min_max / partition_samples(_final)   sklearn/tree/_partitioner.pyx:129-167, :217-279
//   builder loop           sklearn/tree/_tree.pyx:139-336
// (reference call site: ExtraTreesClassifier(random_state=0).fit, experiment.py:96,469)
//
// Two regimes, switched per node by its row count:
//   * GLOBAL (n > S): the node's packed row ids live in a per-tree global index array; the CTA
//     makes three coalesced sweeps with gathers from the L2-resident row matrix - min/max of
//     all features (one float4 per thread), left-count of the <= 4 candidate thresholds,
//     stable out-of-place partition (ping-pong by depth parity).
//   * SHARED (n <= S): the node's rows are copied ONCE into shared memory (column-major,
//     odd stride => conflict-free) and the whole subtree is grown from there by warp 0 alone:
//     no block barriers, no global loads, only node records are stored to HBM.  The median
//     internal node has a few dozen rows, so most nodes of a tree take this path.
// The xorshift stream, the feature permutation and the node numbering are shared by both
// regimes, so the tree is identical to the CPU one regardless of where a node is processed.
// Compile with -fmad=false.
#include "f16_tree_dev.cuh"
#include "f16_tree_random_sub.cuh"

template <int DP> struct SubCfg;
#ifndef F16_S16
#define F16_S16 512
#endif
#ifndef F16_S8
#define F16_S8 1024
#endif
template <> struct SubCfg<16> { static constexpr int S = F16_S16; };
template <> struct SubCfg<8> { static constexpr int S = F16_S8; };
#define k_build_random F16_CAT(k_build_random, F16_VARIANT)

// ------------------------------------------------------------------ kernel
template <int DP>
__global__ void __launch_bounds__(NT, F16_MINB) k_build_random(F16FitParams P) {
    constexpr int Q = DP / 4;
    constexpr int SPI = NT / Q;
    constexpr int S = SubCfg<DP>::S;
    constexpr int SP = S + 1;
    __shared__ Ctl c;
    __shared__ DrawState ds;
    __shared__ F16StackRec s_stack[SSTK];
    __shared__ float s_min[F16_MAX_D], s_max[F16_MAX_D];
    __shared__ float s_wmin[NW][F16_MAX_D], s_wmax[NW][F16_MAX_D];
    __shared__ int s_cand_f[F16_MAX_D];
    __shared__ double s_cand_thr[F16_MAX_D];
    __shared__ unsigned long long s_part[NW][4];
    __shared__ unsigned long long s_cnt[F16_MAX_D];
    __shared__ int s_wcnt[2][PU][NW];
    __shared__ float s_col[DP * SP];
    __shared__ uint16_t s_idx[2][S];
    __shared__ uint8_t s_y[S];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lw = f16_leader_warp(P.salt);                 // the CTA's scalar / subtree warp
    const int rtid = (tid - lw * 32) & (NT - 1);            // rotated: leader warp's lane 0 is 0
    const int t = blockIdx.x;
    const int n = P.n;
    const float* __restrict__ X = P.X;
    uint32_t* buf0 = P.buf + (size_t)t * 2 * n;
    uint32_t* buf1 = buf0 + n;
    F16Node* nodes = P.nodes + (size_t)t * P.node_cap;
    TreeStack stk;
    stk.smem = s_stack;
    stk.gmem = P.stack + (size_t)t * P.stack_cap;
    const double W_total = (double)n;
    PH_DECL

    // ---- root: identity sample list with packed labels; class counts
    unsigned long long cnt = 0;
    for (int i = tid; i < n; i += NT) {
        uint32_t y = P.y[i];
        buf0[i] = f16_pack((uint32_t)i, 1u, y);
        cnt += y ? (1ull << 32) : 1ull;
    }
    cnt = f16_warp_sum_u64(cnt);
    if (lane == 0) s_part[warp][0] = cnt;
    __syncthreads();
    if (rtid == 0) {
        unsigned long long tot = 0;
        for (int q = 0; q < NW; q++) tot += s_part[q][0];
        for (int f = 0; f < F16_MAX_D; f++) { ds.features[f] = f; ds.const_feats[f] = 0; }
        ds.rng = P.rand_r_state[t];
        F16StackRec r;
        r.start = 0; r.end = n; r.parent = -1; r.c0 = (int)(uint32_t)tot; r.c1 = (int)(tot >> 32);
        r.const_mask = 0; r.n_const = 0; r.is_left = 0; r.pad = 0; r.depth = 0;
        c.sp = 0; c.node_count = 0; c.done = 0; c.abort = 0; c.split = 0;
        stk.put(c.sp++, r);
    }
    __syncthreads();
    PH_T(0, 0);

    while (true) {
        if (rtid == 0) pop_node(c, stk);
        __syncthreads();
        if (c.done) break;
        const int start = c.start, nn = c.end - c.start;
        const uint32_t* src = (c.depth & 1) ? buf1 : buf0;
        uint32_t* dst = (c.depth & 1) ? buf0 : buf1;

        if (!c.leaf && nn <= S) {
            // ================= SHARED regime: relocate the node, grow its whole subtree
            {
                const int q = tid % Q;
                for (int i = tid / Q; i < nn; i += SPI) {
                    uint32_t e = src[start + i];
                    float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)f16_id(e) * DP) + q);
                    s_col[(q * 4 + 0) * SP + i] = v.x;
                    s_col[(q * 4 + 1) * SP + i] = v.y;
                    s_col[(q * 4 + 2) * SP + i] = v.z;
                    s_col[(q * 4 + 3) * SP + i] = v.w;
                    if (q == 0) { s_idx[c.depth & 1][i] = (uint16_t)i; s_y[i] = (uint8_t)f16_y(e); }
                }
            }
            if (rtid == 0) {
                F16StackRec r;
                r.start = 0; r.end = nn; r.parent = c.parent; r.c0 = c.c0; r.c1 = c.c1;
                r.const_mask = c.const_mask; r.n_const = (int16_t)c.n_const; r.is_left = (uint8_t)c.is_left;
                r.pad = 1; r.depth = c.depth;
                stk.put(c.sp++, r);
            }
            __syncthreads();
            PH_T(0, 1);
            if (warp == lw) subtree_warp_v2<DP, S>(c, ds, stk, P, nodes, s_col, s_idx, s_y);
            __syncthreads();
            PH_T(0, 5);
            if (c.abort) break;
            continue;
        }

        // ================= GLOBAL regime
        if (!c.leaf) {
            // ---- pass 1: min / max of every feature over the node's rows
            {
                const int q = tid % Q, sl = tid / Q;
                float mn[4], mx[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
                // the row ids of the NEXT step are loaded while this step's row gathers are in flight
                // (ids -> rows is a chain of two memory latencies; overlapped, a step costs one)
                uint32_t idn[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { int i = sl + u * SPI; idn[u] = (sl < nn) ? f16_id(src[start + (i < nn ? i : sl)]) : 0u; }
                for (int i0 = sl; i0 < nn; i0 += SPI * 4) {
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)idn[u] * DP) + q);
                    const int i1 = i0 + SPI * 4;
                    if (i1 < nn) {
#pragma unroll
                        for (int u = 0; u < 4; u++) { int i = i1 + u * SPI; idn[u] = f16_id(src[start + (i < nn ? i : i1)]); }
                    }
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
                    for (int j = 0; j < 4; j++) { s_wmin[warp][lane * 4 + j] = mn[j]; s_wmax[warp][lane * 4 + j] = mx[j]; }
                }
                __syncthreads();
                if (rtid < DP) {
                    float a = s_wmin[0][rtid], b = s_wmax[0][rtid];
#pragma unroll
                    for (int q2 = 1; q2 < NW; q2++) { a = fminf(a, s_wmin[q2][rtid]); b = fmaxf(b, s_wmax[q2][rtid]); }
                    s_min[rtid] = a; s_max[rtid] = b;
                }
                __syncthreads();
            }
            // ---- draw features + thresholds (thread 0; scalar xorshift stream)
            if (rtid == 0) {
                const int d = P.d, max_features = P.max_features;
                int f_i = d, n_visited = 0, n_found = 0, n_drawn = 0;
                const int n_known = c.n_const;
                int n_total = n_known, ncand = 0;
                while (f_i > n_total && (n_visited < max_features || n_visited <= n_found + n_drawn)) {
                    n_visited++;
                    int f_j = f16_rand_int(n_drawn, f_i - n_found, &ds.rng);
                    if (f_j < n_known) {
                        int tmp = ds.features[n_drawn]; ds.features[n_drawn] = ds.features[f_j]; ds.features[f_j] = tmp;
                        n_drawn++;
                        continue;
                    }
                    f_j += n_found;
                    int f = ds.features[f_j];
                    float mn = s_min[f], mx = s_max[f];
                    if (mx <= __fadd_rn(mn, 1e-7f)) {
                        ds.features[f_j] = ds.features[n_total]; ds.features[n_total] = f;
                        n_found++; n_total++;
                        continue;
                    }
                    f_i--;
                    { int tmp = ds.features[f_i]; ds.features[f_i] = ds.features[f_j]; ds.features[f_j] = tmp; }
                    double thr = f16_rand_uniform((double)mn, (double)mx, &ds.rng);
                    if (thr == (double)mx) thr = (double)mn;
                    s_cand_f[ncand] = f; s_cand_thr[ncand] = thr; ncand++;
                }
                for (int i = 0; i < n_known; i++) ds.features[i] = ds.const_feats[i];
                uint32_t m = c.const_mask;
                for (int i = n_known; i < n_total; i++) { ds.const_feats[i] = ds.features[i]; m |= 1u << ds.features[i]; }
                c.ncand = ncand; c.n_const_out = n_total; c.const_mask_out = m;
            }
            __syncthreads();
            PH_T(0, 2);
            const int ncand = c.ncand;
            if (ncand > 0) {
                // ---- pass 2: left counts of every candidate threshold, 4 candidates per sweep.
                //      Four threads share a row: thread t loads the 16-byte quarter of the row that
                //      holds candidate t's feature, so a warp's load touches 8 rows = 8 lines (four
                //      scalar gathers per row cost four times the L1 wavefronts).  `x <= thr` for a
                //      float32 x and a float64 thr is decided in float32 against thr rounded down.
                //      With one or two candidates in the chunk (max_features = 2 for the 7-feature
                //      sets, or constant features) one / two threads share a row instead of four, so no
                //      thread repeats another's work: 1 << sh threads per row.
                for (int k0 = 0; k0 < ncand; k0 += 4) {
                    const int nck = min(4, ncand - k0);
                    const int sh = (nck == 1) ? 0 : (nck == 2) ? 1 : 2;
                    const int tr = tid & ((1 << sh) - 1);
                    const int k = (k0 + tr < ncand) ? k0 + tr : k0;
                    const int fk = s_cand_f[k];
                    const float tk = __double2float_rd(s_cand_thr[k]);
                    const int qk = fk >> 2, ck = fk & 3;
                    unsigned long long acc = 0;
                    const int RPS = NT >> sh;                   // rows per sweep step
                    uint32_t en[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { int i = (tid >> sh) + u * RPS; en[u] = (i < nn) ? src[start + i] : 0xffffffffu; }
                    for (int i0 = tid >> sh; i0 < nn; i0 += RPS * 4) {
                        uint32_t e[4]; float4 v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) e[u] = en[u];
#pragma unroll
                        for (int u = 0; u < 4; u++)
                            v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)f16_id(e[u] == 0xffffffffu ? 0u : e[u]) * DP) + qk);
                        const int i1 = i0 + RPS * 4;
                        if (i1 < nn) {             // next step's entries, while the gathers are in flight
#pragma unroll
                            for (int u = 0; u < 4; u++) { int i = i1 + u * RPS; en[u] = (i < nn) ? src[start + i] : 0xffffffffu; }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const float x = ck == 0 ? v[u].x : ck == 1 ? v[u].y : ck == 2 ? v[u].z : v[u].w;
                            if (e[u] != 0xffffffffu && x <= tk) acc += 1ull | ((unsigned long long)f16_y(e[u]) << 32);
                        }
                    }
                    // lanes with the same tr hold the same candidate: reduce over the others
                    if (sh < 1) acc += __shfl_xor_sync(F16_FULL, acc, 1);
                    if (sh < 2) acc += __shfl_xor_sync(F16_FULL, acc, 2);
                    acc += __shfl_xor_sync(F16_FULL, acc, 4);
                    acc += __shfl_xor_sync(F16_FULL, acc, 8);
                    acc += __shfl_xor_sync(F16_FULL, acc, 16);
                    if (lane < (1 << sh)) s_part[warp][lane] = acc;
                    __syncthreads();
                    if (rtid < (1 << sh) && k0 + rtid < ncand) {
                        unsigned long long s = 0;
                        for (int q2 = 0; q2 < NW; q2++) s += s_part[q2][rtid];
                        s_cnt[k0 + rtid] = s;
                    }
                    __syncthreads();
                }
                // ---- choose the best candidate (strict >, first wins)
                if (rtid == 0) {
                    double best = -INFINITY; int bk = -1;
                    for (int k = 0; k < ncand; k++) {
                        int nl = (int)(uint32_t)s_cnt[k], l1 = (int)(s_cnt[k] >> 32), l0 = nl - l1;
                        double proxy = gini_proxy(l0, l1, c.c0, c.c1);
                        if (proxy > best) { best = proxy; bk = k; }
                    }
                    if (bk >= 0) {
                        int nl = (int)(uint32_t)s_cnt[bk], l1 = (int)(s_cnt[bk] >> 32), l0 = nl - l1;
                        c.best_f = s_cand_f[bk]; c.best_thr = s_cand_thr[bk]; c.node_id_k = bk;
                        c.n_left = nl; c.l0 = l0; c.l1 = l1;
                        c.split = (improvement_certain(c.c0, c.c1, n) || improvement_ok(l0, l1, c.c0, c.c1, W_total)) ? 1 : 0;
                    }
                }
            }
        }
        if (rtid == 0) finish_node(c, P, nodes, stk);
        __syncthreads();
        PH_T(0, c.leaf ? 6 : 3);
        if (c.abort) break;
        if (c.split) {
            const int bf = c.best_f; const float bthr = __double2float_rd(c.best_thr);
            // (storing the candidate sweep's comparison bits per row and partitioning from them,
            //  without this gather, was measured 15 % SLOWER: extra byte stores + register spills)
            block_partition(src, dst, start, nn, c.n_left,
                            [&](uint32_t e, int) { return __ldg(X + (size_t)f16_id(e) * DP + bf) <= bthr; }, s_wcnt);
            __syncthreads();
            if (rtid == 0) c.split = 0;
            PH_T(0, 4);
        }
    }
    if (rtid == 0) { const int nc = min(c.node_count, P.node_cap); P.node_count[t] = nc; atomicMax(P.err + 1, nc); }
}

F16_PHASE_READER(F16_CAT(f16_debug_phases, F16_VARIANT))

int F16_CAT(f16_launch_build_random, F16_VARIANT)(const F16FitParams& P, cudaStream_t st) {
    if (P.dp == 8) k_build_random<8><<<P.n_trees, NT, 0, st>>>(P);
    else k_build_random<16><<<P.n_trees, NT, 0, st>>>(P);
    f16_count_launch(1);
    return cudaGetLastError() == cudaSuccess ? F16_OK : F16_ERR_CUDA;
}
