// Tree growing for ExtraTrees / RandomForest / DecisionTree, one CTA per tree.
//
// What is rebuilt here:
//   DepthFirstTreeBuilder.build      sklearn/tree/_tree.pyx:139-336
//   node_split_best / _random        sklearn/tree/_splitter.pyx:262-504 / :507-736
//   DensePartitioner                 sklearn/tree/_partitioner.pyx
//   Gini criterion                   sklearn/tree/_criterion.pyx:147-199, :622-687
//   bootstrap / per-tree seeds       sklearn/ensemble/_forest.py:94-112,150-166; _base.py
// called by the reference at experiment.py:96-98 (construction) and :469 (fit).
//
// Design (B200): the per-tree xorshift stream is consumed in depth-first node order, so
// the nodes of ONE tree are inherently sequential; parallelism comes from trees x folds x
// configs (one CTA per tree, many forests in flight on separate streams) and from the rows
// inside a node (all threads of the CTA).  The row matrix (<= 11 MB) stays L2 resident;
// per-tree index arrays are streamed with coalesced loads and partitioned out of place
// between two ping-pong buffers selected by node depth parity.
//   * random splitter (ET): 3 passes per node - min/max of all features (one float4 per
//     thread), left-count of <=4 candidate thresholds, stable partition.
//   * best splitter (RF, DT): per-feature index arrays kept SORTED by feature value and
//     stably partitioned at every split, so a node's candidate scan is one linear pass
//     (warp scan of packed class weights); no per-node sort.  A warp owns one feature.
// Compile with -fmad=false: the float64 criterion expressions must round like the CPU.
#include "f16_tree.cuh"
#include <math.h>

#define NT 256
#define NW (NT / 32)
#define F16_EPS 2.220446049250313e-16

// ------------------------------------------------------------------ bootstrap (RF)
// One CTA per tree: MT19937(tree_seed).randint(0, n, n) -> bincount, entirely on device.
// The 624-word state block is regenerated in three dependency-free phases.
__global__ void __launch_bounds__(NT) k_bootstrap(const uint32_t* __restrict__ tree_seed, int n,
                                                 uint32_t* __restrict__ w32 /*[n_trees][ceil(n/4)]*/,
                                                 int words_per_tree) {
    __shared__ uint32_t mt[624];
    __shared__ uint32_t out[624];
    __shared__ int s_wsum[2][NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* w = w32 + (size_t)blockIdx.x * words_per_tree;
    if (tid == 0) {
        uint32_t s = tree_seed[blockIdx.x];
        mt[0] = s;
        for (int i = 1; i < 624; i++) { s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i; mt[i] = s; }
    }
    __syncthreads();
    const uint32_t rng = (uint32_t)(n - 1);
    const uint32_t mask = f16_gen_mask(rng);
    int produced = 0, buf = 0;
    while (produced < n) {
        // ---- regenerate 624 words
        uint32_t a, b, c;
        if (tid < 227) { a = mt[tid]; b = mt[tid + 1]; c = mt[tid + 397]; }
        __syncthreads();
        if (tid < 227) mt[tid] = c ^ f16_mt_twist(a, b);
        __syncthreads();
        if (tid < 227) { int i = tid + 227; a = mt[i]; b = mt[i + 1]; c = mt[i - 227]; }
        __syncthreads();
        if (tid < 227) mt[tid + 227] = c ^ f16_mt_twist(a, b);
        __syncthreads();
        if (tid < 170) { int i = tid + 454; a = mt[i]; b = mt[(i + 1) % 624]; c = mt[i - 227]; }
        __syncthreads();
        if (tid < 170) mt[tid + 454] = c ^ f16_mt_twist(a, b);
        __syncthreads();
        for (int i = tid; i < 624; i += NT) out[i] = f16_mt_temper(mt[i]) & mask;
        __syncthreads();
        // ---- ordered acceptance: only the first (n - produced) accepted draws count
        for (int base = 0; base < 624 && produced < n; base += NT, buf ^= 1) {
            int i = base + tid;
            uint32_t v = (i < 624) ? out[i] : 0xffffffffu;
            bool acc = (i < 624) && (v <= rng);
            unsigned bal = __ballot_sync(F16_FULL, acc);
            if (lane == 0) s_wsum[buf][warp] = __popc(bal);
            __syncthreads();
            int before = 0, tot = 0;
            for (int q = 0; q < NW; q++) { int cq = s_wsum[buf][q]; if (q < warp) before += cq; tot += cq; }
            int rank = produced + before + __popc(bal & ((1u << lane) - 1u));
            if (acc && rank < n) atomicAdd(&w[v >> 2], 1u << ((v & 3u) * 8u));
            produced += tot;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ shared control block
struct Ctl {
    int start, end, parent, c0, c1, n_const, is_left, depth;
    uint32_t const_mask;
    int done, leaf, split;
    int ncand;
    int best_f;
    double best_thr;
    int n_left, l0, l1;
    int n_const_out;
    uint32_t const_mask_out;
    int node_id;
    unsigned long long win_key;
};

__device__ __forceinline__ double gini_of(double a, double b, double w) {
    double sq = 0.0;
    sq = sq + a * a;
    sq = sq + b * b;
    return 1.0 - sq / (w * w);
}

// proxy_impurity_improvement for Gini (sklearn/tree/_criterion.pyx:147-163, :647-687)
__device__ __forceinline__ double gini_proxy(int l0, int l1, int t0, int t1) {
    double L0 = (double)l0, L1 = (double)l1;
    double R0 = (double)(t0 - l0), R1 = (double)(t1 - l1);
    double wl = L0 + L1, wr = R0 + R1;
    double gl = gini_of(L0, L1, wl);
    double gr = gini_of(R0, R1, wr);
    return (-wr * gr) - wl * gl;
}

// impurity_improvement (sklearn/tree/_criterion.pyx:165-199) + the builder's
// `improvement + EPSILON < min_impurity_decrease` test (_tree.pyx:246-252), min_dec = 0.
__device__ __forceinline__ bool improvement_ok(int l0, int l1, int t0, int t1, double W_total) {
    double L0 = (double)l0, L1 = (double)l1;
    double R0 = (double)(t0 - l0), R1 = (double)(t1 - l1);
    double wl = L0 + L1, wr = R0 + R1, wn = (double)t0 + (double)t1;
    double imp = gini_of((double)t0, (double)t1, wn);
    double gl = gini_of(L0, L1, wl);
    double gr = gini_of(R0, R1, wr);
    double improvement = (wn / W_total) * (imp - (wr / wn * gr) - (wl / wn * gl));
    return !(improvement + F16_EPS < 0.0);
}

// thread 0: write node, link to parent, push children
__device__ __forceinline__ void finish_node(Ctl& c, const F16FitParams& P, F16Node* nodes, F16StackRec* stack,
                                            int& sp, int& node_count) {
    int id = node_count++;
    if (id >= P.node_cap) { atomicExch(P.err, F16_ERR_OVERFLOW); c.done = 1; return; }
    F16Node nd;
    nd.thr = c.split ? c.best_thr : -2.0;
    nd.feature = c.split ? c.best_f : -2;
    nd.right = -1;
    nd.c0 = c.c0; nd.c1 = c.c1; nd.n = c.end - c.start; nd.depth = c.depth;
    nodes[id] = nd;
    if (c.parent >= 0 && !c.is_left) nodes[c.parent].right = id;
    c.node_id = id;
    if (c.split) {
        if (sp + 2 > P.stack_cap) { atomicExch(P.err, F16_ERR_OVERFLOW); c.done = 1; return; }
        F16StackRec r;
        r.parent = id; r.depth = c.depth + 1; r.n_const = (int16_t)c.n_const_out;
        r.const_mask = c.const_mask_out; r.pad = 0;
        // right child first so that the left child is popped first
        r.start = c.start + c.n_left; r.end = c.end; r.c0 = c.c0 - c.l0; r.c1 = c.c1 - c.l1; r.is_left = 0;
        stack[sp++] = r;
        r.start = c.start; r.end = c.start + c.n_left; r.c0 = c.l0; r.c1 = c.l1; r.is_left = 1;
        stack[sp++] = r;
    }
}

// thread 0: pop + the builder's leaf pre-test (_tree.pyx:223-240)
__device__ __forceinline__ void pop_node(Ctl& c, F16StackRec* stack, int& sp) {
    if (sp == 0) { c.done = 1; return; }
    F16StackRec r = stack[--sp];
    c.start = r.start; c.end = r.end; c.parent = r.parent; c.c0 = r.c0; c.c1 = r.c1;
    c.n_const = r.n_const; c.const_mask = r.const_mask; c.is_left = r.is_left; c.depth = r.depth;
    double wn = (double)r.c0 + (double)r.c1;
    double imp = gini_of((double)r.c0, (double)r.c1, wn);
    c.leaf = ((r.end - r.start) < 2) || (imp <= F16_EPS);
    c.split = 0; c.ncand = 0;
    c.n_const_out = r.n_const; c.const_mask_out = r.const_mask;
}

// Fisher-Yates feature draw shared by both splitters (_splitter.pyx:332-389 / :573-626).
// `is_const(f)` is consulted only for features that are not known constants.
// Returns the list of features to evaluate, in visit order.
struct DrawState {
    int features[F16_MAX_D];
    int const_feats[F16_MAX_D];
    uint32_t rng;
};

// ------------------------------------------------------------------ stable block partition (one array)
template <class Pred>
__device__ __forceinline__ void block_partition(const uint32_t* src, uint32_t* dst, int start, int n, int n_left,
                                                Pred pred, int (*s_wcnt)[NW]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int run_l = 0, buf = 0;
    for (int base = 0; base < n; base += NT, buf ^= 1) {
        int p = base + tid;
        bool valid = p < n;
        uint32_t e = valid ? src[start + p] : 0u;
        bool left = valid && pred(e);
        unsigned bal = __ballot_sync(F16_FULL, left);
        if (lane == 0) s_wcnt[buf][warp] = __popc(bal);
        __syncthreads();
        int before = 0, tot = 0;
#pragma unroll
        for (int q = 0; q < NW; q++) { int cq = s_wcnt[buf][q]; if (q < warp) before += cq; tot += cq; }
        int lrank = before + __popc(bal & ((1u << lane) - 1u));
        if (valid) {
            if (left) dst[start + run_l + lrank] = e;
            else dst[start + n_left + (base - run_l) + (tid - lrank)] = e;
        }
        run_l += tot;
    }
}

// ================================================================== RANDOM splitter (ExtraTrees)
template <int DP>
__global__ void __launch_bounds__(NT) k_build_random(F16FitParams P) {
    constexpr int Q = DP / 4;
    constexpr int SPI = NT / Q;
    __shared__ Ctl c;
    __shared__ DrawState ds;
    __shared__ float s_min[F16_MAX_D], s_max[F16_MAX_D];
    __shared__ float s_wmin[NW][F16_MAX_D], s_wmax[NW][F16_MAX_D];
    __shared__ int s_cand_f[F16_MAX_D];
    __shared__ double s_cand_thr[F16_MAX_D];
    __shared__ unsigned long long s_part[NW][4];
    __shared__ unsigned long long s_cnt[F16_MAX_D];
    __shared__ int s_wcnt[2][NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    const int n = P.n;
    const float* __restrict__ X = P.X;
    uint32_t* buf0 = P.buf + (size_t)t * 2 * n;
    uint32_t* buf1 = buf0 + n;
    F16Node* nodes = P.nodes + (size_t)t * P.node_cap;
    F16StackRec* stack = P.stack + (size_t)t * P.stack_cap;
    const double W_total = (double)n;

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
    int sp = 0, node_count = 0;
    if (tid == 0) {
        unsigned long long tot = 0;
        for (int q = 0; q < NW; q++) tot += s_part[q][0];
        for (int f = 0; f < F16_MAX_D; f++) { ds.features[f] = f; ds.const_feats[f] = 0; }
        ds.rng = P.rand_r_state[t];
        F16StackRec r;
        r.start = 0; r.end = n; r.parent = -1; r.c0 = (int)(uint32_t)tot; r.c1 = (int)(tot >> 32);
        r.const_mask = 0; r.n_const = 0; r.is_left = 0; r.pad = 0; r.depth = 0;
        stack[sp++] = r;
        c.done = 0;
    }
    __syncthreads();

    while (true) {
        if (tid == 0) pop_node(c, stack, sp);
        __syncthreads();
        if (c.done) break;
        const int start = c.start, nn = c.end - c.start;
        const uint32_t* src = (c.depth & 1) ? buf1 : buf0;
        uint32_t* dst = (c.depth & 1) ? buf0 : buf1;

        if (!c.leaf) {
            // ---- pass 1: min / max of every feature over the node's rows
            {
                const int q = tid % Q, sl = tid / Q;
                float mn[4], mx[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
                for (int i = sl; i < nn; i += SPI) {
                    uint32_t id = f16_id(src[start + i]);
                    float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)id * DP) + q);
                    mn[0] = fminf(mn[0], v.x); mx[0] = fmaxf(mx[0], v.x);
                    mn[1] = fminf(mn[1], v.y); mx[1] = fmaxf(mx[1], v.y);
                    mn[2] = fminf(mn[2], v.z); mx[2] = fmaxf(mx[2], v.z);
                    mn[3] = fminf(mn[3], v.w); mx[3] = fmaxf(mx[3], v.w);
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
                if (tid < DP) {
                    float a = s_wmin[0][tid], b = s_wmax[0][tid];
#pragma unroll
                    for (int q2 = 1; q2 < NW; q2++) { a = fminf(a, s_wmin[q2][tid]); b = fmaxf(b, s_wmax[q2][tid]); }
                    s_min[tid] = a; s_max[tid] = b;
                }
                __syncthreads();
            }
            // ---- draw features + thresholds (thread 0; scalar xorshift stream)
            if (tid == 0) {
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
            const int ncand = c.ncand;
            if (ncand > 0) {
                // ---- pass 2: left counts of every candidate threshold (4 candidates per sweep)
                for (int k0 = 0; k0 < ncand; k0 += 4) {
                    int fk[4]; double tk[4];
                    unsigned long long acc[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        int k = (k0 + j < ncand) ? k0 + j : k0;
                        fk[j] = s_cand_f[k]; tk[j] = s_cand_thr[k]; acc[j] = 0;
                    }
                    for (int i = tid; i < nn; i += NT) {
                        uint32_t e = src[start + i];
                        const float* row = X + (size_t)f16_id(e) * DP;
                        unsigned long long one = 1ull | ((unsigned long long)f16_y(e) << 32);
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            float v = __ldg(row + fk[j]);
                            if ((double)v <= tk[j]) acc[j] += one;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[j] = f16_warp_sum_u64(acc[j]);
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 4; j++) s_part[warp][j] = acc[j];
                    }
                    __syncthreads();
                    if (tid < 4 && k0 + tid < ncand) {
                        unsigned long long s = 0;
                        for (int q2 = 0; q2 < NW; q2++) s += s_part[q2][tid];
                        s_cnt[k0 + tid] = s;
                    }
                    __syncthreads();
                }
                // ---- choose the best candidate (strict >, first wins)
                if (tid == 0) {
                    double best = -INFINITY; int bk = -1;
                    for (int k = 0; k < ncand; k++) {
                        int nl = (int)(uint32_t)s_cnt[k], l1 = (int)(s_cnt[k] >> 32), l0 = nl - l1;
                        double proxy = gini_proxy(l0, l1, c.c0, c.c1);
                        if (proxy > best) { best = proxy; bk = k; }
                    }
                    if (bk >= 0) {
                        int nl = (int)(uint32_t)s_cnt[bk], l1 = (int)(s_cnt[bk] >> 32), l0 = nl - l1;
                        c.best_f = s_cand_f[bk]; c.best_thr = s_cand_thr[bk];
                        c.n_left = nl; c.l0 = l0; c.l1 = l1;
                        c.split = improvement_ok(l0, l1, c.c0, c.c1, W_total) ? 1 : 0;
                    }
                }
            }
        }
        if (tid == 0) finish_node(c, P, nodes, stack, sp, node_count);
        __syncthreads();
        if (c.done) break;
        if (c.split) {
            const int bf = c.best_f; const double bthr = c.best_thr;
            block_partition(src, dst, start, nn, c.n_left,
                            [&](uint32_t e) { return (double)__ldg(X + (size_t)f16_id(e) * DP + bf) <= bthr; }, s_wcnt);
            __syncthreads();
        }
    }
    if (tid == 0) P.node_count[t] = node_count;
}

// ================================================================== BEST splitter (RF, DT)
struct BestCand {
    double proxy;
    unsigned long long key;   // (visit order k << 32) | position p ; smaller wins ties
    float v_prev, v;
    int l0, l1;
};

__device__ __forceinline__ bool side_get(const uint32_t* side, uint32_t id) { return (side[id >> 5] >> (id & 31)) & 1u; }

// warp-cooperative stable partition of one sorted feature array by the side bits
__device__ __forceinline__ void warp_partition(const uint32_t* src, uint32_t* dst, int start, int n, int n_left,
                                               const uint32_t* side) {
    const int lane = threadIdx.x & 31;
    int run_l = 0;
    for (int base = 0; base < n; base += 128) {
        uint32_t e[4]; bool valid[4], left[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int p = base + j * 32 + lane;
            valid[j] = p < n;
            e[j] = valid[j] ? src[start + p] : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) left[j] = valid[j] && side_get(side, f16_id(e[j]));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            unsigned bal = __ballot_sync(F16_FULL, left[j]);
            int lrank = __popc(bal & ((1u << lane) - 1u));
            int pb = base + j * 32;
            if (valid[j]) {
                if (left[j]) dst[start + run_l + lrank] = e[j];
                else dst[start + n_left + (pb - run_l) + (lane - lrank)] = e[j];
            }
            run_l += __popc(bal);
        }
    }
}

template <int DP>
__global__ void __launch_bounds__(NT) k_build_best(F16FitParams P) {
    extern __shared__ uint32_t s_side_dyn[];
    __shared__ Ctl c;
    __shared__ DrawState ds;
    __shared__ float s_min[F16_MAX_D], s_max[F16_MAX_D];
    __shared__ int s_eval_f[F16_MAX_D];
    __shared__ double s_bproxy[NW];
    __shared__ unsigned long long s_bkey[NW];
    __shared__ unsigned long long s_part[NW];
    __shared__ int s_cntn[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;
    const int n = P.n, d = P.d;
    const float* __restrict__ X = P.X;
    uint32_t* ord0 = P.buf + (size_t)t * 2 * d * n;   // [d][n]
    uint32_t* ord1 = ord0 + (size_t)d * n;
    F16Node* nodes = P.nodes + (size_t)t * P.node_cap;
    F16StackRec* stack = P.stack + (size_t)t * P.stack_cap;
    uint32_t* side = P.side_global ? P.side_global + (size_t)t * P.side_words : s_side_dyn;
    const uint8_t* bw = P.boot_w ? P.boot_w + (size_t)t * (((size_t)n + 3) / 4 * 4) : nullptr;
    const double W_total = (double)n;   // sum of bootstrap counts == n; unit weights == n

    // ---- root: per feature, compact the column argsort to the in-bag rows (weight > 0),
    //      packing (id, weight, label).  A warp owns a feature.
    int n_root = 0;
    for (int f = warp; f < d; f += NW) {
        const int32_t* sidx = P.sorted_idx + (size_t)f * n;
        uint32_t* o = ord0 + (size_t)f * n;
        int run = 0;
        unsigned long long cls = 0;
        for (int base = 0; base < n; base += 32) {
            int p = base + lane;
            bool valid = p < n;
            uint32_t id = valid ? (uint32_t)sidx[p] : 0u;
            uint32_t w = valid ? (bw ? (uint32_t)bw[id] : 1u) : 0u;
            uint32_t y = valid ? (uint32_t)P.y[id] : 0u;
            if (w > F16_MAX_W) { atomicExch(P.err, F16_ERR_OVERFLOW); w = F16_MAX_W; }
            bool keep = w > 0;
            unsigned bal = __ballot_sync(F16_FULL, keep);
            if (keep) {
                o[run + __popc(bal & ((1u << lane) - 1u))] = f16_pack(id, w, y);
                cls += (unsigned long long)w << (y ? 32 : 0);
            }
            run += __popc(bal);
        }
        if (f == 0) {
            cls = f16_warp_sum_u64(cls);
            if (lane == 0) { s_part[0] = cls; s_cntn[0] = run; }
        }
    }
    __syncthreads();
    n_root = s_cntn[0];
    int sp = 0, node_count = 0;
    if (tid == 0) {
        unsigned long long tot = s_part[0];
        for (int f = 0; f < F16_MAX_D; f++) { ds.features[f] = f; ds.const_feats[f] = 0; }
        ds.rng = P.rand_r_state[t];
        F16StackRec r;
        r.start = 0; r.end = n_root; r.parent = -1; r.c0 = (int)(uint32_t)tot; r.c1 = (int)(tot >> 32);
        r.const_mask = 0; r.n_const = 0; r.is_left = 0; r.pad = 0; r.depth = 0;
        stack[sp++] = r;
        c.done = 0;
    }
    __syncthreads();

    while (true) {
        if (tid == 0) pop_node(c, stack, sp);
        __syncthreads();
        if (c.done) break;
        const int start = c.start, nn = c.end - c.start;
        const uint32_t* src = (c.depth & 1) ? ord1 : ord0;
        uint32_t* dst = (c.depth & 1) ? ord0 : ord1;
        int n_eval = 0;

        if (!c.leaf) {
            // ---- min / max of every not-yet-constant feature: ends of its sorted slice
            if (tid < d && !((c.const_mask >> tid) & 1u)) {
                const uint32_t* o = src + (size_t)tid * n;
                s_min[tid] = __ldg(X + (size_t)f16_id(o[start]) * DP + tid);
                s_max[tid] = __ldg(X + (size_t)f16_id(o[start + nn - 1]) * DP + tid);
            }
            __syncthreads();
            // ---- Fisher-Yates feature draw (thread 0)
            if (tid == 0) {
                const int max_features = P.max_features;
                int f_i = d, n_visited = 0, n_found = 0, n_drawn = 0;
                const int n_known = c.n_const;
                int n_total = n_known, ne = 0;
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
                    if (s_max[f] <= __fadd_rn(s_min[f], 1e-7f)) {
                        ds.features[f_j] = ds.features[n_total]; ds.features[n_total] = f;
                        n_found++; n_total++;
                        continue;
                    }
                    f_i--;
                    { int tmp = ds.features[f_i]; ds.features[f_i] = ds.features[f_j]; ds.features[f_j] = tmp; }
                    s_eval_f[ne++] = f;
                }
                for (int i = 0; i < n_known; i++) ds.features[i] = ds.const_feats[i];
                uint32_t m = c.const_mask;
                for (int i = n_known; i < n_total; i++) { ds.const_feats[i] = ds.features[i]; m |= 1u << ds.features[i]; }
                c.ncand = ne; c.n_const_out = n_total; c.const_mask_out = m;
            }
            __syncthreads();
            n_eval = c.ncand;
            // ---- candidate scan: warp w owns evaluated features w, w+NW, ...
            BestCand best;
            best.proxy = -INFINITY; best.key = ~0ull; best.v_prev = 0.f; best.v = 0.f; best.l0 = 0; best.l1 = 0;
            const int t0 = c.c0, t1 = c.c1;
            for (int k = warp; k < n_eval; k += NW) {
                const int f = s_eval_f[k];
                const uint32_t* o = src + (size_t)f * n + start;
                unsigned long long carry = 0;
                float prev_last = 0.f;
                for (int base = 0; base < nn; base += 32) {
                    int p = base + lane;
                    bool valid = p < nn;
                    uint32_t e = valid ? o[p] : 0u;
                    float v = valid ? __ldg(X + (size_t)f16_id(e) * DP + f) : INFINITY;
                    unsigned long long my = valid ? ((unsigned long long)f16_w(e) << (f16_y(e) ? 32 : 0)) : 0ull;
                    unsigned long long incl = f16_warp_incl_scan_u64(my);
                    float vp = __shfl_up_sync(F16_FULL, v, 1);
                    if (lane == 0) vp = prev_last;
                    if (valid && p > 0 && v > __fadd_rn(vp, 1e-7f)) {
                        unsigned long long ex = carry + incl - my;
                        int l0 = (int)(uint32_t)ex, l1 = (int)(ex >> 32);
                        double proxy = gini_proxy(l0, l1, t0, t1);
                        if (proxy > best.proxy) {
                            best.proxy = proxy; best.key = ((unsigned long long)k << 32) | (unsigned)p;
                            best.v_prev = vp; best.v = v; best.l0 = l0; best.l1 = l1;
                        }
                    }
                    carry += __shfl_sync(F16_FULL, incl, 31);
                    prev_last = __shfl_sync(F16_FULL, v, 31);
                }
            }
            // ---- block arg-max with (k, p) tie order
            {
                double bp = best.proxy; unsigned long long bk = best.key;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    double op = __shfl_xor_sync(F16_FULL, bp, off);
                    unsigned long long ok = __shfl_xor_sync(F16_FULL, bk, off);
                    if (op > bp || (op == bp && ok < bk)) { bp = op; bk = ok; }
                }
                if (lane == 0) { s_bproxy[warp] = bp; s_bkey[warp] = bk; }
                __syncthreads();
                if (tid == 0) {
                    double wp = s_bproxy[0]; unsigned long long wk = s_bkey[0];
                    for (int q = 1; q < NW; q++)
                        if (s_bproxy[q] > wp || (s_bproxy[q] == wp && s_bkey[q] < wk)) { wp = s_bproxy[q]; wk = s_bkey[q]; }
                    c.win_key = (wp > -INFINITY) ? wk : ~0ull;
                }
                __syncthreads();
                if (c.win_key != ~0ull && best.key == c.win_key) {
                    int k = (int)(c.win_key >> 32), p = (int)(uint32_t)c.win_key;
                    c.best_f = s_eval_f[k];
                    c.best_thr = (double)best.v_prev / 2.0 + (double)best.v / 2.0;
                    c.n_left = p; c.l0 = best.l0; c.l1 = best.l1;
                    c.split = improvement_ok(best.l0, best.l1, t0, t1, W_total) ? 1 : 0;
                }
                __syncthreads();
            }
        }
        if (tid == 0) finish_node(c, P, nodes, stack, sp, node_count);
        __syncthreads();
        if (c.done) break;
        if (c.split) {
            // ---- mark the side of every row of the node (the winning feature's slice is
            //      sorted, so the left rows are its first n_left entries)
            const int n_left = c.n_left;
            const uint32_t* o = src + (size_t)c.best_f * n + start;
            for (int p = tid; p < nn; p += NT) {
                uint32_t id = f16_id(o[p]);
                if (p < n_left) atomicOr(&side[id >> 5], 1u << (id & 31));
                else atomicAnd(&side[id >> 5], ~(1u << (id & 31)));
            }
            __syncthreads();
            // ---- stable partition of every still-useful feature array, a warp per array
            const uint32_t keep_mask = ~c.const_mask_out;
            for (int f = warp; f < d; f += NW) {
                if (!((keep_mask >> f) & 1u)) continue;
                warp_partition(src + (size_t)f * n, dst + (size_t)f * n, start, nn, n_left, side);
            }
            __syncthreads();
        }
    }
    if (tid == 0) P.node_count[t] = node_count;
}

// ================================================================== predict
// walk: blockIdx.y = tree, threads over rows; leaf class sums -> leaf[tree][row]
template <int DP>
__global__ void __launch_bounds__(256) k_predict_walk(const F16Node* __restrict__ nodes, int node_cap,
                                                      const float* __restrict__ X, int n, int2* __restrict__ leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F16Node* tn = nodes + (size_t)blockIdx.y * node_cap;
    const float* row = X + (size_t)i * DP;
    int id = 0;
    while (true) {
        const int4* pn = reinterpret_cast<const int4*>(tn + id);
        int4 a = __ldg(pn);       // thr (2 words), feature, right
        if (a.z < 0) {
            int4 b = __ldg(pn + 1);
            leaf[(size_t)blockIdx.y * n + i] = make_int2(b.x, b.y);
            return;
        }
        double thr = __hiloint2double(a.y, a.x);
        float v = __ldg(row + a.z);
        id = ((double)v <= thr) ? id + 1 : a.w;
    }
}

// reduce: per row, sum tree probabilities IN TREE ORDER (sklearn/ensemble/_forest.py:704-717,
// :961-962), divide by n_trees, argmax with ties -> class 0 (:906).
__global__ void k_predict_reduce(const int2* __restrict__ leaf, int n, int n_trees, uint8_t* __restrict__ pred) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p0 = 0.0, p1 = 0.0;
    for (int t = 0; t < n_trees; t++) {
        int2 cc = leaf[(size_t)t * n + i];
        double w = (double)cc.x + (double)cc.y;
        p0 = p0 + (double)cc.x / w;
        p1 = p1 + (double)cc.y / w;
    }
    if (n_trees > 1) { p0 = p0 / (double)n_trees; p1 = p1 / (double)n_trees; }
    pred[i] = (p1 > p0) ? 1 : 0;
}

// ================================================================== host side
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" int f16_get_profiling(void);

#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

// host: forest seeds. tree_seed[i] = RandomState(seed).randint(2^31 - 1) (one draw per tree,
// in order; sklearn/ensemble/_base.py:_set_random_states), rand_r_state[i] =
// RandomState(tree_seed[i]).randint(0, 2^31 - 1) (sklearn/tree/_splitter.pyx:155).
// DecisionTree: tree_seed = seed itself (sklearn/tree/_classes.py check_random_state).
extern "C" int f16_tree_seeds(uint32_t seed, int32_t kind, int32_t n_trees, uint32_t* tree_seed,
                              uint32_t* rand_r_state) {
    if (n_trees < 1 || !tree_seed || !rand_r_state) { f16_set_error("f16_tree_seeds: bad arguments"); return F16_ERR_INVALID; }
    F16MT* mt = (F16MT*)malloc(sizeof(F16MT));
    if (kind == F16_KIND_DT) {
        tree_seed[0] = seed;
    } else {
        f16_mt_seed(mt, seed);
        for (int i = 0; i < n_trees; i++) tree_seed[i] = f16_mt_randint(mt, 2147483647u);
    }
    for (int i = 0; i < (kind == F16_KIND_DT ? 1 : n_trees); i++) {
        f16_mt_seed(mt, tree_seed[i]);
        rand_r_state[i] = f16_mt_randint(mt, 2147483647u);
    }
    free(mt);
    return F16_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int f16_bootstrap_counts(const uint32_t* tree_seed_host, int32_t n_trees, int64_t n, uint8_t* w_dev,
                                    void* stream) {
    // w_dev: [n_trees][align4(n)] bytes, zeroed here
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 1 || n > F16_MAX_ROWS) { f16_set_error("f16_bootstrap_counts: n out of range"); return F16_ERR_INVALID; }
    size_t stride = align_up((size_t)n, 4);
    uint32_t* seeds_dev = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&seeds_dev, sizeof(uint32_t) * n_trees, st));
    CUDA_TRY(cudaMemcpyAsync(seeds_dev, tree_seed_host, sizeof(uint32_t) * n_trees, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(w_dev, 0, stride * n_trees, st));
    k_bootstrap<<<n_trees, NT, 0, st>>>(seeds_dev, (int)n, (uint32_t*)w_dev, (int)(stride / 4));
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(seeds_dev, st));
    return F16_OK;
}

extern "C" int f16_forest_fit(const float* X_dev, const uint8_t* y_dev, int64_t n, int32_t d,
                              const int32_t* sorted_idx_dev, int32_t kind, int32_t n_estimators,
                              int32_t max_features, uint32_t seed, void* stream, f16_forest** out) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X_dev || !y_dev || !out || n < 1 || n > F16_MAX_ROWS - 1 || d < 1 || d > F16_MAX_D) {
        f16_set_error("f16_forest_fit: bad arguments (n=%lld d=%d)", (long long)n, d);
        return F16_ERR_INVALID;
    }
    if (kind != F16_KIND_DT && kind != F16_KIND_RF && kind != F16_KIND_ET) { f16_set_error("f16_forest_fit: bad kind"); return F16_ERR_INVALID; }
    const int n_trees = (kind == F16_KIND_DT) ? 1 : n_estimators;
    if (n_trees < 1) { f16_set_error("f16_forest_fit: n_estimators < 1"); return F16_ERR_INVALID; }
    if (max_features < 1 || max_features > d) { f16_set_error("f16_forest_fit: bad max_features"); return F16_ERR_INVALID; }
    const bool best = (kind != F16_KIND_ET);
    if (best && !sorted_idx_dev) { f16_set_error("f16_forest_fit: sorted_idx required for DT/RF"); return F16_ERR_INVALID; }
    const int dp = (d <= 8) ? 8 : 16;

    std::vector<uint32_t> tree_seed(n_trees), rr(n_trees);
    int rc = f16_tree_seeds(seed, kind, n_trees, tree_seed.data(), rr.data());
    if (rc) return rc;

    f16_forest* F = (f16_forest*)calloc(1, sizeof(f16_forest));
    F->kind = kind; F->n_trees = n_trees; F->d = d; F->dp = dp; F->n_train = n;
    F->node_cap = (int)(2 * n - 1);
    if (F->node_cap < 1) F->node_cap = 1;

    F16FitParams P;
    memset(&P, 0, sizeof(P));
    P.X = X_dev; P.y = y_dev; P.sorted_idx = sorted_idx_dev;
    P.n = (int)n; P.d = d; P.dp = dp; P.n_trees = n_trees; P.max_features = max_features;
    P.stack_cap = F16_STACK_CAP; P.node_cap = F->node_cap;

    uint32_t* rr_dev = nullptr; uint8_t* bw = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&F->nodes, sizeof(F16Node) * (size_t)n_trees * F->node_cap, st));
    CUDA_TRY(cudaMallocAsync((void**)&F->node_count, sizeof(int32_t) * n_trees, st));
    CUDA_TRY(cudaMallocAsync((void**)&F->err, sizeof(int32_t), st));
    CUDA_TRY(cudaMemsetAsync(F->err, 0, sizeof(int32_t), st));
    CUDA_TRY(cudaMemsetAsync(F->node_count, 0, sizeof(int32_t) * n_trees, st));
    CUDA_TRY(cudaMallocAsync((void**)&rr_dev, sizeof(uint32_t) * n_trees, st));
    CUDA_TRY(cudaMemcpyAsync(rr_dev, rr.data(), sizeof(uint32_t) * n_trees, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMallocAsync((void**)&P.stack, sizeof(F16StackRec) * (size_t)n_trees * P.stack_cap, st));
    size_t buf_words = (size_t)n_trees * 2 * (best ? (size_t)d : 1) * (size_t)n;
    CUDA_TRY(cudaMallocAsync((void**)&P.buf, sizeof(uint32_t) * buf_words, st));
    P.rand_r_state = rr_dev; P.nodes = F->nodes; P.node_count = F->node_count; P.err = F->err;

    if (kind == F16_KIND_RF) {
        size_t stride = align_up((size_t)n, 4);
        CUDA_TRY(cudaMallocAsync((void**)&bw, stride * n_trees, st));
        rc = f16_bootstrap_counts(tree_seed.data(), n_trees, n, bw, stream);
        if (rc) return rc;
        P.boot_w = bw;
    }
    size_t dyn = 0;
    if (best) {
        P.side_words = (int)((n + 31) / 32);
        if (P.side_words > F16_SIDE_SMEM_MAX_WORDS) {
            CUDA_TRY(cudaMallocAsync((void**)&P.side_global, sizeof(uint32_t) * (size_t)n_trees * P.side_words, st));
        } else {
            dyn = sizeof(uint32_t) * (size_t)P.side_words;
        }
    }
    if (f16_get_profiling()) {
        CUDA_TRY(cudaEventCreate(&F->ev0));
        CUDA_TRY(cudaEventCreate(&F->ev1));
        F->has_ev = 1;
        CUDA_TRY(cudaEventRecord(F->ev0, st));
    }
    f16_count_launch(1);
    if (best) {
        if (dp == 8) {
            CUDA_TRY(cudaFuncSetAttribute(k_build_best<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * F16_SIDE_SMEM_MAX_WORDS));
            k_build_best<8><<<n_trees, NT, dyn, st>>>(P);
        } else {
            CUDA_TRY(cudaFuncSetAttribute(k_build_best<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * F16_SIDE_SMEM_MAX_WORDS));
            k_build_best<16><<<n_trees, NT, dyn, st>>>(P);
        }
    } else {
        if (dp == 8) k_build_random<8><<<n_trees, NT, 0, st>>>(P);
        else k_build_random<16><<<n_trees, NT, 0, st>>>(P);
    }
    CUDA_TRY(cudaGetLastError());
    if (F->has_ev) CUDA_TRY(cudaEventRecord(F->ev1, st));
    CUDA_TRY(cudaFreeAsync(P.buf, st));
    CUDA_TRY(cudaFreeAsync(P.stack, st));
    CUDA_TRY(cudaFreeAsync(rr_dev, st));
    if (bw) CUDA_TRY(cudaFreeAsync(bw, st));
    if (P.side_global) CUDA_TRY(cudaFreeAsync(P.side_global, st));
    *out = F;
    return F16_OK;
}

extern "C" int f16_forest_predict(const f16_forest* F, const float* X_dev, int64_t n, uint8_t* pred_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!F || !X_dev || !pred_dev || n < 0) { f16_set_error("f16_forest_predict: bad arguments"); return F16_ERR_INVALID; }
    if (n == 0) return F16_OK;
    int2* leaf = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&leaf, sizeof(int2) * (size_t)F->n_trees * n, st));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)F->n_trees);
    if (F->dp == 8) k_predict_walk<8><<<grid, 256, 0, st>>>(F->nodes, F->node_cap, X_dev, (int)n, leaf);
    else k_predict_walk<16><<<grid, 256, 0, st>>>(F->nodes, F->node_cap, X_dev, (int)n, leaf);
    CUDA_TRY(cudaGetLastError());
    f16_count_launch(2);
    k_predict_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(leaf, (int)n, F->n_trees, pred_dev);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(leaf, st));
    return F16_OK;
}

// Synchronises `stream` and returns the device-side status of the fit (0 ok).
extern "C" int f16_forest_status(const f16_forest* F, void* stream) {
    if (!F) { f16_set_error("f16_forest_status: null forest"); return F16_ERR_INVALID; }
    int32_t e = 0;
    CUDA_TRY(cudaMemcpyAsync(&e, F->err, sizeof(e), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (e) f16_set_error("forest fit failed on device: code %d (node/stack capacity or bootstrap weight > 127)", e);
    return e;
}

extern "C" int f16_forest_n_trees(const f16_forest* F) { return F ? F->n_trees : 0; }

// Synchronises on the build kernel's end event; milliseconds spent in the tree-building
// kernel alone (needs f16_set_profiling(1) before the fit), or -1.
extern "C" double f16_forest_build_ms(const f16_forest* F) {
    if (!F || !F->has_ev) return -1.0;
    float ms = -1.f;
    if (cudaEventSynchronize(F->ev1) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, F->ev0, F->ev1) != cudaSuccess) return -1.0;
    return (double)ms;
}

extern "C" int f16_forest_node_counts(const f16_forest* F, int32_t* counts_host, void* stream) {
    if (!F || !counts_host) { f16_set_error("f16_forest_node_counts: bad arguments"); return F16_ERR_INVALID; }
    CUDA_TRY(cudaMemcpyAsync(counts_host, F->node_count, sizeof(int32_t) * F->n_trees, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return F16_OK;
}

// Exports one tree in sklearn's `tree_` array layout (parity tests).  value: [n_nodes][2].
extern "C" int f16_forest_export(const f16_forest* F, int32_t tree, int64_t n_nodes, int64_t* left, int64_t* right,
                                 int64_t* feature, double* threshold, double* impurity, int64_t* n_node_samples,
                                 double* weighted_n_node_samples, double* value, void* stream) {
    if (!F || tree < 0 || tree >= F->n_trees) { f16_set_error("f16_forest_export: bad tree index"); return F16_ERR_INVALID; }
    std::vector<F16Node> h((size_t)n_nodes);
    CUDA_TRY(cudaMemcpyAsync(h.data(), F->nodes + (size_t)tree * F->node_cap, sizeof(F16Node) * (size_t)n_nodes,
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    for (int64_t i = 0; i < n_nodes; i++) {
        const F16Node& nd = h[i];
        bool leaf = nd.feature < 0;
        left[i] = leaf ? -1 : i + 1;
        right[i] = leaf ? -1 : nd.right;
        feature[i] = nd.feature;
        threshold[i] = nd.thr;
        double w = (double)nd.c0 + (double)nd.c1;
        double sq = 0.0; sq += (double)nd.c0 * (double)nd.c0; sq += (double)nd.c1 * (double)nd.c1;
        impurity[i] = 1.0 - sq / (w * w);
        n_node_samples[i] = nd.n;
        weighted_n_node_samples[i] = w;
        value[2 * i] = (double)nd.c0 / w;
        value[2 * i + 1] = (double)nd.c1 / w;
    }
    return F16_OK;
}

extern "C" void f16_forest_free(f16_forest* F, void* stream) {
    if (!F) return;
    cudaStream_t st = (cudaStream_t)stream;
    if (F->nodes) cudaFreeAsync(F->nodes, st);
    if (F->node_count) cudaFreeAsync(F->node_count, st);
    if (F->err) cudaFreeAsync(F->err, st);
    if (F->has_ev) { cudaEventDestroy(F->ev0); cudaEventDestroy(F->ev1); }
    free(F);
}
