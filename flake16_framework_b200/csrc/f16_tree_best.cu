// RandomForest / DecisionTree: DepthFirstTreeBuilder + BestSplitter + Gini, one CTA per tree,
// plus the RandomForest bootstrap.
//   node_split_best        sklearn/tree/_splitter.pyx:262-504
//   sort_samples_and_feature_values / next_p / partition_samples_final
//                          sklearn/tree/_partitioner.pyx:59-109, :169-215, :248-279
//   builder loop           sklearn/tree/_tree.pyx:139-336
//   bootstrap              sklearn/ensemble/_forest.py:94-112, :150-156
// (reference call sites: RandomForestClassifier / DecisionTreeClassifier(random_state=0).fit,
//  experiment.py:97-98,469)
//
// sklearn sorts the node's rows by every drawn feature at every node.  Here each tree keeps one
// index array PER FEATURE, sorted by that feature's value for the whole training set
// (f16_argsort_columns, once per training set) and STABLY PARTITIONED at every split, so a
// node's rows are a contiguous, still sorted slice of every array: the candidate scan is one
// linear pass (warp prefix-scan of the packed class weights, Gini proxy at every boundary
// between distinct values), min/max are the two ends of the slice, and there is no per-node
// sort.  A warp owns a feature, so a DecisionTree node scans/partitions 8 features at a time.
// Entries pack (row id | bootstrap weight << 24 | label << 31): the scan needs one gather
// (the feature value) per entry.  The side (left/right) of every row of the node is marked in
// a per-tree bit mask held in shared memory (n <= 524288) and looked up by the 16 partitions.
// Compile with -fmad=false.
#include "f16_tree_dev.cuh"

#ifdef F16_WITH_BOOTSTRAP
#define BNT 256
#define BNW 8
// ------------------------------------------------------------------ bootstrap (RF)
// One CTA per tree: MT19937(tree_seed).randint(0, n, n) -> bincount, entirely on device.
// The 624-word state block is regenerated in three dependency-free phases.
__global__ void __launch_bounds__(BNT) k_bootstrap(const uint32_t* __restrict__ tree_seed, int n,
                                                 uint32_t* __restrict__ w32 /*[n_trees][ceil(n/4)]*/,
                                                 int words_per_tree) {
    __shared__ uint32_t mt[624];
    __shared__ uint32_t out[624];
    __shared__ int s_wsum[2][BNW];
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
        for (int i = tid; i < 624; i += BNT) out[i] = f16_mt_temper(mt[i]) & mask;
        __syncthreads();
        // ---- ordered acceptance: only the first (n - produced) accepted draws count
        for (int base = 0; base < 624 && produced < n; base += BNT, buf ^= 1) {
            int i = base + tid;
            uint32_t v = (i < 624) ? out[i] : 0xffffffffu;
            bool acc = (i < 624) && (v <= rng);
            unsigned bal = __ballot_sync(F16_FULL, acc);
            if (lane == 0) s_wsum[buf][warp] = __popc(bal);
            __syncthreads();
            int before = 0, tot = 0;
            for (int q = 0; q < BNW; q++) { int cq = s_wsum[buf][q]; if (q < warp) before += cq; tot += cq; }
            int rank = produced + before + __popc(bal & ((1u << lane) - 1u));
            if (acc && rank < n) atomicAdd(&w[v >> 2], 1u << ((v & 3u) * 8u));
            produced += tot;
        }
        __syncthreads();
    }
}

int f16_launch_bootstrap(const uint32_t* seeds_dev, int n_trees, int n, uint32_t* w32, int words_per_tree, cudaStream_t st) {
    k_bootstrap<<<n_trees, BNT, 0, st>>>(seeds_dev, n, w32, words_per_tree);
    f16_count_launch(1);
    return cudaGetLastError() == cudaSuccess ? F16_OK : F16_ERR_CUDA;
}

#endif

// ------------------------------------------------------------------ best splitter
#ifndef F16_WPU
#define F16_WPU 8             // 32-entry tiles per round of the list partition (this many loads in flight + as many prefetched)
#endif
#ifndef F16_MAX_SEG
#define F16_MAX_SEG 8            // most warps that share one evaluated feature's scan
#endif
#ifndef F16_SPLIT_SCAN
#define F16_SPLIT_SCAN 1     // two warps per evaluated feature when the CTA has them (see the candidate scan)
#endif
#define k_build_best F16_CAT(k_build_best, F16_VARIANT)
struct BestCand {
    double proxy;
    unsigned long long key;   // (visit order k << 32) | position p ; smaller wins ties
    float v_prev, v;
    int l0, l1;
};

__device__ __forceinline__ bool side_get(const uint32_t* side, uint32_t id) { return (side[id >> 5] >> (id & 31)) & 1u; }

// warp-cooperative stable partition of one sorted feature array by the side bits;
// all loads of a round (8 x 32 entries) are issued before the first use
__device__ __forceinline__ void warp_partition(const uint32_t* src, uint32_t* dst, int start, int n, int n_left,
                                               const uint32_t* side) {
    const int lane = threadIdx.x & 31;
    int run_l = 0;
    constexpr int WPU = F16_WPU;
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
    // full rounds of WPU x 32 entries; the entries of the NEXT round are loaded before this round's
    // ranks are computed and stored, so a round never waits for a cold load of its own
    uint32_t en[WPU];
    if (32 * WPU <= n) {
#pragma unroll
        for (int j = 0; j < WPU; j++) en[j] = src[start + j * 32 + lane];
    }
    for (; base + 32 * WPU <= n; base += 32 * WPU) {
        uint32_t e[WPU]; bool left[WPU];
#pragma unroll
        for (int j = 0; j < WPU; j++) e[j] = en[j];
        if (base + 64 * WPU <= n) {
#pragma unroll
            for (int j = 0; j < WPU; j++) en[j] = src[start + base + 32 * WPU + j * 32 + lane];
        }
#pragma unroll
        for (int j = 0; j < WPU; j++) left[j] = side_get(side, f16_id(e[j]));
#pragma unroll
        for (int j = 0; j < WPU; j++) {
            unsigned bal = __ballot_sync(F16_FULL, left[j]);
            int lrank = __popc(bal & lt);
            int pos = left[j] ? (run_l + lrank) : (n_left + (base + j * 32 - run_l) + (lane - lrank));
            dst[start + pos] = e[j];
            run_l += __popc(bal);
        }
    }
    // tail: up to 4 tiles per round, only the tiles that exist are executed (small nodes pay
    // for what they have)
    for (; base < n; base += 128) {
        const int nj = min(4, (n - base + 31) >> 5);
        uint32_t e[4]; bool valid[4], left[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j < nj) { int p = base + j * 32 + lane; valid[j] = p < n; e[j] = valid[j] ? src[start + p] : 0u; }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (j < nj) left[j] = valid[j] && side_get(side, f16_id(e[j]));
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
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

// SHARED regime: a node with at most SB rows is relocated once into shared memory - its d sorted
// lists (entries re-indexed to local row ids), its feature values (column-major, odd stride) - and
// its whole subtree is grown by the same block-wide code on those arrays: no global loads, only
// node records go to HBM.  The region aliases the side bits of the GLOBAL regime (never live at
// the same time: every node marks all of its rows before its partitions read them).
template <int DP> struct BestSub { static constexpr int SB = (DP == 16) ? 256 : 512; };
#define F16_BEST_SUB_BYTES(DP) (2 * (DP) * BestSub<DP>::SB * 4 + (DP) * (BestSub<DP>::SB + 1) * 4 + BestSub<DP>::SB / 8)

template <int DP>
__global__ void __launch_bounds__(NT, F16_MINB) k_build_best(F16FitParams P) {
    constexpr int SB = BestSub<DP>::SB;
    constexpr int SBP = SB + 1;
    extern __shared__ uint32_t s_side_dyn[];
    uint32_t* s_ord = s_side_dyn;                                           // [2][DP][SB]
    float* s_val = reinterpret_cast<float*>(s_side_dyn + 2 * DP * SB);      // [DP][SBP]
    uint32_t* s_side_l = s_side_dyn + 2 * DP * SB + DP * SBP;               // [SB / 32]
    __shared__ Ctl c;
    __shared__ DrawState ds;
    __shared__ F16StackRec s_stack[SSTK];
    __shared__ float s_min[F16_MAX_D], s_max[F16_MAX_D];
    __shared__ int s_eval_f[F16_MAX_D];
    __shared__ double s_bproxy[NW];
    __shared__ unsigned long long s_bkey[NW];
    __shared__ unsigned long long s_part[NW];
    __shared__ int s_cntn[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lw = f16_leader_warp(P.salt);                 // the CTA's scalar-section warp
    const int rtid = (tid - lw * 32) & (NT - 1);            // rotated: leader warp's lane 0 is 0
    const int t = blockIdx.x;
    const int n = P.n, d = P.d;
    const float* __restrict__ X = P.X;
    uint32_t* ord0 = P.buf + (size_t)t * 2 * d * n;   // [d][n]
    uint32_t* ord1 = ord0 + (size_t)d * n;
    F16Node* nodes = P.nodes + (size_t)t * P.node_cap;
    TreeStack stk;
    stk.smem = s_stack;
    stk.gmem = P.stack + (size_t)t * P.stack_cap;
    uint32_t* side = P.side_global ? P.side_global + (size_t)t * P.side_words : s_side_dyn;
    const uint8_t* bw = P.boot_w ? P.boot_w + (size_t)t * (((size_t)n + 3) / 4 * 4) : nullptr;
    const double W_total = (double)n;   // sum of bootstrap counts == n; unit weights == n
    PH_DECL

    // ---- root: per feature, compact the column argsort to the in-bag rows (weight > 0),
    //      packing (id, weight, label).  A warp owns a feature.
    //      First pack (id, weight, label) per row with coalesced reads (pk aliases the not yet
    //      used second ping-pong buffer), so the per-feature pass needs one gather per entry.
    uint32_t* pk = ord1;
    for (int i = tid; i < n; i += NT) {
        uint32_t w = bw ? (uint32_t)bw[i] : 1u;
        if (w > F16_MAX_W) { atomicExch(P.err, F16_ERR_OVERFLOW); w = F16_MAX_W; }
        pk[i] = w ? f16_pack((uint32_t)i, w, (uint32_t)P.y[i]) : 0u;
    }
    __syncthreads();
    for (int f = warp; f < d; f += NW) {
        const int32_t* sidx = P.sorted_idx + (size_t)f * n;
        uint32_t* o = ord0 + (size_t)f * n;
        int run = 0;
        unsigned long long cls = 0;
        for (int base = 0; base < n; base += 256) {
            uint32_t id[8], e[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { int p = base + j * 32 + lane; id[j] = (p < n) ? (uint32_t)__ldg(sidx + p) : 0xffffffffu; }
#pragma unroll
            for (int j = 0; j < 8; j++) e[j] = (id[j] != 0xffffffffu) ? pk[id[j]] : 0u;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                bool keep = e[j] != 0u;
                unsigned bal = __ballot_sync(F16_FULL, keep);
                if (keep) {
                    o[run + __popc(bal & ((1u << lane) - 1u))] = e[j];
                    cls += (unsigned long long)f16_w(e[j]) << (f16_y(e[j]) ? 32 : 0);
                }
                run += __popc(bal);
            }
        }
        if (f == 0) {
            cls = f16_warp_sum_u64(cls);
            if (lane == 0) { s_part[0] = cls; s_cntn[0] = run; }
        }
    }
    __syncthreads();
    if (rtid == 0) {
        unsigned long long tot = s_part[0];
        for (int f = 0; f < F16_MAX_D; f++) { ds.features[f] = f; ds.const_feats[f] = 0; }
        ds.rng = P.rand_r_state[t];
        F16StackRec r;
        r.start = 0; r.end = s_cntn[0]; r.parent = -1; r.c0 = (int)(uint32_t)tot; r.c1 = (int)(tot >> 32);
        r.const_mask = 0; r.n_const = 0; r.is_left = 0; r.pad = 0; r.depth = 0;
        c.sp = 0; c.node_count = 0; c.done = 0; c.abort = 0; c.split = 0;
        stk.put(c.sp++, r);
    }
    __syncthreads();

    PH_T(1, 0);
    while (true) {
        if (rtid == 0) pop_node(c, stk);
        __syncthreads();
        if (c.done) break;
        const int nn = c.end - c.start;
        if (!c.leaf && !c.in_smem && nn <= SB && (~c.const_mask & ((1u << d) - 1u))) {
            // ================= relocate the node into shared memory
            const uint32_t* gsrc = (c.depth & 1) ? ord1 : ord0;
            uint32_t* lid = P.lid + (size_t)t * n;
            const int f0 = __ffs(~c.const_mask & ((1u << d) - 1u)) - 1;   // a list that is maintained
            const int gstart = c.start;
            {
                constexpr int Q = DP / 4;
                const int q = tid % Q;
                const uint32_t* o = gsrc + (size_t)f0 * n + gstart;
                for (int p = tid / Q; p < nn; p += NT / Q) {
                    uint32_t gid = f16_id(o[p]);
                    float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)gid * DP) + q);
                    s_val[(q * 4 + 0) * SBP + p] = v.x;
                    s_val[(q * 4 + 1) * SBP + p] = v.y;
                    s_val[(q * 4 + 2) * SBP + p] = v.z;
                    s_val[(q * 4 + 3) * SBP + p] = v.w;
                    if (q == 0) lid[gid] = (uint32_t)p;
                }
            }
            __syncthreads();
            uint32_t* sdst = s_ord + ((c.depth & 1) ? DP * SB : 0);
            for (int f = warp; f < d; f += NW) {
                if ((c.const_mask >> f) & 1u) continue;
                const uint32_t* o = gsrc + (size_t)f * n + gstart;
                for (int p = lane; p < nn; p += 32) {
                    uint32_t e = o[p];
                    sdst[f * SB + p] = (e & ~F16_ID_MASK) | lid[f16_id(e)];
                }
            }
            if (rtid == 0) { c.in_smem = 1; c.start = 0; c.end = nn; }
            __syncthreads();
            PH_T(1, 1);
        }
        const bool sm = c.in_smem != 0;
        const int start = c.start;
        const int stride = sm ? SB : n;
        const uint32_t* src = sm ? (s_ord + ((c.depth & 1) ? DP * SB : 0)) : ((c.depth & 1) ? ord1 : ord0);
        uint32_t* dst = sm ? (s_ord + ((c.depth & 1) ? 0 : DP * SB)) : ((c.depth & 1) ? ord0 : ord1);
        uint32_t* nside = sm ? s_side_l : side;
        auto value = [&](uint32_t e, int f) -> float {
            return sm ? s_val[f * SBP + f16_id(e)] : __ldg(X + (size_t)f16_id(e) * DP + f);
        };
        int n_eval = 0;

        if (!c.leaf) {
            // ---- min / max of every not-yet-constant feature: ends of its sorted slice
            if (rtid < d && !((c.const_mask >> rtid) & 1u)) {
                const uint32_t* o = src + (size_t)rtid * stride;
                s_min[rtid] = value(o[start], rtid);
                s_max[rtid] = value(o[start + nn - 1], rtid);
            }
            __syncthreads();
            // ---- Fisher-Yates feature draw (thread 0)
            if (rtid == 0) {
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
            PH_T(1, sm ? 5 : 2);
            n_eval = c.ncand;
            // ---- candidate scan.  A warp owns an evaluated feature; when the CTA has at least two
            //      warps per evaluated feature (RandomForest: max_features = 4, 8 warps) the slice is
            //      split in two: the even warp walks the first half upwards with LEFT prefix sums, the
            //      odd warp walks the second half downwards with RIGHT suffix sums (left = node total
            //      - right: the class sums are integers, so this is exact) - every warp of the CTA
            //      works and no pre-pass is needed.  Ties are broken by the key (visit order, position).
            BestCand best;
            best.proxy = -INFINITY; best.key = ~0ull; best.v_prev = 0.f; best.v = 0.f; best.l0 = 0; best.l1 = 0;
            const int t0 = c.c0, t1 = c.c1;
            // SEG warps per evaluated feature (a power of two, NW / n_eval rounded down, at most 8): half of
            // them walk segments of the first half of the slice upwards, half walk segments of the second
            // half downwards.  A segment that does not start at an end of the slice first sums the class
            // weights that precede it (coalesced entry loads, no value gathers) - its prefix offset.
            int seg = 1;
            if (F16_SPLIT_SCAN && n_eval > 0) { while (seg < F16_MAX_SEG && 2 * seg * n_eval <= NW) seg <<= 1; }
            const bool split_scan = seg > 1;
            const int segh = seg >> 1;                           // segments per direction
            const int n_scan = n_eval * seg;
            for (int wk = warp; wk < n_scan; wk += NW) {
                const int k = split_scan ? (wk / seg) : wk;
                const int sidx = split_scan ? (wk % seg) : 0;
                const bool rev = split_scan && (sidx >= segh);
                const int h = split_scan ? (nn + 1) / 2 : nn;   // upwards: positions [1, h); downwards: [h, nn)
                const int f = s_eval_f[k];
                const uint32_t* o = src + (size_t)f * stride + start;
                unsigned long long carry = 0;
                if (!rev) {
                    // this warp's segment of the upward half: entries [lo, hi)
                    const int lo = split_scan ? (int)((long long)sidx * h / segh) : 0;
                    const int hi = split_scan ? (int)((long long)(sidx + 1) * h / segh) : nn;
                    float prev_last = 0.f;
                    if (lo > 0) {
                        unsigned long long pre = 0;
                        for (int p = lane; p < lo; p += 32) { const uint32_t e = o[p]; pre += (unsigned long long)f16_w(e) << (f16_y(e) ? 32 : 0); }
                        carry = f16_warp_sum_u64(pre);
                        prev_last = value(o[lo - 1], f);
                    }
                    // 4 consecutive entries per lane: one warp scan per 128 entries; the entries of the
                    // NEXT round are loaded while this round's value gathers are in flight (entry ->
                    // value is a chain of two memory latencies; overlapped, a round costs one)
                    uint32_t en[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) en[j] = (lo + lane * 4 + j < hi) ? o[lo + lane * 4 + j] : 0u;
                    for (int base = lo; base < hi; base += 128) {
                        const int p0 = base + lane * 4;
                        uint32_t e[4]; float v[4]; unsigned long long my[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) e[j] = en[j];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            v[j] = (p0 + j < hi) ? value(e[j], f) : INFINITY;
                        if (base + 128 < hi) {
#pragma unroll
                            for (int j = 0; j < 4; j++) en[j] = (p0 + 128 + j < hi) ? o[p0 + 128 + j] : 0u;
                        }
                        unsigned long long run = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            my[j] = (p0 + j < hi) ? ((unsigned long long)f16_w(e[j]) << (f16_y(e[j]) ? 32 : 0)) : 0ull;
                            run += my[j];
                        }
                        unsigned long long incl = f16_warp_incl_scan_u64(run);
                        unsigned long long ex = carry + incl - run;
                        float vp = __shfl_up_sync(F16_FULL, v[3], 1);
                        if (lane == 0) vp = prev_last;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int p = p0 + j;
                            if (p < hi && p > 0 && v[j] > __fadd_rn(vp, 1e-7f)) {
                                int l0 = (int)(uint32_t)ex, l1 = (int)(ex >> 32);
                                double proxy = gini_proxy(l0, l1, t0, t1);
                                if (proxy > best.proxy) {
                                    best.proxy = proxy; best.key = ((unsigned long long)k << 32) | (unsigned)p;
                                    best.v_prev = vp; best.v = v[j]; best.l0 = l0; best.l1 = l1;
                                }
                            }
                            ex += my[j]; vp = v[j];
                        }
                        carry += __shfl_sync(F16_FULL, incl, 31);
                        prev_last = __shfl_sync(F16_FULL, v[3], 31);
                    }
                } else {
                    // reversed element i stands for entry q = nn - 1 - i; its inclusive prefix is the
                    // RIGHT sum of candidate position q (rows q .. nn-1); the candidate test compares
                    // entry q with entry q - 1 = reversed element i + 1 (the lane's next element, the
                    // next lane's first one, or - for lane 31 - one extra load)
                    const int m = nn - h;                  // candidate positions h .. nn-1 = reversed elements 0 .. m-1
                    // this warp's segment of the downward half: reversed elements [i_lo, i_hi)
                    const int rs = sidx - segh;
                    const int i_lo = (int)((long long)rs * m / segh), i_hi = (int)((long long)(rs + 1) * m / segh);
                    if (i_lo > 0) {
                        unsigned long long pre = 0;
                        for (int i = lane; i < i_lo; i += 32) { const uint32_t e = o[nn - 1 - i]; pre += (unsigned long long)f16_w(e) << (f16_y(e) ? 32 : 0); }
                        carry = f16_warp_sum_u64(pre);
                    }
                    // element m (entry h - 1) is loaded for its value only
                    uint32_t en[4], en_after;
#pragma unroll
                    for (int j = 0; j < 4; j++) en[j] = (i_lo + lane * 4 + j <= m) ? o[nn - 1 - (i_lo + lane * 4 + j)] : 0u;
                    en_after = (lane == 31 && i_lo + lane * 4 + 4 <= m) ? o[nn - 1 - (i_lo + lane * 4 + 4)] : 0u;
                    for (int base = i_lo; base < i_hi; base += 128) {
                        const int i0 = base + lane * 4;
                        uint32_t e[4]; float v[4]; unsigned long long my[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) e[j] = en[j];
                        const uint32_t e_after = en_after;
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            v[j] = (i0 + j <= m) ? value(e[j], f) : -INFINITY;
                        float v_after = (lane == 31 && i0 + 4 <= m) ? value(e_after, f) : -INFINITY;
                        if (base + 128 < i_hi) {
#pragma unroll
                            for (int j = 0; j < 4; j++) en[j] = (i0 + 128 + j <= m) ? o[nn - 1 - (i0 + 128 + j)] : 0u;
                            en_after = (lane == 31 && i0 + 132 <= m) ? o[nn - 1 - (i0 + 132)] : 0u;
                        }
                        unsigned long long run = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            my[j] = (i0 + j < m) ? ((unsigned long long)f16_w(e[j]) << (f16_y(e[j]) ? 32 : 0)) : 0ull;
                            run += my[j];
                        }
                        unsigned long long incl = f16_warp_incl_scan_u64(run);
                        unsigned long long r = carry + incl - run;
                        const float vdown = __shfl_down_sync(F16_FULL, v[0], 1);
                        if (lane != 31) v_after = vdown;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            r += my[j];
                            const float vn = (j < 3) ? v[j < 3 ? j + 1 : 3] : v_after;     // entry q - 1
                            if (i0 + j < i_hi && v[j] > __fadd_rn(vn, 1e-7f)) {
                                const int q = nn - 1 - (i0 + j);
                                int l0 = t0 - (int)(uint32_t)r, l1 = t1 - (int)(r >> 32);
                                double proxy = gini_proxy(l0, l1, t0, t1);
                                const unsigned long long key = ((unsigned long long)k << 32) | (unsigned)q;
                                if (proxy > best.proxy || (proxy == best.proxy && key < best.key)) {
                                    best.proxy = proxy; best.key = key;
                                    best.v_prev = vn; best.v = v[j]; best.l0 = l0; best.l1 = l1;
                                }
                            }
                        }
                        carry += __shfl_sync(F16_FULL, incl, 31);
                    }
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
                if (rtid == 0) {
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
                    c.split = (improvement_certain(t0, t1, n) || improvement_ok(best.l0, best.l1, t0, t1, W_total)) ? 1 : 0;
                }
                __syncthreads();
                PH_T(1, sm ? 6 : 3);
            }
        }
        if (rtid == 0) finish_node(c, P, nodes, stk);
        __syncthreads();
        if (c.abort) break;
        if (c.split) {
            // ---- mark the side of every row of the node (the winning feature's slice is
            //      sorted, so the left rows are its first n_left entries)
            const int n_left = c.n_left;
            const uint32_t* o = src + (size_t)c.best_f * stride + start;
            for (int p = tid; p < nn; p += NT) {
                uint32_t id = f16_id(o[p]);
                if (p < n_left) atomicOr(&nside[id >> 5], 1u << (id & 31));
                else atomicAnd(&nside[id >> 5], ~(1u << (id & 31)));
            }
            __syncthreads();
            // ---- stable partition of every still-useful feature array, a warp per array
            const uint32_t keep_mask = ~c.const_mask_out;
            for (int f = warp; f < d; f += NW) {
                if (!((keep_mask >> f) & 1u)) continue;
                warp_partition(src + (size_t)f * stride, dst + (size_t)f * stride, start, nn, n_left, nside);
            }
            __syncthreads();
            if (rtid == 0) c.split = 0;
            PH_T(1, sm ? 7 : 4);
        } else {
            PH_T(1, sm ? 9 : 8);
        }
    }
    if (rtid == 0) { const int nc = min(c.node_count, P.node_cap); P.node_count[t] = nc; atomicMax(P.err + 1, nc); }
}

F16_PHASE_READER(F16_CAT(f16_debug_phases, F16_VARIANT))

int F16_CAT(f16_launch_build_best, F16_VARIANT)(const F16FitParams& P, size_t dyn_smem, cudaStream_t st) {
    cudaError_t e;
    const size_t max_dyn = 4 * F16_SIDE_SMEM_MAX_WORDS;     // 64 KiB >= both regimes' needs
    if (P.dp == 8) {
        size_t need = F16_BEST_SUB_BYTES(8);
        if (dyn_smem < need) dyn_smem = need;
        e = cudaFuncSetAttribute(k_build_best<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn);
        if (e != cudaSuccess) return F16_ERR_CUDA;
        k_build_best<8><<<P.n_trees, NT, dyn_smem, st>>>(P);
    } else {
        size_t need = F16_BEST_SUB_BYTES(16);
        if (dyn_smem < need) dyn_smem = need;
        e = cudaFuncSetAttribute(k_build_best<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn);
        if (e != cudaSuccess) return F16_ERR_CUDA;
        k_build_best<16><<<P.n_trees, NT, dyn_smem, st>>>(P);
    }
    f16_count_launch(1);
    return cudaGetLastError() == cudaSuccess ? F16_OK : F16_ERR_CUDA;
}
