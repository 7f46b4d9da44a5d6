// ExtraTrees, WARP-PER-TREE variant: one 32-thread CTA grows one tree.
//
// Same algorithm and same results as f16_tree_random.cu (sklearn node_split_random,
// _splitter.pyx:507-736; builder _tree.pyx:139-336), different mapping to the machine.
// The nodes of a tree are sequential (the xorshift stream is consumed depth-first), and most
// nodes hold a few dozen rows, so a tree can only keep about one warp busy.  Giving every tree
// exactly one warp - no block barriers anywhere, up to 32 trees resident per SM, thousands per
// GPU when several forests are in flight - turns the latency-bound DFS of each tree into a
// throughput problem for the SM schedulers.
//   GLOBAL regime (n > S rows): the warp sweeps the node's slice of the per-tree sample list
//     three times (min/max of all features; left-counts of the <= 4 candidates; stable
//     partition), 4 x 32 loads in flight per sweep step, gathers from the L2-resident rows.
//   SHARED regime (n <= S): subtree_warp_v2 on a shared-memory copy of the node's rows.
// Compile with -fmad=false.
#include "f16_tree_dev.cuh"
#include "f16_tree_random_sub.cuh"

#ifndef F16_WS16
#define F16_WS16 64
#endif
#ifndef F16_WS8
#define F16_WS8 128
#endif
#define WSTK 24   // stack records cached in shared memory

template <int DP> struct WCfg;
template <> struct WCfg<16> { static constexpr int S = F16_WS16; };
template <> struct WCfg<8> { static constexpr int S = F16_WS8; };

struct WStack {
    F16StackRec* smem;
    F16StackRec* gmem;
    __device__ __forceinline__ F16StackRec get(int i) const { return i < WSTK ? smem[i] : gmem[i]; }
    __device__ __forceinline__ void put(int i, const F16StackRec& r) const { if (i < WSTK) smem[i] = r; else gmem[i] = r; }
};

template <int DP>
__global__ void __launch_bounds__(32, 32) k_build_random_w(F16FitParams P) {
    constexpr int Q = DP / 4;          // float4 quads per row
    constexpr int RPI = 32 / Q;        // rows per warp step in the min/max sweep
    constexpr int S = WCfg<DP>::S;
    constexpr int SP = S + 1;
    __shared__ Ctl c;
    __shared__ DrawState ds;
    __shared__ F16StackRec s_stack[SSTK];
    __shared__ float s_col[DP * SP];
    __shared__ uint16_t s_idx[2][S];
    __shared__ uint8_t s_y[S];

    const int lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const int t = blockIdx.x;
    const int n = P.n, d = P.d, max_features = P.max_features;
    const float* __restrict__ X = P.X;
    uint32_t* buf0 = P.buf + (size_t)t * 2 * n;
    uint32_t* buf1 = buf0 + n;
    F16Node* nodes = P.nodes + (size_t)t * P.node_cap;
    TreeStack stk;                     // same record layout / accessors as the CTA kernels
    stk.smem = s_stack;
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
            // ---- SHARED regime: the shared-memory subtree routine keeps its scalars in `c`/`ds`
            if (lane == 0) { c.sp = sp; c.node_count = node_count; ds.rng = rng; }
            __syncwarp();
            subtree_warp_v2<DP, S>(c, ds, stk, P, nodes, s_col, s_idx, s_y);
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
            // ---- relocate the node into shared memory, then fall into the SHARED regime
            const int q = lane % Q;
            for (int i = lane / Q; i < nn; i += RPI) {
                uint32_t e = src[start + i];
                float4 v = __ldg(reinterpret_cast<const float4*>(X + (size_t)f16_id(e) * DP) + q);
                s_col[(q * 4 + 0) * SP + i] = v.x;
                s_col[(q * 4 + 1) * SP + i] = v.y;
                s_col[(q * 4 + 2) * SP + i] = v.z;
                s_col[(q * 4 + 3) * SP + i] = v.w;
                if (q == 0) { s_idx[r.depth & 1][i] = (uint16_t)i; s_y[i] = (uint8_t)f16_y(e); }
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

        if (!leaf) {
            // ---- sweep 1: min / max of every feature (lane = (row in step, quad))
            float mn[4], mx[4];
#pragma unroll
            for (int j = 0; j < 4; j++) { mn[j] = INFINITY; mx[j] = -INFINITY; }
            {
                const int q = lane % Q, rl = lane / Q;
                for (int i0 = rl; i0 < nn; i0 += RPI * 4) {
                    uint32_t id[4]; float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { int i = i0 + u * RPI; id[u] = f16_id(src[start + (i < nn ? i : i0)]); }
#pragma unroll
                    for (int u = 0; u < 4; u++) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)id[u] * DP) + q);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
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

            auto eval_chunk = [&](int cnt) {
                // ---- sweep 2: left counts of up to 4 candidates, 2 x 32 rows in flight
                int nl[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0};
                for (int base = 0; base < nn; base += 64) {
                    const int i0 = base + lane, i1 = i0 + 32;
                    const bool v0 = i0 < nn, v1 = i1 < nn;
                    const uint32_t e0 = v0 ? src[start + i0] : 0u, e1 = v1 ? src[start + i1] : 0u;
                    const float* r0 = X + (size_t)f16_id(e0) * DP;
                    const float* r1 = X + (size_t)f16_id(e1) * DP;
                    float a[4], b[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) { a[k] = __ldg(r0 + cf[k]); b[k] = __ldg(r1 + cf[k]); }
                    const unsigned y0 = __ballot_sync(F16_FULL, v0 && f16_y(e0)), y1 = __ballot_sync(F16_FULL, v1 && f16_y(e1));
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (k < cnt) {
                            unsigned b0 = __ballot_sync(F16_FULL, v0 && ((double)a[k] <= ct[k]));
                            unsigned b1 = __ballot_sync(F16_FULL, v1 && ((double)b[k] <= ct[k]));
                            nl[k] += __popc(b0) + __popc(b1);
                            l1[k] += __popc(b0 & y0) + __popc(b1 & y1);
                        }
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

            while (f_i > n_total && (n_visited < max_features || n_visited <= n_found + n_drawn)) {
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
            if (ncand > 0) eval_chunk(ncand);
            __syncwarp();
            if (lane == 0) {
                for (int i = 0; i < n_known; i++) ds.features[i] = ds.const_feats[i];
                for (int i = n_known; i < n_total; i++) ds.const_feats[i] = ds.features[i];
            }
            __syncwarp();
            for (int i = n_known; i < n_total; i++) cmask |= 1u << ds.const_feats[i];
            if (best_f >= 0) split = improvement_ok(bl0, bl1, t0, t1, W_total);
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
            // ---- sweep 3: stable partition, 4 x 32 entries in flight
            int run_l = 0;
            for (int base = 0; base < nn; base += 128) {
                const int nj = min(4, (nn - base + 31) >> 5);
                uint32_t e[4]; bool valid[4], left[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j < nj) { int p = base + j * 32 + lane; valid[j] = p < nn; e[j] = valid[j] ? src[start + p] : 0u; }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    if (j < nj) left[j] = valid[j] && ((double)__ldg(X + (size_t)f16_id(e[j]) * DP + best_f) <= best_thr);
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

int f16_launch_build_random_w(const F16FitParams& P, cudaStream_t st) {
    if (P.dp == 8) k_build_random_w<8><<<P.n_trees, 32, 0, st>>>(P);
    else k_build_random_w<16><<<P.n_trees, 32, 0, st>>>(P);
    f16_count_launch(1);
    return cudaGetLastError() == cudaSuccess ? F16_OK : F16_ERR_CUDA;
}
