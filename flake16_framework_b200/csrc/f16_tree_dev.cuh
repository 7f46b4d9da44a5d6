// Device-side pieces shared by the two tree-building kernels (random / best splitter).
#pragma once
#include "f16_tree.cuh"
#include <math.h>

// Threads per tree-building CTA and the kernel-name suffix are set per compiled variant
// (-DNT=.. -DF16_VARIANT=..): ExtraTrees / RandomForest trees are plentiful, so they use small
// CTAs (many trees resident per SM); the single DecisionTree uses a wide one.
#ifndef NT
#define NT 256
#endif
#ifndef F16_VARIANT
#define F16_VARIANT _v
#endif
#ifndef F16_MINB
#define F16_MINB 3
#endif
#define F16_CAT_(a, b) a##b
#define F16_CAT(a, b) F16_CAT_(a, b)
#define NW (NT / 32)
#define F16_EPS 2.220446049250313e-16
#define SSTK 40     // stack records cached in shared memory (deeper ones spill to global)

// The scalar sections of a tree (pop, feature draw, node record) and the single-warp subtree regime
// run on ONE warp of the CTA.  A warp's scheduler (SM sub-partition) is its index within the CTA
// modulo 4, so if that warp were warp 0 in every CTA, the serial sections of all CTAs resident on
// an SM would queue on sub-partition 0 while the other three idle.  The leader warp is therefore
// drawn per CTA and per launch; `rtid` is the thread index rotated so that the leader warp's
// lane 0 is 0 - role tests use rtid, data-parallel loops keep tid.
#ifndef F16_ROT
#define F16_ROT 1
#endif
__device__ __forceinline__ int f16_leader_warp(int salt) {
#if F16_ROT
    return (int)(((((uint32_t)blockIdx.x + (uint32_t)salt) * 2654435761u) >> 16) % (uint32_t)NW);
#else
    return 0;
#endif
}

// ------------------------------------------------------------------ shared control block
struct Ctl {
    int start, end, parent, c0, c1, n_const, is_left, depth;
    uint32_t const_mask;
    int in_smem;
    // done: written by pop_node only; abort: written by the node writers only; split: set by the
    // winner, cleared by the split path itself - so no flag is rewritten while a slower warp
    // may still be reading it after the last barrier of an iteration
    int done, abort, leaf, split;
    int ncand;
    int best_f;
    double best_thr;
    int n_left, l0, l1;
    int n_const_out;
    uint32_t const_mask_out;
    int node_id;
    int node_id_k;          // index of the winning candidate within the first chunk (ET)
    unsigned long long win_key;
    int sp, node_count;
};

// Fisher-Yates feature-draw state that persists across the nodes of a tree
// (sklearn Splitter.features / constant_features / rand_r_state).
struct DrawState {
    int features[F16_MAX_D];
    int const_feats[F16_MAX_D];
    uint32_t rng;
};

struct TreeStack {
    F16StackRec* smem;   // [SSTK]
    F16StackRec* gmem;   // [stack_cap]
    __device__ __forceinline__ F16StackRec get(int i) const { return i < SSTK ? smem[i] : gmem[i]; }
    __device__ __forceinline__ void put(int i, const F16StackRec& r) const { if (i < SSTK) smem[i] = r; else gmem[i] = r; }
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

// The improvement test cannot fail for a node that holds at most 1/16 of the training weight W
// (= n: unit weights, or bootstrap counts that sum to n): the computed (imp - B - C) is its
// non-negative true value (Gini is concave) minus at most eight roundings of quantities <= 1,
// i.e. > -9e-16; scaled by w_node / W <= 1/16 that is > -6e-17, four times smaller than the
// EPSILON it is compared with.  Skips four float64 divisions on the per-node critical path.
__device__ __forceinline__ bool improvement_certain(int c0, int c1, int n) {
    return (long long)(c0 + c1) * 16 <= (long long)n;
}

// the builder's leaf pre-test (_tree.pyx:223-240): n_node_samples < 2 or impurity <= EPSILON
// `impurity <= EPSILON` is decided without the float64 division: the class sums are integers
// below 2^24 (row-id limit), so for a pure node sq == w*w exactly and the Gini evaluates to
// exactly 0, while an impure node has Gini = 2 c0 c1 / w^2 >= 2 (w - 1) / w^2 > 1e-7, twelve
// orders of magnitude above EPSILON and its rounding error.
__device__ __forceinline__ bool leaf_pretest(int n_node, int c0, int c1) {
    return (n_node < 2) || (c0 == 0) || (c1 == 0);
}

// one thread: write node, link to parent, push children (right first: left is popped first)
__device__ __forceinline__ void finish_node(Ctl& c, const F16FitParams& P, F16Node* nodes, const TreeStack& stk) {
    int id = c.node_count++;
    if (id >= P.node_cap) { atomicExch(P.err, F16_ERR_OVERFLOW); c.abort = 1; return; }
    F16Node nd;
    nd.thr = c.split ? c.best_thr : -2.0;
    nd.feature = c.split ? c.best_f : -2;
    nd.right = -1;
    nd.c0 = c.c0; nd.c1 = c.c1; nd.n = c.end - c.start; nd.depth = c.depth;
    nodes[id] = nd;
    if (c.parent >= 0 && !c.is_left) nodes[c.parent].right = id;
    c.node_id = id;
    if (c.split) {
        if (c.sp + 2 > P.stack_cap) { atomicExch(P.err, F16_ERR_OVERFLOW); c.abort = 1; return; }
        F16StackRec r;
        r.parent = id; r.depth = c.depth + 1; r.n_const = (int16_t)c.n_const_out;
        r.const_mask = c.const_mask_out; r.pad = (uint8_t)c.in_smem;
        r.start = c.start + c.n_left; r.end = c.end; r.c0 = c.c0 - c.l0; r.c1 = c.c1 - c.l1; r.is_left = 0;
        stk.put(c.sp++, r);
        r.start = c.start; r.end = c.start + c.n_left; r.c0 = c.l0; r.c1 = c.l1; r.is_left = 1;
        stk.put(c.sp++, r);
    }
}

// one thread: pop + leaf pre-test
__device__ __forceinline__ void pop_node(Ctl& c, const TreeStack& stk) {
    if (c.sp == 0) { c.done = 1; return; }
    F16StackRec r = stk.get(--c.sp);
    c.start = r.start; c.end = r.end; c.parent = r.parent; c.c0 = r.c0; c.c1 = r.c1;
    c.n_const = r.n_const; c.const_mask = r.const_mask; c.is_left = r.is_left; c.depth = r.depth;
    c.in_smem = r.pad;
    c.leaf = leaf_pretest(r.end - r.start, r.c0, r.c1);
    c.ncand = 0;
    c.n_const_out = r.n_const; c.const_mask_out = r.const_mask;
}

// ------------------------------------------------------------------ stable block partition (one array)
// PU sub-tiles of NT elements per barrier: all loads of a round are issued before use.
#define PU 4
template <class Pred>
__device__ __forceinline__ void block_partition(const uint32_t* src, uint32_t* dst, int start, int n, int n_left,
                                                Pred pred, int (*s_wcnt)[PU][NW]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int run_l = 0, buf = 0;
    // the entries of the NEXT round are loaded while this round's predicate gathers are in flight
    uint32_t en[PU];
#pragma unroll
    for (int j = 0; j < PU; j++) { int p = j * NT + tid; en[j] = (p < n) ? src[start + p] : 0u; }
    for (int base = 0; base < n; base += NT * PU, buf ^= 1) {
        uint32_t e[PU]; bool valid[PU], left[PU]; unsigned bal[PU];
        const int nj = min(PU, (n - base + NT - 1) / NT);     // sub-tiles that exist (uniform)
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) { valid[j] = base + j * NT + tid < n; e[j] = en[j]; }
        }
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) left[j] = valid[j] && pred(e[j], start + base + j * NT + tid);
        }
        if (base + NT * PU < n) {
#pragma unroll
            for (int j = 0; j < PU; j++) { int p = base + NT * PU + j * NT + tid; en[j] = (p < n) ? src[start + p] : 0u; }
        }
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) {
                bal[j] = __ballot_sync(F16_FULL, left[j]);
                if (lane == 0) s_wcnt[buf][j][warp] = __popc(bal[j]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < PU; j++) {
            if (j < nj) {
                int before = 0, tot = 0;
#pragma unroll
                for (int q = 0; q < NW; q++) { int cq = s_wcnt[buf][j][q]; if (q < warp) before += cq; tot += cq; }
                int lrank = before + __popc(bal[j] & ((1u << lane) - 1u));
                if (valid[j]) {
                    if (left[j]) dst[start + run_l + lrank] = e[j];
                    else dst[start + n_left + (base + j * NT - run_l) + (tid - lrank)] = e[j];
                }
                run_l += tot;
            }
        }
    }
}

// Optional per-phase cycle accounting of the tree builders (tools/phase_probe.py builds a
// variant with -DF16_PHASE_TIMING; the product library is compiled without it).
#ifdef F16_PHASE_TIMING
#define F16_NPH 12
static __device__ unsigned long long f16_phase_cycles[2][F16_NPH];   // one copy per translation unit
static __device__ unsigned long long f16_phase_nodes[2][F16_NPH];
#define F16_PHASE_READER(name) extern "C" int name(unsigned long long* cycles, unsigned long long* nodes, int reset) { \
    if (cudaMemcpyFromSymbol(cycles, f16_phase_cycles, sizeof(f16_phase_cycles)) != cudaSuccess) return 1; \
    if (cudaMemcpyFromSymbol(nodes, f16_phase_nodes, sizeof(f16_phase_nodes)) != cudaSuccess) return 1; \
    if (reset) { unsigned long long z[2][F16_NPH] = {}; cudaMemcpyToSymbol(f16_phase_cycles, z, sizeof(z)); \
                 cudaMemcpyToSymbol(f16_phase_nodes, z, sizeof(z)); } \
    return 0; }
#define PH_DECL long long ph_last = clock64();
#define PH_T(kind, k) do { if (threadIdx.x == 0) { long long ph_now = clock64(); \
    atomicAdd(&f16_phase_cycles[kind][k], (unsigned long long)(ph_now - ph_last)); \
    atomicAdd(&f16_phase_nodes[kind][k], 1ull); ph_last = ph_now; } } while (0)
#else
#define PH_DECL
#define PH_T(kind, k) do { } while (0)
#define F16_PHASE_READER(name)
#endif

// launchers implemented in f16_tree_random.cu / f16_tree_best.cu
int f16_launch_build_random_et(const F16FitParams& P, cudaStream_t st);
int f16_launch_build_best_rf(const F16FitParams& P, size_t dyn_smem, cudaStream_t st);
int f16_launch_build_best_dt(const F16FitParams& P, size_t dyn_smem, cudaStream_t st);
int f16_launch_bootstrap(const uint32_t* seeds_dev, int n_trees, int n, uint32_t* w32, int words_per_tree, cudaStream_t st);
