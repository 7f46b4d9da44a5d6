// Host side of the forest API (fit / predict / export) + the prediction kernels.
//   BaseForest.fit / _parallel_build_trees   sklearn/ensemble/_forest.py:303-523, :132-179
//   ForestClassifier.predict(_proba)         sklearn/ensemble/_forest.py:882-967, :704-717
//   Tree._apply_dense                        sklearn/tree/_tree.pyx:954-996
// (reference call sites: model.fit experiment.py:469, model.predict experiment.py:473)
// The tree-growing kernels live in f16_tree_random.cu (ExtraTrees) and f16_tree_best.cu
// (RandomForest / DecisionTree).
#include "f16_tree_dev.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" int f16_get_profiling(void);
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st);

#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

// ================================================================== predict
// walk: blockIdx.y = tree, threads over rows; leaf class sums -> leaf[tree][row]
template <int DP>
__global__ void __launch_bounds__(256) k_predict_walk(const F16Node* __restrict__ nodes, int node_cap,
                                                      const int32_t* __restrict__ node_count,
                                                      const float* __restrict__ X, int n, int2* __restrict__ leaf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F16Node* tn = nodes + (size_t)blockIdx.y * node_cap;
    const int nc = min(node_count[blockIdx.y], node_cap);     // an aborted fit counts the node it could not store
    const float* row = X + (size_t)i * DP;
    int id = 0;
    // a complete tree never leaves [0, nc) and is at most F16_STACK_CAP deep; a tree whose fit was
    // aborted (capacity overflow, reported by f16_forest_status) may hold dangling links - the
    // bounds keep the walk inside the tree's records and finite, its output is then meaningless
    for (int step = 0; step <= F16_STACK_CAP && id >= 0 && id < nc; step++) {
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
    leaf[(size_t)blockIdx.y * n + i] = make_int2(1, 0);
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
    CUDA_TRY(f16_malloc_async((void**)&seeds_dev, sizeof(uint32_t) * n_trees, st));
    CUDA_TRY(cudaMemcpyAsync(seeds_dev, tree_seed_host, sizeof(uint32_t) * n_trees, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemsetAsync(w_dev, 0, stride * n_trees, st));
    if (f16_launch_bootstrap(seeds_dev, n_trees, (int)n, (uint32_t*)w_dev, (int)(stride / 4), st)) {
        f16_set_error("k_bootstrap launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return F16_ERR_CUDA;
    }
    CUDA_TRY(cudaFreeAsync(seeds_dev, st));
    return F16_OK;
}

// Everything a fit allocates besides the forest itself; released (stream-ordered) on every path.
struct FitScratch {
    cudaStream_t st;
    void* p[8];
    int np;
    explicit FitScratch(cudaStream_t s) : st(s), np(0) {}
    cudaError_t alloc(void** out, size_t bytes) {
        cudaError_t e = f16_malloc_async(out, bytes, st);
        if (e == cudaSuccess) p[np++] = *out;
        return e;
    }
    ~FitScratch() { for (int i = 0; i < np; i++) cudaFreeAsync(p[i], st); }
};

static void forest_release(f16_forest* F, cudaStream_t st) {
    if (!F) return;
    if (F->nodes) cudaFreeAsync(F->nodes, st);
    if (F->node_count) cudaFreeAsync(F->node_count, st);
    if (F->err) cudaFreeAsync(F->err, st);
    if (F->has_ev) { cudaEventDestroy(F->ev0); cudaEventDestroy(F->ev1); }
    free(F);
}

// node_cap: capacity of one tree's node array.  A binary tree over n rows has at most 2n - 1
// nodes; callers that know better (f16_forest_fit_cap: a measured count of an earlier fit on a
// training set of the same kind) pass a smaller capacity, and a tree that outgrows it fails the
// fit with F16_ERR_OVERFLOW (f16_forest_status) instead of writing out of bounds.
static int forest_fit_impl(const float* X_dev, const uint8_t* y_dev, int64_t n, int32_t d,
                           const int32_t* sorted_idx_dev, int32_t kind, int32_t n_estimators,
                           int32_t max_features, uint32_t seed, int64_t node_cap, cudaStream_t st, f16_forest** out) {
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
    if (!F) { f16_set_error("f16_forest_fit: out of host memory"); return F16_ERR_NOMEM; }
    F->kind = kind; F->n_trees = n_trees; F->d = d; F->dp = dp; F->n_train = n;
    const int64_t full_cap = 2 * n - 1;
    F->node_cap = (int)((node_cap > 0 && node_cap < full_cap) ? node_cap : full_cap);
    if (F->node_cap < 1) F->node_cap = 1;

    F16FitParams P;
    memset(&P, 0, sizeof(P));
    P.X = X_dev; P.y = y_dev; P.sorted_idx = sorted_idx_dev;
    P.n = (int)n; P.d = d; P.dp = dp; P.n_trees = n_trees; P.max_features = max_features;
    P.stack_cap = F16_STACK_CAP; P.node_cap = F->node_cap;
    static unsigned launch_serial = 0;
    P.salt = (int)(__sync_fetch_and_add(&launch_serial, 1u) * 7u);

    FitScratch S(st);           // freed when this function returns, whatever the path
#define FIT_TRY(x)                                                                             \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            forest_release(F, st);                                                             \
            return e_ == cudaErrorMemoryAllocation ? F16_ERR_NOMEM : F16_ERR_CUDA;             \
        }                                                                                      \
    } while (0)
    uint32_t* rr_dev = nullptr; uint8_t* bw = nullptr;
    FIT_TRY(f16_malloc_async((void**)&F->nodes, sizeof(F16Node) * (size_t)n_trees * F->node_cap, st));
    FIT_TRY(f16_malloc_async((void**)&F->node_count, sizeof(int32_t) * n_trees, st));
    FIT_TRY(f16_malloc_async((void**)&F->err, 2 * sizeof(int32_t), st));
    FIT_TRY(cudaMemsetAsync(F->err, 0, 2 * sizeof(int32_t), st));
    FIT_TRY(cudaMemsetAsync(F->node_count, 0, sizeof(int32_t) * n_trees, st));
    FIT_TRY(S.alloc((void**)&rr_dev, sizeof(uint32_t) * n_trees));
    FIT_TRY(cudaMemcpyAsync(rr_dev, rr.data(), sizeof(uint32_t) * n_trees, cudaMemcpyHostToDevice, st));
    FIT_TRY(S.alloc((void**)&P.stack, sizeof(F16StackRec) * (size_t)n_trees * P.stack_cap));
    size_t buf_words = (size_t)n_trees * 2 * (best ? (size_t)d : 1) * (size_t)n;
    FIT_TRY(S.alloc((void**)&P.buf, sizeof(uint32_t) * buf_words));
    P.rand_r_state = rr_dev; P.nodes = F->nodes; P.node_count = F->node_count; P.err = F->err;

    if (kind == F16_KIND_RF) {
        size_t stride = align_up((size_t)n, 4);
        FIT_TRY(S.alloc((void**)&bw, stride * n_trees));
        rc = f16_bootstrap_counts(tree_seed.data(), n_trees, n, bw, (void*)st);
        if (rc) { forest_release(F, st); return rc; }
        P.boot_w = bw;
    }
    size_t dyn = 0;
    if (best) {
        FIT_TRY(S.alloc((void**)&P.lid, sizeof(uint32_t) * (size_t)n_trees * (size_t)n));
        P.side_words = (int)((n + 31) / 32);
        // F16_FORCE_GLOBAL_SIDE=1 exercises the > 524288-row path on small inputs (tests)
        static const int force_global_side = getenv("F16_FORCE_GLOBAL_SIDE") ? atoi(getenv("F16_FORCE_GLOBAL_SIDE")) : 0;
        if (P.side_words > F16_SIDE_SMEM_MAX_WORDS || force_global_side) {
            FIT_TRY(S.alloc((void**)&P.side_global, sizeof(uint32_t) * (size_t)n_trees * P.side_words));
        } else {
            dyn = sizeof(uint32_t) * (size_t)P.side_words;
        }
    }
    if (f16_get_profiling()) {
        FIT_TRY(cudaEventCreate(&F->ev0));
        cudaError_t e1 = cudaEventCreate(&F->ev1);
        if (e1 != cudaSuccess) { cudaEventDestroy(F->ev0); FIT_TRY(e1); }
        F->has_ev = 1;
        FIT_TRY(cudaEventRecord(F->ev0, st));
    }
    rc = (kind == F16_KIND_ET) ? f16_launch_build_random_et(P, st)
       : (kind == F16_KIND_RF) ? f16_launch_build_best_rf(P, dyn, st) : f16_launch_build_best_dt(P, dyn, st);
    if (rc) {
        f16_set_error("tree build kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        forest_release(F, st);
        return rc;
    }
    if (F->has_ev) FIT_TRY(cudaEventRecord(F->ev1, st));
#undef FIT_TRY
    *out = F;
    return F16_OK;
}

extern "C" int f16_forest_fit(const float* X_dev, const uint8_t* y_dev, int64_t n, int32_t d,
                              const int32_t* sorted_idx_dev, int32_t kind, int32_t n_estimators,
                              int32_t max_features, uint32_t seed, void* stream, f16_forest** out) {
    return forest_fit_impl(X_dev, y_dev, n, d, sorted_idx_dev, kind, n_estimators, max_features, seed, 0,
                           (cudaStream_t)stream, out);
}

// Same fit with an explicit per-tree node capacity (0 = the 2n - 1 worst case).  The worst case
// costs 32 B x (2n - 1) per tree - 57 GB for a 500-tree forest on 1.8 M rows - while real trees
// of this path hold 0.05 n - 0.15 n nodes; the grid engine passes a capacity derived from the
// node counts it has measured and retries with the worst case if a fit reports F16_ERR_OVERFLOW.
extern "C" int f16_forest_fit_cap(const float* X_dev, const uint8_t* y_dev, int64_t n, int32_t d,
                                  const int32_t* sorted_idx_dev, int32_t kind, int32_t n_estimators,
                                  int32_t max_features, uint32_t seed, int64_t node_cap, void* stream, f16_forest** out) {
    return forest_fit_impl(X_dev, y_dev, n, d, sorted_idx_dev, kind, n_estimators, max_features, seed, node_cap,
                           (cudaStream_t)stream, out);
}

extern "C" int f16_forest_predict(const f16_forest* F, const float* X_dev, int64_t n, uint8_t* pred_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (F && n == 0) return F16_OK;     // empty batch (pointers may be null)
    if (!F || !X_dev || !pred_dev || n < 0) { f16_set_error("f16_forest_predict: bad arguments"); return F16_ERR_INVALID; }
    int2* leaf = nullptr;
    CUDA_TRY(f16_malloc_async((void**)&leaf, sizeof(int2) * (size_t)F->n_trees * n, st));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)F->n_trees);
    if (F->dp == 8) k_predict_walk<8><<<grid, 256, 0, st>>>(F->nodes, F->node_cap, F->node_count, X_dev, (int)n, leaf);
    else k_predict_walk<16><<<grid, 256, 0, st>>>(F->nodes, F->node_cap, F->node_count, X_dev, (int)n, leaf);
    CUDA_TRY(cudaGetLastError());
    k_predict_reduce<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(leaf, (int)n, F->n_trees, pred_dev);
    f16_count_launch(2);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(leaf, st));
    return F16_OK;
}

// Synchronises `stream` and returns the device-side status of the fit (0 ok).
extern "C" int f16_forest_status(const f16_forest* F, void* stream) {
    if (!F) { f16_set_error("f16_forest_status: null forest"); return F16_ERR_INVALID; }
    int32_t ev[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(ev, F->err, sizeof(ev), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    const int32_t e = ev[0];
    const_cast<f16_forest*>(F)->max_nodes = ev[1];
    if (e) f16_set_error("forest fit failed on device: code %d (node/stack capacity or bootstrap weight > 127)", e);
    return e;
}

extern "C" int f16_forest_n_trees(const f16_forest* F) { return F ? F->n_trees : 0; }
extern "C" int f16_forest_max_nodes(const f16_forest* F) { return F ? F->max_nodes : 0; }

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

extern "C" void f16_forest_free(f16_forest* F, void* stream) { forest_release(F, (cudaStream_t)stream); }
