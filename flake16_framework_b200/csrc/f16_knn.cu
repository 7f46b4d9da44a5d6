// Exact brute-force k-nearest-neighbour search in float64 (Euclidean), k <= 8, d <= 16.
//
// Replaces sklearn.neighbors.NearestNeighbors(n_neighbors=K).fit(A).kneighbors(Q) as used
// by imbalanced-learn's SMOTE (K=6, minority rows only), TomekLinks (K=2) and
// EditedNearestNeighbours (K=4) - reference call site experiment.py:463-466; sklearn picks
// brute force (ArgKmin) for d > 15 and a KD-tree for d <= 15 (sklearn/neighbors/_base.py:615-648);
// both are exact searches, so one exact kernel serves both feature sets.
//
// The distance that DECIDES is the direct sum of squared differences accumulated in float64
// (no |x|^2 - 2xy + |y|^2 expansion: raw Flake16 columns reach 1e8 and the expansion cancels
// catastrophically).  Tie rule: smaller distance first, then lower reference index.
//
// This is an N^2*d FP64 CUDA-core problem (roofline: the FP64 pipe, not HBM).  To halve the
// FP64 instructions per pair, every pair first goes through a CONSERVATIVE FILTER in the
// expanded form: acc = x.y - (1-eps)/2 (|x|^2 + |y|^2), one DFMA per coordinate; a pair can
// only enter the top-k if acc > -bd_k/2, where eps bounds the rounding error of the expanded
// form (|error| <= 2(d+6) u (|x|^2+|y|^2), u = 2^-53; eps = 2e-14 >= 180 u).  The few pairs
// that pass are recomputed with the exact direct form, so the result is identical to the
// unfiltered search.  The filter is evaluated on the first half of the coordinates first
// (columns ordered by descending variance by the caller): a partial distance already beyond
// the current k-th best for every lane of the warp skips the second half.
// Reference rows are staged through shared memory together with their pre-scaled squared
// norms; every thread keeps its query and its running top-k in registers.
#include "f16_common.cuh"
#include <math.h>
#include <stdio.h>

extern "C" void f16_set_error(const char* fmt, ...);
#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define KT 128      // queries (threads) per block
#define KTILE 128   // reference rows per shared-memory tile (two tiles in flight)
#define KMAX 8
#define KNN_SCALE (-0.5 * (1.0 - 2.0e-14))

struct KnnPerm { int c[F16_MAX_D]; };

int f16_knn_tc_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                      cudaStream_t st);
int f16_knn_umma_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                        cudaStream_t st);
int f16_knn_sweep_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                         cudaStream_t st);

// first-half coordinate count: even, so that halves fall on 16-byte (double2) boundaries
template <int D> struct KnnCfg {
    static constexpr int H = 2;    // prefix tested first: the two highest-variance coordinates
    static constexpr int HH = (H < D) ? H : D;          // coordinates really in the first half
    static constexpr int DPAD = (D + 1) & ~1;           // coordinates padded to a pair
    static constexpr int TS = DPAD + 2;                 // tile row: {nh, nr, c0, c1, ..., pad}
};

// Reference rows in tile layout, written once: Ap[i] = {KNN_SCALE * |first H coords|^2,
// KNN_SCALE * |remaining coords|^2, c0, c1, ..., (pad)} with the coordinates in the permuted
// order, so the search kernel stages tiles with plain 16-byte asynchronous copies.
template <int D>
__global__ void k_knn_prep(const double* __restrict__ A, int n, KnnPerm perm, double* __restrict__ Ap) {
    constexpr int HH = KnnCfg<D>::HH, DPAD = KnnCfg<D>::DPAD, TS = KnnCfg<D>::TS;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double h = 0.0, r = 0.0;
    double* o = Ap + (size_t)i * TS;
#pragma unroll
    for (int c = 0; c < DPAD; c++) {
        double v = (c < D) ? A[(size_t)i * D + perm.c[c]] : 0.0;
        if (c < HH) h = fma(v, v, h); else r = fma(v, v, r);
        o[2 + c] = v;
    }
    o[0] = KNN_SCALE * h; o[1] = KNN_SCALE * r;
}

__device__ __forceinline__ void knn_cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void knn_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void knn_cp_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

// Each thread owns KQ = 2 queries (registers), so every 16-byte shared-memory load of a
// reference row feeds 4 DFMAs: the kernel stays bound by the FP64 pipe, not by LDS issue.
#ifndef KQ
#define KQ 2
#endif
#ifndef KNN_UNROLL
#define KNN_UNROLL 1
#endif
// PREFIX: test the two leading (highest-variance) coordinates first and skip the rest of the
// row when no lane of the warp can improve - pays off when those columns dominate the distance
// (raw, unscaled features); without it the row is one uninterrupted DFMA stream.
template <int D, int K, bool PREFIX>
__global__ void __launch_bounds__(KT) k_knn(const double* __restrict__ A, int n, const double* __restrict__ Q, int nq,
                                            int32_t* __restrict__ out, KnnPerm perm, const double* __restrict__ Ap) {
    constexpr int HH = KnnCfg<D>::HH, H2 = KnnCfg<D>::H / 2, DPAD = KnnCfg<D>::DPAD, P2 = DPAD / 2, TS = KnnCfg<D>::TS;
    __shared__ __align__(16) double tiles[2][KTILE * TS];      // double buffered (cp.async)
    const int tid = threadIdx.x;
    double q[KQ][DPAD], qh[KQ], qr[KQ], bd[KQ][K], thr[KQ];
    int bi[KQ][K], qi[KQ];
#pragma unroll
    for (int u = 0; u < KQ; u++) {
        qi[u] = blockIdx.x * (KT * KQ) + u * KT + tid;
        qh[u] = 0.0; qr[u] = 0.0;
#pragma unroll
        for (int c = 0; c < DPAD; c++) {
            q[u][c] = (c < D && qi[u] < nq) ? Q[(size_t)qi[u] * D + perm.c[c]] : 0.0;
            if (c < HH) qh[u] = fma(q[u][c], q[u][c], qh[u]); else qr[u] = fma(q[u][c], q[u][c], qr[u]);
        }
        qh[u] *= KNN_SCALE; qr[u] *= KNN_SCALE;
#pragma unroll
        for (int m = 0; m < K; m++) { bd[u][m] = INFINITY; bi[u][m] = -1; }
        thr[u] = -INFINITY;                   // -bd[K-1] / 2
    }

    // stage tile 0
    {
        const int cnt = min(KTILE, n);
        for (int i = tid; i < cnt * (TS / 2); i += KT) knn_cp_async16(&tiles[0][2 * i], Ap + 2 * (size_t)i);
        knn_cp_commit();
    }
    int buf = 0;
    for (int base = 0; base < n; base += KTILE, buf ^= 1) {
        const int cnt = min(KTILE, n - base);
        knn_cp_wait_all();
        __syncthreads();                  // tile `buf` is complete; everyone is done with `buf ^ 1`
        if (base + KTILE < n) {           // prefetch the next tile while this one is consumed
            const int ncnt = min(KTILE, n - base - KTILE);
            const double* g = Ap + (size_t)(base + KTILE) * TS;
            for (int i = tid; i < ncnt * (TS / 2); i += KT) knn_cp_async16(&tiles[buf ^ 1][2 * i], g + 2 * (size_t)i);
            knn_cp_commit();
        }
        const double* tile = tiles[buf];
        constexpr int kRowUnroll = KNN_UNROLL;
#pragma unroll kRowUnroll
        for (int j = 0; j < cnt; j++) {
            const double2* t2 = reinterpret_cast<const double2*>(tile + j * TS);
            const double2 nn = t2[0];
            double acc[KQ];
            const double2 v0 = t2[1];
            bool go = false;
#pragma unroll
            for (int u = 0; u < KQ; u++) {
                acc[u] = fma(q[u][1], v0.y, fma(q[u][0], v0.x, qh[u] + nn.x));
                go = go || (acc[u] > thr[u]);
            }
            if (!PREFIX || __any_sync(0xffffffffu, go)) {
                // remaining coordinates: two independent accumulators per query (the filter's
                // error bound does not depend on the summation order)
                double acb[KQ];
#pragma unroll
                for (int u = 0; u < KQ; u++) { acc[u] += qr[u] + nn.y; acb[u] = 0.0; }
#pragma unroll
                for (int p = H2; p < P2; p++) {
                    const double2 v = t2[1 + p];
#pragma unroll
                    for (int u = 0; u < KQ; u++) { acc[u] = fma(q[u][2 * p], v.x, acc[u]); acb[u] = fma(q[u][2 * p + 1], v.y, acb[u]); }
                }
#pragma unroll
                for (int u = 0; u < KQ; u++) acc[u] += acb[u];
#pragma unroll
                for (int u = 0; u < KQ; u++) {
                    if (acc[u] > thr[u]) {
                        // survivor of the conservative filter: the exact direct distance decides
                        const double* t = tile + j * TS + 2;
                        double s = 0.0;
#pragma unroll
                        for (int c = 0; c < D; c++) {
                            double df = q[u][c] - t[c];
                            s = fma(df, df, s);
                        }
                        if (s < bd[u][K - 1]) {
                            bd[u][K - 1] = s; bi[u][K - 1] = base + j;
#pragma unroll
                            for (int m = K - 1; m > 0; m--) {
                                if (bd[u][m] < bd[u][m - 1]) {
                                    double td = bd[u][m]; bd[u][m] = bd[u][m - 1]; bd[u][m - 1] = td;
                                    int ti = bi[u][m]; bi[u][m] = bi[u][m - 1]; bi[u][m - 1] = ti;
                                }
                            }
                            thr[u] = -0.5 * bd[u][K - 1];
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KQ; u++) {
        if (qi[u] < nq) {
#pragma unroll
            for (int m = 0; m < K; m++) out[(size_t)qi[u] * K + m] = bi[u][m];
        }
    }
}

template <int D>
static int launch_k(const double* A, int n, const double* Q, int nq, int k, int32_t* out, const KnnPerm& pm,
                    double* nrm, bool prefix, cudaStream_t st) {
    k_knn_prep<D><<<(n + 255) / 256, 256, 0, st>>>(A, n, pm, nrm);
    int grid = (nq + KT * KQ - 1) / (KT * KQ);
#define KNN_LAUNCH(KK)                                                                          \
    case KK:                                                                                    \
        if (prefix) k_knn<D, KK, true><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm, nrm);          \
        else k_knn<D, KK, false><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm, nrm);                \
        break;
    switch (k) {
        KNN_LAUNCH(1) KNN_LAUNCH(2) KNN_LAUNCH(3) KNN_LAUNCH(4) KNN_LAUNCH(5) KNN_LAUNCH(6) KNN_LAUNCH(7) KNN_LAUNCH(8)
        default: return F16_ERR_INVALID;
    }
#undef KNN_LAUNCH
    f16_count_launch(2);
    return F16_OK;
}

// idx_dev: [nq][k] int32, row-major; neighbours in increasing (distance, index) order.
// A query that is itself a row of A finds itself in column 0 (distance 0) unless an exact
// duplicate with a lower index exists - the same convention imbalanced-learn relies on.
// col_order (host, d ints, or NULL): order in which the coordinates are accumulated; put the
// highest-variance columns first to make the early exit effective.
extern "C" int f16_knn(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, int32_t k,
                       const int32_t* col_order, int32_t prefix_test, int32_t* idx_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!A_dev || !Q_dev || !idx_dev || n < 1 || nq < 0 || n > 0x7fffffff || nq > 0x7fffffff || d < 1 || d > F16_MAX_D) {
        f16_set_error("f16_knn: bad arguments"); return F16_ERR_INVALID;
    }
    if (k < 1 || k > KMAX || k > n) { f16_set_error("f16_knn: k=%d out of range (n=%lld)", k, (long long)n); return F16_ERR_INVALID; }
    if (nq == 0) return F16_OK;
    KnnPerm pm;
    unsigned seen = 0;
    for (int c = 0; c < F16_MAX_D; c++) pm.c[c] = c;
    if (col_order) {
        for (int c = 0; c < d; c++) {
            if (col_order[c] < 0 || col_order[c] >= d || ((seen >> col_order[c]) & 1u)) { f16_set_error("f16_knn: col_order is not a permutation"); return F16_ERR_INVALID; }
            seen |= 1u << col_order[c];
            pm.c[c] = col_order[c];
        }
    }
    if (prefix_test == 6) {     // sweep over the rows sorted by the leading (dominant) column (f16_knn_sweep.cu)
        int r6 = f16_knn_sweep_launch(A_dev, (int)n, Q_dev, (int)nq, d, k, pm.c, idx_dev, st);
        if (r6 == F16_OK) return F16_OK;
        if (r6 != F16_ERR_INVALID) return r6;
        prefix_test = 1;
    }
    if (prefix_test >= 3 && prefix_test <= 5) {     // tensor-core candidate filter + exact float64 selection
        // small problems are faster on the plain float64 kernel (three launches, candidate lists);
        // 4 / 5 = use the filter whatever the size (parity tests): 3, 4 tcgen05 + TMEM (f16_knn_umma.cu),
        // 5 mma.sync (f16_knn_tc.cu)
        if (prefix_test != 3 || (double)n * (double)nq >= 2.5e8) {
            int r3 = (prefix_test == 5) ? f16_knn_tc_launch(A_dev, (int)n, Q_dev, (int)nq, d, k, pm.c, idx_dev, st)
                                        : f16_knn_umma_launch(A_dev, (int)n, Q_dev, (int)nq, d, k, pm.c, idx_dev, st);
            if (r3 == F16_OK) return F16_OK;
            if (r3 != F16_ERR_INVALID) return r3;
        }
        prefix_test = 0;
    }
    if (prefix_test == 2) prefix_test = 0;      // (strategy 2, a float32 filter, was measured never faster and is no longer built)
    double* nrm = nullptr;      // reference rows in tile layout (norms + permuted coordinates)
    CUDA_TRY(cudaMallocAsync((void**)&nrm, sizeof(double) * (size_t)n * (F16_MAX_D + 2), st));
    int rc;
    switch (d) {
#define CASE_D(DD) case DD: rc = launch_k<DD>(A_dev, (int)n, Q_dev, (int)nq, k, idx_dev, pm, nrm, prefix_test != 0, st); break;
        CASE_D(1) CASE_D(2) CASE_D(3) CASE_D(4) CASE_D(5) CASE_D(6) CASE_D(7) CASE_D(8)
        CASE_D(9) CASE_D(10) CASE_D(11) CASE_D(12) CASE_D(13) CASE_D(14) CASE_D(15) CASE_D(16)
#undef CASE_D
        default: rc = F16_ERR_INVALID;
    }
    cudaError_t le = cudaGetLastError();
    cudaError_t fe = cudaFreeAsync(nrm, st);             // released on every path, failed launches included
    if (rc) { f16_set_error("f16_knn: unsupported (d=%d, k=%d)", d, k); return rc; }
    if (le != cudaSuccess || fe != cudaSuccess) {
        f16_set_error("f16_knn: %s", cudaGetErrorString(le != cudaSuccess ? le : fe));
        return F16_ERR_CUDA;
    }
    return F16_OK;
}
