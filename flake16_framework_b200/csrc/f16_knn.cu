// Exact brute-force k-nearest-neighbour search in float64 (Euclidean), k <= 8, d <= 16.
//
// Replaces sklearn.neighbors.NearestNeighbors(n_neighbors=K).fit(A).kneighbors(Q) as used
// by imbalanced-learn's SMOTE (K=6, minority rows only), TomekLinks (K=2) and
// EditedNearestNeighbours (K=4) - reference call site experiment.py:463-466; sklearn picks
// brute force (ArgKmin) for d > 15 and a KD-tree for d <= 15 (sklearn/neighbors/_base.py:615-648);
// both are exact searches, so one exact kernel serves both feature sets.
//
// Distances are direct sums of squared differences accumulated in float64 (no
// |x|^2 - 2xy + |y|^2 expansion: raw Flake16 columns reach 1e8 and the expansion cancels
// catastrophically).  Tie rule: smaller distance first, then lower reference index.
//
// This is an N^2*d FP64 CUDA-core problem (no tensor-core path for float64 on sm_100a that
// keeps exactness); the reference rows are staged through shared memory and every thread
// keeps its query and its running top-k in registers.
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
#define KTILE 64    // reference rows per shared-memory tile
#define KMAX 8

struct KnnPerm { int c[F16_MAX_D]; };

// Partial-distance early exit: the first H coordinates (the caller orders columns by
// descending variance) are accumulated first; the remaining D-H are only evaluated when at
// least one query of the warp could still beat its current k-th best distance.  The result
// is exact: a skipped pair has a partial sum already >= the k-th best of every lane.
template <int D, int K>
__global__ void __launch_bounds__(KT) k_knn(const double* __restrict__ A, int n, const double* __restrict__ Q, int nq,
                                            int32_t* __restrict__ out, KnnPerm perm) {
    constexpr int H = (D + 1) / 2;
    __shared__ double tile[KTILE * D];
    const int tid = threadIdx.x;
    const int qi = blockIdx.x * KT + tid;
    double q[D];
#pragma unroll
    for (int c = 0; c < D; c++) q[c] = (qi < nq) ? Q[(size_t)qi * D + perm.c[c]] : 0.0;
    double bd[K];
    int bi[K];
#pragma unroll
    for (int m = 0; m < K; m++) { bd[m] = INFINITY; bi[m] = -1; }

    for (int base = 0; base < n; base += KTILE) {
        int cnt = min(KTILE, n - base);
        for (int i = tid; i < cnt * D; i += KT) {
            int j = i / D, c = i - j * D;
            tile[i] = A[(size_t)(base + j) * D + perm.c[c]];
        }
        __syncthreads();
        for (int j = 0; j < cnt; j++) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < H; c++) {
                double df = q[c] - tile[j * D + c];
                s = fma(df, df, s);
            }
            if (__any_sync(0xffffffffu, s < bd[K - 1])) {
#pragma unroll
                for (int c = H; c < D; c++) {
                    double df = q[c] - tile[j * D + c];
                    s = fma(df, df, s);
                }
                if (s < bd[K - 1]) {
                    bd[K - 1] = s; bi[K - 1] = base + j;
#pragma unroll
                    for (int m = K - 1; m > 0; m--) {
                        if (bd[m] < bd[m - 1]) {
                            double td = bd[m]; bd[m] = bd[m - 1]; bd[m - 1] = td;
                            int ti = bi[m]; bi[m] = bi[m - 1]; bi[m - 1] = ti;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (qi < nq) {
#pragma unroll
        for (int m = 0; m < K; m++) out[(size_t)qi * K + m] = bi[m];
    }
}

template <int D>
static int launch_k(const double* A, int n, const double* Q, int nq, int k, int32_t* out, const KnnPerm& pm, cudaStream_t st) {
    int grid = (nq + KT - 1) / KT;
    switch (k) {
        case 2: k_knn<D, 2><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 4: k_knn<D, 4><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 6: k_knn<D, 6><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 1: k_knn<D, 1><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 3: k_knn<D, 3><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 5: k_knn<D, 5><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 7: k_knn<D, 7><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        case 8: k_knn<D, 8><<<grid, KT, 0, st>>>(A, n, Q, nq, out, pm); break;
        default: return F16_ERR_INVALID;
    }
    f16_count_launch(1);
    return F16_OK;
}

// idx_dev: [nq][k] int32, row-major; neighbours in increasing (distance, index) order.
// A query that is itself a row of A finds itself in column 0 (distance 0) unless an exact
// duplicate with a lower index exists - the same convention imbalanced-learn relies on.
// col_order (host, d ints, or NULL): order in which the coordinates are accumulated; put the
// highest-variance columns first to make the early exit effective.
extern "C" int f16_knn(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, int32_t k,
                       const int32_t* col_order, int32_t* idx_dev, void* stream) {
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
    int rc;
    switch (d) {
#define CASE_D(DD) case DD: rc = launch_k<DD>(A_dev, (int)n, Q_dev, (int)nq, k, idx_dev, pm, st); break;
        CASE_D(1) CASE_D(2) CASE_D(3) CASE_D(4) CASE_D(5) CASE_D(6) CASE_D(7) CASE_D(8)
        CASE_D(9) CASE_D(10) CASE_D(11) CASE_D(12) CASE_D(13) CASE_D(14) CASE_D(15) CASE_D(16)
#undef CASE_D
        default: rc = F16_ERR_INVALID;
    }
    if (rc) { f16_set_error("f16_knn: unsupported (d=%d, k=%d)", d, k); return rc; }
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}
