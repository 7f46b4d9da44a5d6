// Exact k-NN by a sweep over the rows sorted by ONE dominant coordinate (f16_knn, strategy 6).
//
// For raw, unscaled features one column carries almost all of the variance (Flake16's peak-memory
// column spans 1e3 .. 1e9), so the tensor-core filter's norm-relative band is useless and the
// float64 search tests all n^2 pairs.  But then the k nearest rows are also near in that one
// coordinate: with the reference rows sorted by it, a query walks outwards from its own position,
// left and right, and stops on a side as soon as the squared gap in the sorted coordinate alone
// exceeds the current k-th best distance - a few hundred rows instead of n.  Every visited pair is
// evaluated with the same float64 direct sum, in the same coordinate order, as k_knn, and kept by
// (distance, index): identical output.
//
// The rows are ordered by the float32 rounding of the coordinate (the library's radix argsort,
// f16_sort.cu), so the float64 values are only nearly sorted; the stopping rule therefore uses a
// bound that is monotone in the float32 key: every row at or beyond position j has a float64
// coordinate >= key_j - |key_j| 2^-23 (right side; mirrored on the left).
// Whether this strategy pays is decided per dataset by measurement (ops.calibrate_knn).
#include "f16_common.cuh"
#include <math.h>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st);
extern "C" int f16_argsort_columns(const float* X_dev, int64_t n, int32_t d, int32_t* sorted_idx_dev, void* stream);
#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

struct SwPerm { int c[F16_MAX_D]; };

// float32 key rows for the radix argsort: [n][8], column 0 = the dominant coordinate
__global__ void k_sweep_keys(const double* __restrict__ A, int n, int d, int c0, float* __restrict__ keyrow) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4* o = reinterpret_cast<float4*>(keyrow + (size_t)i * 8);
    o[0] = make_float4((float)A[(size_t)i * d + c0], 0.f, 0.f, 0.f);
    o[1] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// rows in sorted order, coordinates permuted (perm[0] = dominant), padded to 16; key as float64
__global__ void k_sweep_gather(const double* __restrict__ A, int n, int d, SwPerm perm, const int32_t* __restrict__ order,
                               double* __restrict__ S, double* __restrict__ key) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = order[p];
    const double* row = A + (size_t)i * d;
    double* o = S + (size_t)p * F16_MAX_D;
#pragma unroll
    for (int c = 0; c < F16_MAX_D; c++) o[c] = (c < d) ? row[perm.c[c]] : 0.0;
    key[p] = (double)(float)row[perm.c[0]];
}

template <int K, bool SAME>
__global__ void __launch_bounds__(128) k_sweep_search(const double* __restrict__ S, const double* __restrict__ key,
                                                      const int32_t* __restrict__ order, int n,
                                                      const double* __restrict__ Q, int nq, int d, SwPerm perm,
                                                      int32_t* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    double q[F16_MAX_D];
    int qi, p0;
    if (SAME) {
        // query t is the row at sorted position t: neighbouring threads sweep neighbouring rows
        qi = order[t];
        p0 = t;
#pragma unroll
        for (int c = 0; c < F16_MAX_D; c++) q[c] = S[(size_t)t * F16_MAX_D + c];
    } else {
        qi = t;
#pragma unroll
        for (int c = 0; c < F16_MAX_D; c++) q[c] = (c < d) ? Q[(size_t)t * d + perm.c[c]] : 0.0;
        const double qk = (double)(float)q[0];
        int lo = 0, hi = n;                       // first position with key >= qk
        while (lo < hi) { int mid = (lo + hi) >> 1; if (key[mid] < qk) lo = mid + 1; else hi = mid; }
        p0 = lo;
    }
    const double q0 = q[0];
    double bd[K]; int bi[K];
#pragma unroll
    for (int k = 0; k < K; k++) { bd[k] = INFINITY; bi[k] = 0x7fffffff; }
    int r = p0, l = p0 - 1;
    const double ULP = 1.1920928955078125e-07;    // 2^-23
    while (true) {
        // lower bounds of the squared distance of every row at or beyond r (right) / l (left)
        double br = INFINITY, bl = INFINITY;
        // (1e-37 covers float32 subnormal keys; a non-finite key gives no bound at all)
        if (r < n) { double kk = key[r]; double g = (kk - fabs(kk) * ULP - 1e-37) - q0; br = (isfinite(kk) && g > 0.0) ? g * g : 0.0; }
        if (l >= 0) { double kk = key[l]; double g = q0 - (kk + fabs(kk) * ULP + 1e-37); bl = (isfinite(kk) && g > 0.0) ? g * g : 0.0; }
        const bool go_r = (r < n) && !(br > bd[K - 1]);
        const bool go_l = (l >= 0) && !(bl > bd[K - 1]);
        if (!go_r && !go_l) break;
        const int p = (go_r && (!go_l || br <= bl)) ? r++ : l--;
        const double* row = S + (size_t)p * F16_MAX_D;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < F16_MAX_D; c++) {
            if (c < d) { double df = q[c] - row[c]; s = fma(df, df, s); }
        }
        const int j = order[p];
        if (s < bd[K - 1] || (s == bd[K - 1] && j < bi[K - 1])) {
            bd[K - 1] = s; bi[K - 1] = j;
#pragma unroll
            for (int k = K - 1; k > 0; k--) {
                if (bd[k] < bd[k - 1] || (bd[k] == bd[k - 1] && bi[k] < bi[k - 1])) {
                    double td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                    int ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; k++) out[(size_t)qi * K + k] = (bi[k] == 0x7fffffff) ? -1 : bi[k];
}

// Returns F16_OK, or F16_ERR_INVALID when the caller should use another strategy.
int f16_knn_sweep_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                         cudaStream_t st) {
    if (k < 1 || k > 8 || d < 1 || d > F16_MAX_D) return F16_ERR_INVALID;
    const bool same = (A == Q) && (n == nq);
    SwPerm pm;
    for (int c = 0; c < F16_MAX_D; c++) pm.c[c] = (c < d) ? perm[c] : c;
    float* keyrow = nullptr; int32_t* order = nullptr; double *S = nullptr, *key = nullptr;
    CUDA_TRY(f16_malloc_async((void**)&keyrow, sizeof(float) * 8 * (size_t)n, st));
    CUDA_TRY(f16_malloc_async((void**)&order, sizeof(int32_t) * (size_t)n, st));
    CUDA_TRY(f16_malloc_async((void**)&S, sizeof(double) * F16_MAX_D * (size_t)n, st));
    CUDA_TRY(f16_malloc_async((void**)&key, sizeof(double) * (size_t)n, st));
    k_sweep_keys<<<(n + 255) / 256, 256, 0, st>>>(A, n, d, pm.c[0], keyrow);
    int rc = f16_argsort_columns(keyrow, n, 1, order, st);
    if (rc == F16_OK) {
        k_sweep_gather<<<(n + 255) / 256, 256, 0, st>>>(A, n, d, pm, order, S, key);
        const int grid = (nq + 127) / 128;
#define SW_LAUNCH(KK)                                                                                              \
    case KK:                                                                                                       \
        if (same) k_sweep_search<KK, true><<<grid, 128, 0, st>>>(S, key, order, n, Q, nq, d, pm, out);              \
        else k_sweep_search<KK, false><<<grid, 128, 0, st>>>(S, key, order, n, Q, nq, d, pm, out);                  \
        break;
        switch (k) { SW_LAUNCH(1) SW_LAUNCH(2) SW_LAUNCH(3) SW_LAUNCH(4) SW_LAUNCH(5) SW_LAUNCH(6) SW_LAUNCH(7) SW_LAUNCH(8) }
#undef SW_LAUNCH
        f16_count_launch(3);
    }
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(keyrow, st); cudaFreeAsync(order, st); cudaFreeAsync(S, st); cudaFreeAsync(key, st);
    if (rc != F16_OK) return rc;
    if (e != cudaSuccess) { f16_set_error("f16_knn (sorted sweep): %s", cudaGetErrorString(e)); return F16_ERR_CUDA; }
    return F16_OK;
}
