// Small data-movement / bookkeeping kernels of the `scores` hot path + library plumbing.
//   gather+cast   experiment.py:459-460 (features[train]) fused with sklearn's float32 cast
//                 (sklearn/ensemble/_forest.py:334-342 validate_data dtype=float32)
//   SMOTE / TomekLinks / ENN decisions + row compaction   experiment.py:463-466
//                 (imbalanced-learn 0.9.0)
//   confusion counts   experiment.py:476-483
// Compile with -fmad=false (SMOTE's `X[rows] + steps * diffs` rounds twice on the CPU).
#include "f16_common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

// ------------------------------------------------------------------ error plumbing
static thread_local char g_err[512] = "";

extern "C" void f16_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* f16_last_error(void) { return g_err; }
extern "C" int f16_version(void) { return 100; }

#include <atomic>
static std::atomic<long long> g_launches{0};
extern "C" void f16_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long f16_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}
static std::atomic<int> g_profiling{0};
extern "C" void f16_set_profiling(int on) { g_profiling.store(on); }
extern "C" int f16_get_profiling(void) { return g_profiling.load(); }

#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

// Keeps stream-ordered allocations cached in the device pool (no trim at synchronisation).
extern "C" int f16_init(int device) {
    CUDA_TRY(cudaSetDevice(device));
    cudaMemPool_t pool;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thr = ~0ull;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    // Many forests are fitted concurrently on different streams.  With the default
    // cudaMemPoolReuseAllowInternalDependencies the pool satisfies an allocation on stream A with
    // a block whose free is still pending on stream B by inserting a hidden B -> A dependency,
    // which serialises the streams (measured: 4.4-5.3 s vs 2.8-2.9 s per 18-config grid slice).
    // With it off every stream keeps its own blocks (bounded by lanes x per-fit scratch, ~100 GB
    // worst case at 178 k rows); f16_malloc_async trims the pool and retries if that ever
    // exhausts the device.
    int off = 0;
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
    return F16_OK;
}

// Stream-ordered allocation that survives pool exhaustion: on cudaErrorMemoryAllocation the device
// is synchronised, every cached block of the pool is released and the request retried once.
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st) {
    cudaError_t e = cudaMallocAsync(p, bytes, st);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool;
        if (cudaDeviceSynchronize() == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            cudaMemPoolTrimTo(pool, 0);
            e = cudaMallocAsync(p, bytes, st);
        }
    }
    return e;
}

// ------------------------------------------------------------------ gather rows (+ cast)
// out[j][0..dp) = (float) X[idx ? idx[j] : j][0..d), zero padded to dp columns.
__global__ void k_gather_cast(const double* __restrict__ X, int d, const int64_t* __restrict__ idx, int64_t n_out,
                              int dp, float* __restrict__ out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t j = g / dp;
    int c = (int)(g % dp);
    if (j >= n_out) return;
    int64_t r = idx ? idx[j] : j;
    out[g] = (c < d) ? __double2float_rn(X[r * d + c]) : 0.f;
}
__global__ void k_gather_f64(const double* __restrict__ X, int d, const int64_t* __restrict__ idx, int64_t n_out,
                             double* __restrict__ out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t j = g / d;
    int c = (int)(g % d);
    if (j >= n_out) return;
    out[g] = X[idx[j] * d + c];
}
__global__ void k_gather_u8(const uint8_t* __restrict__ y, const int64_t* __restrict__ idx, int64_t n_out,
                            uint8_t* __restrict__ out) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n_out) out[j] = y[idx[j]];
}

extern "C" int f16_gather_rows_f32(const double* X_dev, int32_t d, const int64_t* idx_dev, int64_t n_out,
                                   float* out_dev, void* stream) {
    if (n_out == 0) return F16_OK;        // empty batch: nothing to stage (pointers may be null)
    if (!X_dev || !out_dev || d < 1 || d > F16_MAX_D || n_out < 0) { f16_set_error("f16_gather_rows_f32: bad arguments"); return F16_ERR_INVALID; }
    int dp = (d <= 8) ? 8 : 16;
    int64_t tot = n_out * dp;
    k_gather_cast<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X_dev, d, idx_dev, n_out, dp, out_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}
extern "C" int f16_gather_rows_f64(const double* X_dev, int32_t d, const int64_t* idx_dev, int64_t n_out,
                                   double* out_dev, void* stream) {
    if (!X_dev || !out_dev || !idx_dev || d < 1 || n_out < 0) { f16_set_error("f16_gather_rows_f64: bad arguments"); return F16_ERR_INVALID; }
    if (n_out == 0) return F16_OK;
    int64_t tot = n_out * d;
    k_gather_f64<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X_dev, d, idx_dev, n_out, out_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}
extern "C" int f16_gather_u8(const uint8_t* y_dev, const int64_t* idx_dev, int64_t n_out, uint8_t* out_dev, void* stream) {
    if (!y_dev || !out_dev || !idx_dev || n_out < 0) { f16_set_error("f16_gather_u8: bad arguments"); return F16_ERR_INVALID; }
    if (n_out == 0) return F16_OK;
    k_gather_u8<<<(unsigned)((n_out + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y_dev, idx_dev, n_out, out_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}

__global__ void k_gather_i32(const int32_t* __restrict__ v, const int64_t* __restrict__ idx, int64_t n, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v[idx[i]];
}
extern "C" int f16_gather_i32(const int32_t* v_dev, const int64_t* idx_dev, int64_t n_out, int32_t* out_dev, void* stream) {
    if (!v_dev || !out_dev || !idx_dev || n_out < 0) { f16_set_error("f16_gather_i32: bad arguments"); return F16_ERR_INVALID; }
    if (n_out == 0) return F16_OK;
    k_gather_i32<<<(unsigned)((n_out + 255) / 256), 256, 0, (cudaStream_t)stream>>>(v_dev, idx_dev, n_out, out_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}

// labels of the synthetic rows SMOTE appends (imblearn: y_new = full(n_new, minority class))
extern "C" int f16_fill_u8(uint8_t* dst_dev, int32_t value, int64_t n, void* stream) {
    if (n < 0 || (n > 0 && !dst_dev)) { f16_set_error("f16_fill_u8: bad arguments"); return F16_ERR_INVALID; }
    if (n == 0) return F16_OK;
    CUDA_TRY(cudaMemsetAsync(dst_dev, value, (size_t)n, (cudaStream_t)stream));
    return F16_OK;
}

// ------------------------------------------------------------------ SMOTE interpolation
// X_new[j] = C[row] + step[j] * (C[nn[row][1 + col]] - C[row]),  row = idx[j] / k, col = idx[j] % k
// (imblearn/over_sampling/_smote/base.py _make_samples/_generate_samples); nn has k+1
// columns, column 0 being the query row itself.
__global__ void k_smote(const double* __restrict__ C, int d, const int32_t* __restrict__ nn, int k,
                        const int64_t* __restrict__ sidx, const double* __restrict__ steps, int64_t n_new,
                        double* __restrict__ out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t j = g / d;
    int c = (int)(g % d);
    if (j >= n_new) return;
    int64_t s = sidx[j];
    int64_t row = s / k;
    int col = (int)(s % k);
    int32_t nb = nn[row * (k + 1) + 1 + col];
    double a = C[row * d + c];
    double diff = __dsub_rn(C[(int64_t)nb * d + c], a);
    out[g] = __dadd_rn(a, __dmul_rn(steps[j], diff));
}
extern "C" int f16_smote_generate(const double* C_dev, int64_t n_min, int32_t d, const int32_t* nn_dev, int32_t k,
                                  const int64_t* sample_idx_dev, const double* steps_dev, int64_t n_new,
                                  double* Xnew_dev, void* stream) {
    if (!C_dev || !nn_dev || !sample_idx_dev || !steps_dev || !Xnew_dev || n_min < 1 || d < 1 || k < 1 || n_new < 0) {
        f16_set_error("f16_smote_generate: bad arguments"); return F16_ERR_INVALID;
    }
    if (n_new == 0) return F16_OK;
    int64_t tot = n_new * d;
    k_smote<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(C_dev, d, nn_dev, k, sample_idx_dev, steps_dev, n_new, Xnew_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}

// ------------------------------------------------------------------ Tomek links / ENN keep masks
// clean_mask: bit c set <=> class c is cleaned.  nn: [n][kk] neighbour table incl. self in col 0.
__global__ void k_tomek_keep(const int32_t* __restrict__ nn, int kk, const uint8_t* __restrict__ y, int64_t n,
                             int clean_mask, uint8_t* __restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int yi = y[i];
    bool link = false;
    if ((clean_mask >> yi) & 1) {
        int32_t j = nn[i * kk + 1];
        if (y[j] != yi && nn[(int64_t)j * kk + 1] == (int32_t)i) link = true;
    }
    keep[i] = link ? 0 : 1;
}
__global__ void k_enn_keep(const int32_t* __restrict__ nn, int kk, const uint8_t* __restrict__ y, int64_t n,
                           int clean_mask, uint8_t* __restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int yi = y[i];
    bool ok = true;
    if ((clean_mask >> yi) & 1) {
        for (int c = 1; c < kk; c++) ok = ok && (y[nn[i * kk + c]] == yi);
    }
    keep[i] = ok ? 1 : 0;
}
extern "C" int f16_tomek_keep(const int32_t* nn_dev, int32_t kk, const uint8_t* y_dev, int64_t n, int32_t clean_mask,
                              uint8_t* keep_dev, void* stream) {
    if (!nn_dev || !y_dev || !keep_dev || kk < 2 || n < 0) { f16_set_error("f16_tomek_keep: bad arguments"); return F16_ERR_INVALID; }
    if (n == 0) return F16_OK;
    k_tomek_keep<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nn_dev, kk, y_dev, n, clean_mask, keep_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}
extern "C" int f16_enn_keep(const int32_t* nn_dev, int32_t kk, const uint8_t* y_dev, int64_t n, int32_t clean_mask,
                            uint8_t* keep_dev, void* stream) {
    if (!nn_dev || !y_dev || !keep_dev || kk < 2 || n < 0) { f16_set_error("f16_enn_keep: bad arguments"); return F16_ERR_INVALID; }
    if (n == 0) return F16_OK;
    k_enn_keep<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(nn_dev, kk, y_dev, n, clean_mask, keep_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}

// ------------------------------------------------------------------ row compaction
// grouped = 0: kept rows in original order (TomekLinks: flatnonzero(~links)).
// grouped = 1: kept class-0 rows first, then kept class-1 rows (ENN concatenates per class).
#define CB 1024
__global__ void __launch_bounds__(CB) k_compact_count(const uint8_t* __restrict__ keep, const uint8_t* __restrict__ y,
                                                      int64_t n, unsigned long long* __restrict__ block_sums) {
    __shared__ unsigned long long s[CB / 32];
    int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
    unsigned long long v = 0;
    if (i < n && keep[i]) v = y[i] ? (1ull << 32) : 1ull;
    v = f16_warp_sum_u64(v);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int q = 0; q < CB / 32; q++) t += s[q];
        block_sums[blockIdx.x] = t;
    }
}
// single block: exclusive scan of block sums; total -> block_sums[nb]
__global__ void __launch_bounds__(1024) k_compact_scan(unsigned long long* block_sums, int nb, int64_t* n_out) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + threadIdx.x;
        unsigned long long v = (i < nb) ? block_sums[i] : 0ull;
        unsigned long long incl = f16_warp_incl_scan_u64(v);
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned long long before = 0;
        for (int q = 0; q < (threadIdx.x >> 5); q++) before += s_w[q];
        unsigned long long carry = s_carry;
        if (i < nb) block_sums[i] = carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        unsigned long long t = s_carry;
        block_sums[nb] = t;
        n_out[0] = (int64_t)((t & 0xffffffffull) + (t >> 32));
        n_out[1] = (int64_t)(t & 0xffffffffull);
    }
}
__global__ void __launch_bounds__(CB) k_compact_scatter(const double* __restrict__ X, const uint8_t* __restrict__ y,
                                                        const uint8_t* __restrict__ keep, int64_t n, int d, int grouped,
                                                        const unsigned long long* __restrict__ block_sums, int nb,
                                                        double* __restrict__ Xout, uint8_t* __restrict__ yout,
                                                        int64_t* __restrict__ src_index) {
    __shared__ unsigned long long s_w[CB / 32];
    int64_t i = (int64_t)blockIdx.x * CB + threadIdx.x;
    bool k = (i < n) && keep[i];
    int yi = (i < n) ? y[i] : 0;
    unsigned long long v = k ? (yi ? (1ull << 32) : 1ull) : 0ull;
    unsigned long long incl = f16_warp_incl_scan_u64(v);
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long before = block_sums[blockIdx.x];
    for (int q = 0; q < (threadIdx.x >> 5); q++) before += s_w[q];
    if (!k) return;
    unsigned long long ex = before + incl - v;
    unsigned long long tot = block_sums[nb];
    int64_t pos;
    if (grouped) pos = yi ? (int64_t)((tot & 0xffffffffull) + (ex >> 32)) : (int64_t)(ex & 0xffffffffull);
    else pos = (int64_t)((ex & 0xffffffffull) + (ex >> 32));
    for (int c = 0; c < d; c++) Xout[pos * d + c] = X[i * d + c];
    yout[pos] = (uint8_t)yi;
    if (src_index) src_index[pos] = i;
}
// n_out_dev: int64[2] = {rows kept, class-0 rows kept}.  scratch: (ceil(n/1024)+1) u64.
extern "C" int f16_compact_rows(const double* X_dev, const uint8_t* y_dev, const uint8_t* keep_dev, int64_t n, int32_t d,
                                int32_t grouped, double* Xout_dev, uint8_t* yout_dev, int64_t* src_index_dev,
                                int64_t* n_out_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X_dev || !y_dev || !keep_dev || !Xout_dev || !yout_dev || !n_out_dev || n < 1 || d < 1) {
        f16_set_error("f16_compact_rows: bad arguments"); return F16_ERR_INVALID;
    }
    int nb = (int)((n + CB - 1) / CB);
    unsigned long long* bs = nullptr;
    CUDA_TRY(cudaMallocAsync((void**)&bs, sizeof(unsigned long long) * (nb + 1), st));
    k_compact_count<<<nb, CB, 0, st>>>(keep_dev, y_dev, n, bs);
    f16_count_launch(1);
    k_compact_scan<<<1, 1024, 0, st>>>(bs, nb, n_out_dev);
    f16_count_launch(1);
    k_compact_scatter<<<nb, CB, 0, st>>>(X_dev, y_dev, keep_dev, n, d, grouped, bs, nb, Xout_dev, yout_dev, src_index_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(bs, st));
    return F16_OK;
}

// ------------------------------------------------------------------ confusion counts
// k = 2*label + pred - 1: 0 FP, 1 FN, 2 TP (-1 TN skipped); per project and total
// (experiment.py:476-483).  counts: int64 [n_proj + 1][3], ACCUMULATED (folds add up).
__global__ void k_confusion(const uint8_t* __restrict__ y, const uint8_t* __restrict__ pred,
                            const int32_t* __restrict__ proj, int64_t n, int n_proj,
                            unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int s_c[];   // [(n_proj + 1) * 3]
    int m = (n_proj + 1) * 3;
    for (int i = threadIdx.x; i < m; i += blockDim.x) s_c[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int k = 2 * (int)y[i] + (int)pred[i] - 1;
        if (k < 0) continue;
        atomicAdd(&s_c[proj[i] * 3 + k], 1u);
        atomicAdd(&s_c[n_proj * 3 + k], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        if (s_c[i]) atomicAdd(&counts[i], (unsigned long long)s_c[i]);
}
extern "C" int f16_confusion(const uint8_t* y_dev, const uint8_t* pred_dev, const int32_t* proj_dev, int64_t n,
                             int32_t n_proj, int64_t* counts_dev, void* stream) {
    if (!y_dev || !pred_dev || !proj_dev || !counts_dev || n < 0 || n_proj < 1 || n_proj > 4000) {
        f16_set_error("f16_confusion: bad arguments"); return F16_ERR_INVALID;
    }
    if (n == 0) return F16_OK;
    int grid = (int)((n + 255) / 256);
    if (grid > 296) grid = 296;
    k_confusion<<<grid, 256, sizeof(unsigned int) * (n_proj + 1) * 3, (cudaStream_t)stream>>>(
        y_dev, pred_dev, proj_dev, n, n_proj, (unsigned long long*)counts_dev);
    f16_count_launch(1);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}
