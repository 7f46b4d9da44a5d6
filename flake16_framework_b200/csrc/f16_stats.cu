// Spearman rank-correlation table of the feature matrix.
//   reference: write_figures, experiment.py:661-663 - scipy.stats.spearmanr(features).correlation
//   (scipy: rankdata(method="average") per column, then numpy.corrcoef of the ranks).
//
// Per column: the float64 values are mapped to order-preserving 64-bit keys and sorted (bitonic
// network: tiles of 1024 keys in shared memory, wider strides in global memory - all columns in one
// grid), every row then finds its tie group in the sorted column with two binary searches and takes
// the group's mean rank (lower + upper + 1) / 2.  The centred ranks are multiplied out into the
// d x d covariance by row blocks (partials summed in block order: deterministic) and normalised the
// way numpy.corrcoef does (cov / (n - 1), then / stddev_a / stddev_b, clipped to [-1, 1]).
#include "f16_common.cuh"
#include <math.h>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st);

#define SP_TILE 1024      // keys per shared-memory tile (512 threads, 2 keys each)
#define SP_COV_ROWS 4096  // rows per covariance block

__device__ __forceinline__ unsigned long long sp_key(double x) {
    x = x + 0.0;                                              // -0.0 -> +0.0: they tie, like x != y on the CPU
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void k_sp_keys(const double* __restrict__ X, long long n, int d, long long n_pad, unsigned long long* __restrict__ keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= n_pad) return;
    keys[(size_t)c * n_pad + i] = (i < n) ? sp_key(X[(size_t)i * d + c]) : ~0ull;
}

// all compare-exchange steps with stride < SP_TILE of the merges k_lo .. k_hi, inside one tile
__global__ void __launch_bounds__(SP_TILE / 2) k_sp_bitonic_local(unsigned long long* __restrict__ keys, long long n_pad,
                                                                  long long k_lo, long long k_hi) {
    __shared__ unsigned long long s[SP_TILE];
    unsigned long long* col = keys + (size_t)blockIdx.y * n_pad;
    const long long base = (long long)blockIdx.x * SP_TILE;
    const int t = threadIdx.x;
    s[t] = col[base + t];
    s[t + SP_TILE / 2] = col[base + t + SP_TILE / 2];
    __syncthreads();
    for (long long k = k_lo; k <= k_hi; k <<= 1) {
        for (int j = (int)min((long long)SP_TILE / 2, k >> 1); j > 0; j >>= 1) {
            const int i = 2 * t - (t & (j - 1));              // lower index of this thread's pair
            const int p = i + j;
            const bool up = (((base + i) & k) == 0);
            const unsigned long long a = s[i], b = s[p];
            if ((a > b) == up) { s[i] = b; s[p] = a; }
            __syncthreads();
        }
    }
    col[base + t] = s[t];
    col[base + t + SP_TILE / 2] = s[t + SP_TILE / 2];
}

__global__ void k_sp_bitonic_global(unsigned long long* __restrict__ keys, long long n_pad, long long j, long long k) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // pair index
    if (t >= n_pad / 2) return;
    unsigned long long* col = keys + (size_t)blockIdx.y * n_pad;
    const long long i = 2 * t - (t & (j - 1));
    const long long p = i + j;
    const bool up = ((i & k) == 0);
    const unsigned long long a = col[i], b = col[p];
    if ((a > b) == up) { col[i] = b; col[p] = a; }
}

// centred average ranks: R[i][c] = (lower + upper + 1) / 2 - (n + 1) / 2
__global__ void k_sp_rank(const double* __restrict__ X, long long n, int d, long long n_pad,
                          const unsigned long long* __restrict__ keys, double* __restrict__ R) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y;
    if (i >= n) return;
    const unsigned long long* col = keys + (size_t)c * n_pad;
    const unsigned long long key = sp_key(X[(size_t)i * d + c]);
    long long lo = 0, hi = n;                                  // first position with col[pos] >= key
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (col[mid] < key) lo = mid + 1; else hi = mid; }
    const long long lower = lo;
    hi = n;                                                    // first position with col[pos] > key
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (col[mid] <= key) lo = mid + 1; else hi = mid; }
    R[(size_t)i * d + c] = 0.5 * (double)(lower + lo + 1) - 0.5 * (double)(n + 1);
}

// partial[block][a][b] = sum over the block's rows of R[i][a] * R[i][b]
__global__ void __launch_bounds__(256) k_sp_cov_partial(const double* __restrict__ R, long long n, int d, double* __restrict__ partial) {
    __shared__ double s[64][F16_MAX_D + 1];
    const int t = threadIdx.x, a = t / F16_MAX_D, b = t % F16_MAX_D;
    const long long r0 = (long long)blockIdx.x * SP_COV_ROWS;
    const long long r1 = min(n, r0 + SP_COV_ROWS);
    double acc = 0.0;
    for (long long base = r0; base < r1; base += 64) {
        __syncthreads();
        for (int i = t; i < 64 * d; i += 256) {
            const int r = i / d, c = i - r * d;
            s[r][c] = (base + r < r1) ? R[(size_t)(base + r) * d + c] : 0.0;
        }
        __syncthreads();
        if (a < d && b < d) {
#pragma unroll 8
            for (int r = 0; r < 64; r++) acc = fma(s[r][a], s[r][b], acc);
        }
    }
    partial[(size_t)blockIdx.x * 256 + t] = acc;
}

__global__ void __launch_bounds__(256) k_sp_final(const double* __restrict__ partial, int n_blocks, long long n, int d, double* __restrict__ rho) {
    __shared__ double cov[F16_MAX_D][F16_MAX_D];
    const int t = threadIdx.x, a = t / F16_MAX_D, b = t % F16_MAX_D;
    double s = 0.0;
    for (int k = 0; k < n_blocks; k++) s += partial[(size_t)k * 256 + t];
    cov[a][b] = s / (double)(n - 1);
    __syncthreads();
    if (a < d && b < d) {
        double c = cov[a][b] / sqrt(cov[a][a]);
        c = c / sqrt(cov[b][b]);
        rho[a * d + b] = fmin(1.0, fmax(-1.0, c));              // NaN (a constant column) passes through, as in numpy
    }
}

extern "C" int f16_spearman(const double* X_dev, int64_t n, int32_t d, double* rho_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X_dev || !rho_dev || n < 2 || n > (1ll << 30) || d < 1 || d > F16_MAX_D) {
        f16_set_error("f16_spearman: bad arguments (n=%lld d=%d)", (long long)n, d);
        return F16_ERR_INVALID;
    }
    long long n_pad = SP_TILE;
    while (n_pad < n) n_pad <<= 1;
    unsigned long long* keys = nullptr;
    double *R = nullptr, *partial = nullptr;
    const int n_blocks = (int)((n + SP_COV_ROWS - 1) / SP_COV_ROWS);
    int rc = F16_OK, launches = 0;
    cudaError_t e;
    if ((e = f16_malloc_async((void**)&keys, sizeof(unsigned long long) * (size_t)d * n_pad, st)) != cudaSuccess ||
        (e = f16_malloc_async((void**)&R, sizeof(double) * (size_t)n * d, st)) != cudaSuccess ||
        (e = f16_malloc_async((void**)&partial, sizeof(double) * (size_t)n_blocks * 256, st)) != cudaSuccess) {
        f16_set_error("f16_spearman: %s", cudaGetErrorString(e));
        rc = F16_ERR_NOMEM;
    } else {
        dim3 gk((unsigned)((n_pad + 255) / 256), (unsigned)d);
        k_sp_keys<<<gk, 256, 0, st>>>(X_dev, n, d, n_pad, keys);
        dim3 gl((unsigned)(n_pad / SP_TILE), (unsigned)d), gg((unsigned)((n_pad / 2 + 255) / 256), (unsigned)d);
        k_sp_bitonic_local<<<gl, SP_TILE / 2, 0, st>>>(keys, n_pad, 2, SP_TILE);         // sorted tiles
        launches += 2;
        for (long long k = 2 * SP_TILE; k <= n_pad; k <<= 1) {
            for (long long j = k >> 1; j >= SP_TILE; j >>= 1) { k_sp_bitonic_global<<<gg, 256, 0, st>>>(keys, n_pad, j, k); launches++; }
            k_sp_bitonic_local<<<gl, SP_TILE / 2, 0, st>>>(keys, n_pad, k, k);
            launches++;
        }
        dim3 gr((unsigned)((n + 255) / 256), (unsigned)d);
        k_sp_rank<<<gr, 256, 0, st>>>(X_dev, n, d, n_pad, keys, R);
        k_sp_cov_partial<<<n_blocks, 256, 0, st>>>(R, n, d, partial);
        k_sp_final<<<1, 256, 0, st>>>(partial, n_blocks, n, d, rho_dev);
        launches += 3;
        f16_count_launch(launches);
        if ((e = cudaGetLastError()) != cudaSuccess) { f16_set_error("f16_spearman: %s", cudaGetErrorString(e)); rc = F16_ERR_CUDA; }
    }
    if (keys) cudaFreeAsync(keys, st);
    if (R) cudaFreeAsync(R, st);
    if (partial) cudaFreeAsync(partial, st);
    return rc;
}
