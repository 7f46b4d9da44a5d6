// Per-column argsort of the float32 training matrix: LSD radix sort, 8-bit digits, 4 passes,
// all d columns in one launch (blockIdx.y = column).
//
// Replaces the per-node `simultaneous_sort` introsort of sklearn's BestSplitter
// (sklearn/tree/_partitioner.pyx:59-109 sort_samples_and_feature_values,
// sklearn/utils/_sorting.pyx:8-157): the best-split builder keeps one index array per
// feature sorted by value for the whole tree and only partitions it stably at each split,
// so each training set is sorted exactly once per column instead of once per (node, feature).
// Order among equal values is irrelevant to the tree.
#include "f16_common.cuh"
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

#define ST 256                 // threads per block
#define SW (ST / 32)
#define ITEMS 16
#define CHUNK (ST * ITEMS)     // elements per block

__device__ __forceinline__ uint32_t f32_key(float v) {
    uint32_t u = __float_as_uint(v);
    return u ^ ((u >> 31) ? 0xffffffffu : 0x80000000u);
}

// first pass reads the keys straight from the row-major matrix
__device__ __forceinline__ uint32_t load_key(const float* X, int dp, int col, const uint32_t* keys, int64_t i, bool first) {
    return first ? f32_key(X[i * dp + col]) : keys[i];
}

__global__ void __launch_bounds__(ST) k_radix_hist(const float* __restrict__ X, int dp, const uint32_t* __restrict__ keys_in,
                                                   int n, int shift, int first, uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t h[256];
    const int col = blockIdx.y;
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t* kin = keys_in + (size_t)col * n;
    int base = blockIdx.x * CHUNK;
    for (int r = 0; r < ITEMS; r++) {
        int i = base + r * ST + threadIdx.x;
        if (i < n) atomicAdd(&h[(load_key(X, dp, col, kin, i, first) >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[((size_t)col * 256 + threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

// one block per column: exclusive scan over (digit major, block minor)
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    uint32_t* h = hist + (size_t)blockIdx.x * 256 * nblocks;
    int m = 256 * nblocks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        int i = base + threadIdx.x;
        uint32_t v = (i < m) ? h[i] : 0u, incl = v;
        int lane = threadIdx.x & 31;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t t = __shfl_up_sync(F16_FULL, incl, dlt); if (lane >= dlt) incl += t; }
        if (lane == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (int q = 0; q < (threadIdx.x >> 5); q++) before += s_w[q];
        uint32_t carry = s_carry;
        if (i < m) h[i] = carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + before + incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(ST) k_radix_scatter(const float* __restrict__ X, int dp, const uint32_t* __restrict__ keys_in,
                                                      const uint32_t* __restrict__ vals_in, uint32_t* __restrict__ keys_out,
                                                      uint32_t* __restrict__ vals_out, int n, int shift, int first,
                                                      const uint32_t* __restrict__ hist, int nblocks) {
    __shared__ uint32_t s_run[256];          // running offset per digit (global base + earlier rounds)
    __shared__ uint32_t s_cnt[SW][256];      // per-warp digit counts of the current round
    const int col = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* kin = keys_in + (size_t)col * n;
    const uint32_t* vin = vals_in + (size_t)col * n;
    uint32_t* kout = keys_out + (size_t)col * n;
    uint32_t* vout = vals_out + (size_t)col * n;
    s_run[tid] = hist[((size_t)col * 256 + tid) * nblocks + blockIdx.x];
    int base = blockIdx.x * CHUNK;
    for (int r = 0; r < ITEMS; r++) {
        for (int q = 0; q < SW; q++) s_cnt[q][tid] = 0;
        __syncthreads();
        int i = base + r * ST + tid;
        bool valid = i < n;
        uint32_t key = valid ? load_key(X, dp, col, kin, i, first) : 0xffffffffu;
        uint32_t val = valid ? (first ? (uint32_t)i : vin[i]) : 0u;
        uint32_t dg = (key >> shift) & 255u;
        // invalid lanes get a private pseudo-digit so that they never match valid ones
        unsigned peers = __match_any_sync(F16_FULL, valid ? dg : (256u + lane));
        int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) s_cnt[warp][dg] = __popc(peers);
        __syncthreads();
        // thread == digit: prefix over warps, advance running offset
        uint32_t run = s_run[tid], acc = run;
        for (int q = 0; q < SW; q++) { uint32_t cq = s_cnt[q][tid]; s_cnt[q][tid] = acc; acc += cq; }
        s_run[tid] = acc;
        __syncthreads();
        if (valid) {
            uint32_t pos = s_cnt[warp][dg] + rank;
            kout[pos] = key;
            vout[pos] = val;
        }
        __syncthreads();
    }
}

// sorted_idx_dev: [d][n] int32; ascending by X[:, col]; ties in arbitrary (stable) order.
extern "C" int f16_argsort_columns(const float* X_dev, int64_t n, int32_t d, int32_t* sorted_idx_dev, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!X_dev || !sorted_idx_dev || n < 1 || n > F16_MAX_ROWS - 1 || d < 1 || d > F16_MAX_D) {
        f16_set_error("f16_argsort_columns: bad arguments"); return F16_ERR_INVALID;
    }
    int dp = (d <= 8) ? 8 : 16;
    int nblocks = (int)((n + CHUNK - 1) / CHUNK);
    uint32_t *ka = nullptr, *va = nullptr, *kb = nullptr, *hist = nullptr;
    size_t cn = (size_t)d * (size_t)n;
    CUDA_TRY(cudaMallocAsync((void**)&ka, 4 * cn, st));
    CUDA_TRY(cudaMallocAsync((void**)&va, 4 * cn, st));
    CUDA_TRY(cudaMallocAsync((void**)&kb, 4 * cn, st));
    CUDA_TRY(cudaMallocAsync((void**)&hist, 4 * (size_t)d * 256 * nblocks, st));
    uint32_t* vb = (uint32_t*)sorted_idx_dev;
    dim3 grid(nblocks, d);
    for (int pass = 0; pass < 4; pass++) {
        const uint32_t* kin = (pass & 1) ? ka : kb;
        const uint32_t* vin = (pass & 1) ? va : vb;
        uint32_t* kout = (pass & 1) ? kb : ka;
        uint32_t* vout = (pass & 1) ? vb : va;
        int first = pass == 0;
        k_radix_hist<<<grid, ST, 0, st>>>(X_dev, dp, kin, (int)n, pass * 8, first, hist, nblocks);
        f16_count_launch(1);
        k_radix_scan<<<d, 1024, 0, st>>>(hist, nblocks);
        f16_count_launch(1);
        k_radix_scatter<<<grid, ST, 0, st>>>(X_dev, dp, kin, vin, kout, vout, (int)n, pass * 8, first, hist, nblocks);
        f16_count_launch(1);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaFreeAsync(ka, st));
    CUDA_TRY(cudaFreeAsync(va, st));
    CUDA_TRY(cudaFreeAsync(kb, st));
    CUDA_TRY(cudaFreeAsync(hist, st));
    return F16_OK;
}
