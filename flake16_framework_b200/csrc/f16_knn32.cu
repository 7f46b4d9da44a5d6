// k-NN search with a float32 conservative filter (exact float64 result).
//
// Same contract as k_knn in f16_knn.cu - the float64 direct distance decides, ties by index - but
// the pair filter  acc = x.y - (1-eps)/2 (|x|^2 + |y|^2)  is accumulated in FLOAT32 (FFMA runs at
// twice the FP64 rate on sm_100a and halves the register footprint), with the norms kept in
// float64.  eps = 4e-6 covers the float32 rounding of the inputs and of the 16-term dot product
// (|error| <= 8e-7 (|x|^2+|y|^2)), so a true neighbour always survives the filter; survivors
// (about k ln n per query) are recomputed exactly from the float64 rows in global memory.
// The filter is only *useful* when eps (|x|^2+|y|^2) is small against nearest-neighbour distances,
// i.e. for standardised / PCA-rotated data - the caller selects it per dataset (mode 2 of f16_knn);
// on badly scaled data it stays correct but degenerates to the exact recomputation.
#include "f16_common.cuh"
#include <math.h>

#define K32T 128          // threads per block
#define K32Q 2            // queries per thread
#define K32TILE 128       // reference rows per tile
#define KNN32_SCALE (-0.5 * (1.0 - 4.0e-6))

struct Knn32Perm { int c[F16_MAX_D]; };

template <int D>
__global__ void k_knn32_prep(const double* __restrict__ A, int n, Knn32Perm perm, float* __restrict__ c32,
                             double* __restrict__ an) {
    constexpr int DP4 = (D + 3) / 4 * 4;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < DP4; c++) {
        double v = (c < D) ? A[(size_t)i * D + perm.c[c]] : 0.0;
        s = fma(v, v, s);
        c32[(size_t)i * DP4 + c] = (float)v;
    }
    an[i] = KNN32_SCALE * s;
}

__device__ __forceinline__ void k32_cp16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void k32_cp8(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}

template <int D, int K>
__global__ void __launch_bounds__(K32T) k_knn32(const double* __restrict__ A, int n, const double* __restrict__ Q, int nq,
                                                int32_t* __restrict__ out, Knn32Perm perm, const float* __restrict__ c32,
                                                const double* __restrict__ an) {
    constexpr int DP4 = (D + 3) / 4 * 4, V4 = DP4 / 4;
    __shared__ __align__(16) float s_c[2][K32TILE * DP4];
    __shared__ __align__(16) double s_n[2][K32TILE];
    const int tid = threadIdx.x;
    float q[K32Q][DP4];
    double qn[K32Q], bd[K32Q][K], thr[K32Q];
    int bi[K32Q][K], qi[K32Q];
#pragma unroll
    for (int u = 0; u < K32Q; u++) {
        qi[u] = blockIdx.x * (K32T * K32Q) + u * K32T + tid;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < DP4; c++) {
            double v = (c < D && qi[u] < nq) ? Q[(size_t)qi[u] * D + perm.c[c]] : 0.0;
            s = fma(v, v, s);
            q[u][c] = (float)v;
        }
        qn[u] = KNN32_SCALE * s;
#pragma unroll
        for (int m = 0; m < K; m++) { bd[u][m] = INFINITY; bi[u][m] = -1; }
        thr[u] = -INFINITY;
    }
    auto stage = [&](int b, int base) {
        const int cnt = min(K32TILE, n - base);
        const float* g = c32 + (size_t)base * DP4;
        for (int i = tid; i < cnt * V4; i += K32T) k32_cp16(&s_c[b][4 * i], g + 4 * (size_t)i);
        for (int i = tid; i < cnt; i += K32T) k32_cp8(&s_n[b][i], an + base + i);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    stage(0, 0);
    int buf = 0;
    for (int base = 0; base < n; base += K32TILE, buf ^= 1) {
        const int cnt = min(K32TILE, n - base);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        if (base + K32TILE < n) stage(buf ^ 1, base + K32TILE);
        const float4* t4 = reinterpret_cast<const float4*>(s_c[buf]);
        for (int j = 0; j < cnt; j++) {
            float a0[K32Q], a1[K32Q];
#pragma unroll
            for (int u = 0; u < K32Q; u++) { a0[u] = 0.f; a1[u] = 0.f; }
#pragma unroll
            for (int p = 0; p < V4; p++) {
                const float4 v = t4[j * V4 + p];
#pragma unroll
                for (int u = 0; u < K32Q; u++) {
                    a0[u] = fmaf(q[u][4 * p], v.x, a0[u]); a1[u] = fmaf(q[u][4 * p + 1], v.y, a1[u]);
                    a0[u] = fmaf(q[u][4 * p + 2], v.z, a0[u]); a1[u] = fmaf(q[u][4 * p + 3], v.w, a1[u]);
                }
            }
            const double nj = s_n[buf][j];
#pragma unroll
            for (int u = 0; u < K32Q; u++) {
                const double acc = (double)(a0[u] + a1[u]) + (qn[u] + nj);
                if (acc > thr[u] && qi[u] < nq) {
                    // survivor: the exact float64 direct distance decides
                    const double* xr = Q + (size_t)qi[u] * D;
                    const double* yr = A + (size_t)(base + j) * D;
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < D; c++) {
                        double df = xr[perm.c[c]] - yr[perm.c[c]];
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
#pragma unroll
    for (int u = 0; u < K32Q; u++) {
        if (qi[u] < nq) {
#pragma unroll
            for (int m = 0; m < K; m++) out[(size_t)qi[u] * K + m] = bi[u][m];
        }
    }
}

template <int D>
static int launch32(const double* A, int n, const double* Q, int nq, int k, int32_t* out, const Knn32Perm& pm,
                    float* c32, double* an, cudaStream_t st) {
    k_knn32_prep<D><<<(n + 255) / 256, 256, 0, st>>>(A, n, pm, c32, an);
    int grid = (nq + K32T * K32Q - 1) / (K32T * K32Q);
    switch (k) {
        case 2: k_knn32<D, 2><<<grid, K32T, 0, st>>>(A, n, Q, nq, out, pm, c32, an); break;
        case 4: k_knn32<D, 4><<<grid, K32T, 0, st>>>(A, n, Q, nq, out, pm, c32, an); break;
        case 6: k_knn32<D, 6><<<grid, K32T, 0, st>>>(A, n, Q, nq, out, pm, c32, an); break;
        default: return F16_ERR_INVALID;
    }
    f16_count_launch(2);
    return F16_OK;
}

// Returns F16_ERR_INVALID when (d, k) has no float32-filter instantiation (caller falls back to
// the float64 filter).  scratch: c32 float [n][16], an double [n].
int f16_knn32_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                     float* c32, double* an, cudaStream_t st) {
    Knn32Perm pm;
    for (int c = 0; c < F16_MAX_D; c++) pm.c[c] = perm[c];
    switch (d) {
        case 7: return launch32<7>(A, n, Q, nq, k, out, pm, c32, an, st);
        case 16: return launch32<16>(A, n, Q, nq, k, out, pm, c32, an, st);
        default: return F16_ERR_INVALID;
    }
}
