// Tensor-core candidate filter for the exact k-nearest-neighbour search: the mma.sync version
// (f16_knn strategy 5) and the pieces it shares with the tcgen05 version (strategy 3,
// f16_knn_umma.cu): column sums, candidate-list format, exact float64 selection.
//
// The float64 search (f16_knn.cu) is bound by the FP64 pipe: n * nq * d DFMAs.  For centred,
// moderately scaled data (StandardScaler / PCA outputs) almost all of that arithmetic only proves
// that a pair is NOT among the k nearest.  Here that proof is done on the tensor cores:
//
//   1. prep    every point is split into two float16 vectors (x = hi + lo + r, |r| <= 2^-22 |x|)
//              and its squared norm is rounded to float32;
//   2. filter  s~(q, j) = |q|^2 + |x_j|^2 - 2 (hi_q.hi_j + hi_q.lo_j + lo_q.hi_j) is evaluated with
//              mma.sync.m16n8k16 (float32 accumulation) for ALL pairs; its distance from the true
//              squared distance is bounded by E = eps (|q|^2 + |x_j|^2) + slack, so
//                  s~ - E  <=  d^2  <=  s~ + E.
//              Every thread keeps, per query row, the k smallest UPPER bounds it has seen; the
//              k-th of them (U) bounds the true k-th smallest distance from above, so a pair
//              whose LOWER bound exceeds U can never be among the k nearest (nor tie with the
//              k-th).  Every other pair is appended to the query's candidate list (a few dozen
//              entries out of n);
//   3. select  the candidates are evaluated with the same float64 direct sum of squared
//              differences, in the same coordinate order, as the float64 kernel, and the k best
//              by (distance, index) are written - bit-identical output.  A query whose list
//              overflowed (or whose data does not fit float16) is searched exhaustively.
//
// Error budget of the filter (u16 = 2^-11, u32 = 2^-24): split remainder 2 * 2^-22, dropped
// lo.lo 2^-22, 48 float32 accumulations <= 48 * 2^-22 (allowing truncation instead of rounding),
// all relative to sum |q_i x_i| <= (|q|^2 + |x|^2) / 2, times 2 for the -2 q.x term, plus three
// float32 roundings of the norms / final sum: < 1.3e-5 (|q|^2 + |x|^2).  eps = 6e-5 leaves a
// factor > 4; slack covers float16 underflow of tiny coordinates.  tests/test_ops_gpu.py measures
// the real error on the device (f16_knn_tc_probe) and checks it against eps / 4.
#include "f16_common.cuh"
#include <cuda_fp16.h>
#include <math.h>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st);
#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define TC_TILE 128       // reference points per shared-memory tile (double buffered)
#define TC_RS 24          // halfs per shared row: 48-byte stride => conflict-free fragment loads
#define TC_CAP 256        // candidate slots per query
#define TC_EPS 6.0e-5f
#define TC_SLACK 1.0e-6f
#define TC_MAXABS 60000.0
#ifndef TC_MT
#define TC_MT 2           // m16 query tiles per warp (4 measured no faster: fewer resident warps)
#endif

struct TcPerm { int c[F16_MAX_D]; };

// ------------------------------------------------------------------ 1. prep
// hi / lo: [n][16] halfs (zero padded), nrm: [n] float.  A coordinate that does not fit float16
// (or is not finite) raises *bad: the filter is skipped and every query is searched exhaustively
// in float64 - slow, but the result never depends on the filter being applicable.
// Euclidean distances are translation invariant: both point sets are shifted by the reference
// set's column means (colsum / n) first, which minimises the norms the filter's error band is
// relative to (uncentred raw columns would make the band wider than the neighbour distances).
__global__ void k_knn_tc_colsum(const double* __restrict__ A, int n, int d, double* __restrict__ colsum) {
    __shared__ double s_sum[F16_MAX_D];
    if (threadIdx.x < F16_MAX_D) s_sum[threadIdx.x] = 0.0;
    __syncthreads();
    const int c = threadIdx.x % F16_MAX_D;               // blockDim is a multiple of 16
    double acc = 0.0;
    if (c < d) {
        for (size_t i = (size_t)blockIdx.x * (blockDim.x / F16_MAX_D) + threadIdx.x / F16_MAX_D; i < (size_t)n;
             i += (size_t)gridDim.x * (blockDim.x / F16_MAX_D))
            acc += A[i * d + c];
        atomicAdd(&s_sum[c], acc);
    }
    __syncthreads();
    if (threadIdx.x < d) atomicAdd(&colsum[threadIdx.x], s_sum[threadIdx.x]);
}

__global__ void k_knn_tc_prep(const double* __restrict__ A, int n, int d, const double* __restrict__ colsum, double inv_n,
                              __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ nrm, int* __restrict__ bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    bool ok = true;
    __half h[16], l[16];
#pragma unroll
    for (int c = 0; c < 16; c++) {
        double v = (c < d) ? A[(size_t)i * d + c] - colsum[c] * inv_n : 0.0;
        if (!(fabs(v) <= TC_MAXABS)) ok = false;
        __half hv = __double2half(v);
        h[c] = hv;
        l[c] = __double2half(v - (double)__half2float(hv));
        s = fma(v, v, s);
    }
#pragma unroll
    for (int c = 0; c < 16; c++) {
        hi[(size_t)i * 16 + c] = ok ? h[c] : __float2half(0.f);
        lo[(size_t)i * 16 + c] = ok ? l[c] : __float2half(0.f);
    }
    // stored as the filter's accumulator seed: -|x|^2 (1 - eps) / 2
    nrm[i] = ok ? -0.5f * ((float)s * (1.0f - TC_EPS)) : 0.f;
    if (!ok) atomicExch(bad, 1);
}

// ------------------------------------------------------------------ 2. filter
__device__ __forceinline__ void tc_cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void tc_cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void tc_cp_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }

__device__ __forceinline__ void tc_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// PROBE: instead of filtering, record max |s~ - d^2| / (|q|^2 + |x|^2 + slack') over all pairs
// (d^2 recomputed in float64) - the measured error the tests compare with TC_EPS.
//
// A warp owns MT m16 tiles of queries (MT*16 rows) and sweeps all references in n8 chunks:
//   acc = -nr'/2                      (accumulator initialised from the reference norms)
//   acc += lo_q.hi_x + hi_q.lo_x + hi_q.hi_x          (3 HMMA per tile)
// so that -2 acc = nr' - 2 q.x and the filter test "lower bound <= U" is ONE compare per
// output: acc >= thr[row] with thr = (nq' - U) / 2.  Rows that pass take the slow path
// (candidate append, top-k of upper bounds, new threshold).
template <int K, int MT, bool PROBE>
__global__ void __launch_bounds__(128) k_knn_tc_filter2(const __half* __restrict__ Ah, const __half* __restrict__ Al,
                                                        const float* __restrict__ An, int n,
                                                        const __half* __restrict__ Qh, const __half* __restrict__ Ql,
                                                        const float* __restrict__ Qn, int nq,
                                                        uint32_t* __restrict__ cand, int* __restrict__ cand_cnt,
                                                        const double* __restrict__ A64, const double* __restrict__ Q64, int d,
                                                        float* __restrict__ probe_out, const double* __restrict__ colsum,
                                                        double inv_n) {
    constexpr int QPB = 64 * MT;          // queries per CTA: 4 warps x MT tiles x 16 rows
    constexpr int R = 2 * MT;             // query rows per thread
    __shared__ __align__(16) __half s_hi[2][TC_TILE * TC_RS];
    __shared__ __align__(16) __half s_lo[2][TC_TILE * TC_RS];
    __shared__ __align__(16) float s_nh[2][TC_TILE];
    __shared__ int s_cnt[QPB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int q_base = blockIdx.x * QPB + warp * (MT * 16);

    for (int i = tid; i < QPB; i += 128) s_cnt[i] = 0;
    if (!PROBE && cand_cnt[nq] != 0) return;      // data outside the float16 range: no filter (uniform exit)

    uint32_t ah[MT][4], al[MT][4];
    float nqp[R], nq2[R], thr[R], ub[R];
    float tk[R][K];
    int qrow[R];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = mt * 2 + h;
            const int q = q_base + mt * 16 + h * 8 + g;
            qrow[r] = q;
            const bool ok = q < nq;
            const uint32_t* ph = reinterpret_cast<const uint32_t*>(Qh + (size_t)(ok ? q : 0) * 16);
            const uint32_t* pl = reinterpret_cast<const uint32_t*>(Ql + (size_t)(ok ? q : 0) * 16);
            ah[mt][h] = ok ? ph[t] : 0u;          // row g (+8 for h = 1), k = 2t, 2t+1
            ah[mt][h + 2] = ok ? ph[t + 4] : 0u;  // k = 2t+8, 2t+9
            al[mt][h] = ok ? pl[t] : 0u;
            al[mt][h + 2] = ok ? pl[t + 4] : 0u;
            nq2[r] = ok ? -2.f * Qn[q] : 0.f;     // |q|^2 (1 - eps)
            nqp[r] = nq2[r] - TC_SLACK;
            ub[r] = INFINITY;
            thr[r] = ok ? -INFINITY : INFINITY;   // rows past nq never produce candidates
#pragma unroll
            for (int m = 0; m < K; m++) tk[r][m] = INFINITY;
        }
    }
    float perr = 0.f;

    auto stage = [&](int buf, int base) {
        const int cnt = min(TC_TILE, n - base);
        if (cnt == TC_TILE) {
            for (int i = tid; i < TC_TILE * 2; i += 128) {
                const int row = i >> 1, piece = i & 1;
                tc_cp_async16(&s_hi[buf][row * TC_RS + piece * 8], Ah + (size_t)(base + row) * 16 + piece * 8);
                tc_cp_async16(&s_lo[buf][row * TC_RS + piece * 8], Al + (size_t)(base + row) * 16 + piece * 8);
            }
            if (tid < TC_TILE / 4) tc_cp_async16(&s_nh[buf][tid * 4], An + base + tid * 4);
        } else {
            // last, partial tile: plain loads; padding rows get -inf (never above a finite threshold)
            for (int i = tid; i < TC_TILE * 16; i += 128) {
                const int row = i >> 4, c = i & 15;
                s_hi[buf][row * TC_RS + c] = (row < cnt) ? Ah[(size_t)(base + row) * 16 + c] : __float2half(0.f);
                s_lo[buf][row * TC_RS + c] = (row < cnt) ? Al[(size_t)(base + row) * 16 + c] : __float2half(0.f);
            }
            for (int i = tid; i < TC_TILE; i += 128) s_nh[buf][i] = (i < cnt) ? An[base + i] : -INFINITY;
        }
        tc_cp_commit();
    };

    stage(0, 0);
    int buf = 0;
    for (int base = 0; base < n; base += TC_TILE, buf ^= 1) {
        tc_cp_wait_all();
        __syncthreads();
        if (base + TC_TILE < n) stage(buf ^ 1, base + TC_TILE);
        // B fragments of chunk ch: reference ch*8 + g, k = 2t, 2t+1 (+8); norms of columns 2t, 2t+1
        const uint32_t* rh = reinterpret_cast<const uint32_t*>(s_hi[buf] + g * TC_RS) + t;
        const uint32_t* rl = reinterpret_cast<const uint32_t*>(s_lo[buf] + g * TC_RS) + t;
        const float2* pn = reinterpret_cast<const float2*>(s_nh[buf]) + t;
#pragma unroll 2
        for (int ch = 0; ch < TC_TILE / 8; ch++, rh += 8 * TC_RS / 2, rl += 8 * TC_RS / 2, pn += 4) {
            const uint32_t bh0 = rh[0], bh1 = rh[4], bl0 = rl[0], bl1 = rl[4];
            const float2 nh = *pn;
            float acc[MT][4];
            bool hit = false;
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                acc[mt][0] = nh.x; acc[mt][1] = nh.y; acc[mt][2] = nh.x; acc[mt][3] = nh.y;
                tc_mma(acc[mt], al[mt], bh0, bh1);      // small terms first
                tc_mma(acc[mt], ah[mt], bl0, bl1);
                tc_mma(acc[mt], ah[mt], bh0, bh1);
                hit = hit || (fmaxf(acc[mt][0], acc[mt][1]) >= thr[mt * 2]) || (fmaxf(acc[mt][2], acc[mt][3]) >= thr[mt * 2 + 1]);
            }
            if (PROBE) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int r = mt * 2 + (o >> 1);
                        const int j = base + ch * 8 + 2 * t + (o & 1);
                        if (qrow[r] < nq && j < n) {
                            double s = 0.0, na = 0.0, nb = 0.0;
                            for (int c = 0; c < d; c++) {
                                const double mu = colsum[c] * inv_n;
                                double a = Q64[(size_t)qrow[r] * d + c] - mu, b = A64[(size_t)j * d + c] - mu;
                                s = fma(a - b, a - b, s); na = fma(a, a, na); nb = fma(b, b, nb);
                            }
                            // plain estimate |q|^2 + |x|^2 - 2 q.x: undo the (1 - eps) scaling of the norms
                            const double nrp = -2.0 * (double)((o & 1) ? nh.y : nh.x);
                            double est = -2.0 * (double)acc[mt][o] - nrp + (nrp + (double)nq2[r]) / (1.0 - (double)TC_EPS);
                            perr = fmaxf(perr, (float)(fabs(est - s) / (na + nb + 1e-3)));
                        }
                    }
                }
                continue;
            }
            if (__any_sync(F16_FULL, hit)) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                    for (int o = 0; o < 4; o++) {
                        const int r = mt * 2 + (o >> 1);
                        const int j = base + ch * 8 + 2 * t + (o & 1);
                        if (acc[mt][o] >= thr[r] && j < n) {
                            const float l = -2.f * acc[mt][o];                   // nr' - 2 q.x
                            const float nrv = -2.f * ((o & 1) ? nh.y : nh.x);    // nr'
                            // upper bound of the true squared distance
                            const float u = (l + nqp[r]) + 2.2f * TC_EPS * (nrv + nq2[r]) + 4.f * TC_SLACK;
                            if (u < tk[r][K - 1]) {
                                tk[r][K - 1] = u;
#pragma unroll
                                for (int m = K - 1; m > 0; m--) {
                                    if (tk[r][m] < tk[r][m - 1]) { float x = tk[r][m]; tk[r][m] = tk[r][m - 1]; tk[r][m - 1] = x; }
                                }
                                ub[r] = fminf(ub[r], tk[r][K - 1]);
                                thr[r] = 0.5f * (nqp[r] - ub[r]);
                            }
                            const int ql = warp * (MT * 16) + (r >> 1) * 16 + (r & 1) * 8 + g;
                            const int pos = atomicAdd(&s_cnt[ql], 1);
                            if (pos < TC_CAP) cand[(size_t)qrow[r] * TC_CAP + pos] = (uint32_t)j;
                        }
                    }
                }
            }
        }
        // the four threads of a quad hold disjoint column subsets of the same rows: the smallest
        // of their k-th upper bounds is still an upper bound of the true k-th distance
#pragma unroll
        for (int r = 0; r < R; r++) {
            float m = ub[r];
            m = fminf(m, __shfl_xor_sync(F16_FULL, m, 1));
            m = fminf(m, __shfl_xor_sync(F16_FULL, m, 2));
            ub[r] = m;
            if (qrow[r] < nq) thr[r] = 0.5f * (nqp[r] - m);
        }
    }
    __syncthreads();
    if (PROBE) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) perr = fmaxf(perr, __shfl_xor_sync(F16_FULL, perr, off));
        if (lane == 0) atomicMax(reinterpret_cast<int*>(probe_out), __float_as_int(perr));   // perr >= 0
        return;
    }
    for (int i = tid; i < QPB; i += 128) {
        const int q = blockIdx.x * QPB + i;
        if (q < nq) cand_cnt[q] = s_cnt[i];
    }
}

// ------------------------------------------------------------------ 3. select
// One warp per query: exact float64 distances of the candidates (same arithmetic and coordinate
// order as k_knn), best K by (distance, index).
template <int K>
__global__ void __launch_bounds__(256) k_knn_tc_select(const double* __restrict__ A, int n, const double* __restrict__ Q, int nq,
                                                       int d, TcPerm perm, const uint32_t* __restrict__ cand,
                                                       const int* __restrict__ cand_cnt, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    double q[F16_MAX_D];
#pragma unroll
    for (int c = 0; c < F16_MAX_D; c++) q[c] = (c < d) ? Q[(size_t)qi * d + perm.c[c]] : 0.0;
    const bool bad = cand_cnt[nq] != 0;         // filter not applicable
    const int cnt = bad ? 0 : cand_cnt[qi];
    const bool all = bad || cnt > TC_CAP;       // overflow: exhaustive search
    const int m = all ? n : cnt;
    double bd[K]; int bi[K];
#pragma unroll
    for (int k = 0; k < K; k++) { bd[k] = INFINITY; bi[k] = 0x7fffffff; }
    for (int i = lane; i < m; i += 32) {
        const int j = all ? i : (int)cand[(size_t)qi * TC_CAP + i];
        const double* row = A + (size_t)j * d;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < F16_MAX_D; c++) {
            if (c < d) { double df = q[c] - row[perm.c[c]]; s = fma(df, df, s); }
        }
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
    // K rounds of warp arg-min over the lanes' sorted heads
#pragma unroll
    for (int k = 0; k < K; k++) {
        double hd = bd[0]; int hi = bi[0]; int hl = lane;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double od = __shfl_xor_sync(F16_FULL, hd, off);
            int oi = __shfl_xor_sync(F16_FULL, hi, off);
            int ol = __shfl_xor_sync(F16_FULL, hl, off);
            if (od < hd || (od == hd && oi < hi)) { hd = od; hi = oi; hl = ol; }
        }
        if (lane == 0) out[(size_t)qi * K + k] = (hi == 0x7fffffff) ? -1 : hi;
        if (lane == hl) {
#pragma unroll
            for (int x = 0; x < K - 1; x++) { bd[x] = bd[x + 1]; bi[x] = bi[x + 1]; }
            bd[K - 1] = INFINITY; bi[K - 1] = 0x7fffffff;
        }
    }
}

// ------------------------------------------------------------------ host
struct TcBuffers {
    __half *ah = nullptr, *al = nullptr, *qh = nullptr, *ql = nullptr;
    float *an = nullptr, *qn = nullptr;
    uint32_t* cand = nullptr;
    int* cnt = nullptr;
    double* colsum = nullptr;
};

// column sums of the reference set, then the float16 split of both sets about the column means
static void tc_prepare(TcBuffers& b, const double* A, int n, const double* Q, int nq, int d, bool same, cudaStream_t st) {
    k_knn_tc_colsum<<<296, 256, 0, st>>>(A, n, d, b.colsum);
    const double inv_n = 1.0 / (double)n;
    k_knn_tc_prep<<<(n + 127) / 128, 128, 0, st>>>(A, n, d, b.colsum, inv_n, b.ah, b.al, b.an, b.cnt + nq);
    if (!same) k_knn_tc_prep<<<(nq + 127) / 128, 128, 0, st>>>(Q, nq, d, b.colsum, inv_n, b.qh, b.ql, b.qn, b.cnt + nq);
}

static int tc_alloc(TcBuffers& b, int n, int nq, bool same, cudaStream_t st) {
    CUDA_TRY(f16_malloc_async((void**)&b.ah, sizeof(__half) * 16 * (size_t)n, st));
    CUDA_TRY(f16_malloc_async((void**)&b.al, sizeof(__half) * 16 * (size_t)n, st));
    CUDA_TRY(f16_malloc_async((void**)&b.an, sizeof(float) * ((size_t)n + 4), st));
    if (!same) {
        CUDA_TRY(f16_malloc_async((void**)&b.qh, sizeof(__half) * 16 * (size_t)nq, st));
        CUDA_TRY(f16_malloc_async((void**)&b.ql, sizeof(__half) * 16 * (size_t)nq, st));
        CUDA_TRY(f16_malloc_async((void**)&b.qn, sizeof(float) * ((size_t)nq + 4), st));
    }
    CUDA_TRY(f16_malloc_async((void**)&b.cand, sizeof(uint32_t) * TC_CAP * (size_t)nq, st));
    CUDA_TRY(f16_malloc_async((void**)&b.cnt, sizeof(int) * ((size_t)nq + 1), st));     // [nq] = 'bad data' flag
    CUDA_TRY(cudaMemsetAsync(b.cnt + nq, 0, sizeof(int), st));
    CUDA_TRY(f16_malloc_async((void**)&b.colsum, sizeof(double) * F16_MAX_D, st));
    CUDA_TRY(cudaMemsetAsync(b.colsum, 0, sizeof(double) * F16_MAX_D, st));
    return F16_OK;
}

static void tc_free(TcBuffers& b, cudaStream_t st) {
    void* p[] = {b.ah, b.al, b.an, b.qh, b.ql, b.qn, b.cand, b.cnt, b.colsum};
    for (void* x : p) if (x) cudaFreeAsync(x, st);
}

// Returns F16_OK when the search was done here, F16_ERR_INVALID when the caller should use the
// float64 kernel (unsupported k / d).
int f16_knn_tc_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                      cudaStream_t st) {
    if (k < 1 || k > 8 || d < 1 || d > F16_MAX_D) return F16_ERR_INVALID;
    const bool same = (A == Q) && (n == nq);
    TcBuffers b;
    int rc = tc_alloc(b, n, nq, same, st);
    if (rc != F16_OK) { tc_free(b, st); return rc; }
    tc_prepare(b, A, n, Q, nq, d, same, st);
    const __half* qh = same ? b.ah : b.qh;
    const __half* ql = same ? b.al : b.ql;
    const float* qn = same ? b.an : b.qn;
    TcPerm pm;
    for (int c = 0; c < F16_MAX_D; c++) pm.c[c] = (c < d) ? perm[c] : c;
    const int grid = (nq + 64 * TC_MT - 1) / (64 * TC_MT);
    const int sgrid = (nq + 7) / 8;
#define TC_LAUNCH(KK)                                                                                            \
    case KK:                                                                                                     \
        k_knn_tc_filter2<KK, (KK <= 4 ? TC_MT : 2), false><<<(KK <= 4 ? grid : (nq + 127) / 128), 128, 0, st>>>(  \
            b.ah, b.al, b.an, n, qh, ql, qn, nq, b.cand, b.cnt, nullptr, nullptr, d, nullptr, nullptr, 0.0);     \
        k_knn_tc_select<KK><<<sgrid, 256, 0, st>>>(A, n, Q, nq, d, pm, b.cand, b.cnt, out);                       \
        break;
    switch (k) { TC_LAUNCH(1) TC_LAUNCH(2) TC_LAUNCH(3) TC_LAUNCH(4) TC_LAUNCH(5) TC_LAUNCH(6) TC_LAUNCH(7) TC_LAUNCH(8) }
#undef TC_LAUNCH
    f16_count_launch(same ? 4 : 5);
    cudaError_t e = cudaGetLastError();
    tc_free(b, st);
    if (e != cudaSuccess) { f16_set_error("f16_knn (tensor filter): %s", cudaGetErrorString(e)); return F16_ERR_CUDA; }
    return F16_OK;
}

// ---- pieces shared with the tcgen05 variant of the filter (f16_knn_umma.cu)
int f16_knn_tc_cap() { return TC_CAP; }

void f16_knn_tc_colsum_launch(const double* A, int n, int d, double* colsum, cudaStream_t st) {
    k_knn_tc_colsum<<<296, 256, 0, st>>>(A, n, d, colsum);
}

// cand: [nq][TC_CAP] candidate indices, cnt: [nq + 1] counts (> TC_CAP = overflow), cnt[nq] = 'bad data'
int f16_knn_tc_select_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm,
                             const uint32_t* cand, const int* cnt, int32_t* out, cudaStream_t st) {
    TcPerm pm;
    for (int c = 0; c < F16_MAX_D; c++) pm.c[c] = (c < d) ? perm[c] : c;
    const int sgrid = (nq + 7) / 8;
#define TC_SELECT(KK) case KK: k_knn_tc_select<KK><<<sgrid, 256, 0, st>>>(A, n, Q, nq, d, pm, cand, cnt, out); break;
    switch (k) {
        TC_SELECT(1) TC_SELECT(2) TC_SELECT(3) TC_SELECT(4) TC_SELECT(5) TC_SELECT(6) TC_SELECT(7) TC_SELECT(8)
        default: return F16_ERR_INVALID;
    }
#undef TC_SELECT
    return F16_OK;
}

// Test hook: largest observed |s~ - d^2| / (|q|^2 + |x|^2 + 1e-3) of the tensor-core estimate over
// all (query, reference) pairs (float64 recomputation on the device; small inputs only).
extern "C" int f16_knn_tc_probe(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, float* err_host,
                                void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!A_dev || !Q_dev || !err_host || n < 1 || nq < 1 || d < 1 || d > F16_MAX_D || n > (1 << 22) || nq > (1 << 22)) {
        f16_set_error("f16_knn_tc_probe: bad arguments"); return F16_ERR_INVALID;
    }
    TcBuffers b;
    int rc = tc_alloc(b, (int)n, (int)nq, false, st);
    if (rc != F16_OK) { tc_free(b, st); return rc; }
    float* perr = nullptr;
    CUDA_TRY(f16_malloc_async((void**)&perr, sizeof(float), st));
    CUDA_TRY(cudaMemsetAsync(perr, 0, sizeof(float), st));
    tc_prepare(b, A_dev, (int)n, Q_dev, (int)nq, d, false, st);
    k_knn_tc_filter2<4, TC_MT, true><<<((int)nq + 64 * TC_MT - 1) / (64 * TC_MT), 128, 0, st>>>(
        b.ah, b.al, b.an, (int)n, b.qh, b.ql, b.qn, (int)nq, b.cand, b.cnt, A_dev, Q_dev, d, perr, b.colsum, 1.0 / (double)n);
    CUDA_TRY(cudaMemcpyAsync(err_host, perr, sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaFreeAsync(perr, st));
    tc_free(b, st);
    CUDA_TRY(cudaGetLastError());
    return F16_OK;
}
