// tcgen05 / TMEM variant of the k-NN candidate filter (see f16_knn_tc.cu for the method and its
// error budget; the candidate lists, the exact float64 selection and the result are the same).
//
// One CTA owns NB = 2 blocks of 128 queries (UMMA M = 128) and sweeps the references in tiles of
// 128 (UMMA N); per block and tile
//   D[128 x 128] (TMEM, float32)  =  lo_q.hi_x + hi_q.lo_x + [2048 1 2^-11 0..].[s0 s1 s2 0..] + hi_q.hi_x
// four tcgen05.mma.kind::f16 (K = 16), issued by one thread.  The third product adds the
// reference's accumulator seed nh = -|x|^2 (1 - eps) / 2 (split in three float16), so that a TMEM
// lane (= query row) holds  q.x - nr'/2  and the filter test is one compare per element against
// the row's threshold (nq' - U) / 2.
// Warp roles: warp 0 = bulk-copy producer (cp.async.bulk + mbarrier complete_tx), warp 1 = TMEM
// owner and MMA issuer, warps 2..9 = epilogue, four per query block (tcgen05.ld 32 lanes x 32
// columns, one query per thread: the running top-k of upper bounds, the threshold and the
// candidate counter are plain registers; no atomics).  Two TMEM accumulators per block (NB x 2 x
// 128 = 512 columns) let the MMAs of tile t+1 overlap the epilogue of tile t; a 4-stage
// shared-memory ring feeds the MMAs, and every reference tile is used by both query blocks.
// Operands are stored by the prep kernel in the canonical K-major no-swizzle layout of the
// tensor core ("core matrices" of 8 rows x 16 bytes): per 8 points 256 bytes =
// [k 0..7 of the 8 points][k 8..15 of the 8 points], so a tile is one contiguous 4 KB block,
// LBO (K direction) = 128 B, SBO (row-group direction) = 256 B.
#include "f16_common.cuh"
#include <cuda_fp16.h>
#include <math.h>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st);
int f16_knn_tc_cap();
void f16_knn_tc_colsum_launch(const double* A, int n, int d, double* colsum, cudaStream_t st);
int f16_knn_tc_select_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm,
                             const uint32_t* cand, const int* cnt, int32_t* out, cudaStream_t st);
#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return F16_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)

#define UM_M 128
#define UM_N 128
#ifndef UM_STAGES
#define UM_STAGES 4
#endif
#define UM_TILE_BYTES 4096                 // 128 points x 16 halfs
#define UM_CAP 256                         // == TC_CAP (checked at launch)
#define UM_EPS 6.0e-5f                     // == TC_EPS
#define UM_SLACK 1.0e-6f
#define UM_MAXABS 60000.0
#ifndef UM_NB
#define UM_NB 2                            // query blocks of 128 rows per CTA
#endif

// halfs offset of (point i, coordinate k) in the core-matrix layout
__host__ __device__ __forceinline__ size_t um_off(size_t i, int k) {
    return (i >> 3) * 128 + (size_t)(k >> 3) * 64 + (i & 7) * 8 + (k & 7);
}

// ------------------------------------------------------------------ prep
// n_pad = n rounded up to 128; padding points: zero vectors, seed -inf (can never pass the filter)
__global__ void k_knn_umma_prep(const double* __restrict__ A, int n, int n_pad, int d, const double* __restrict__ colsum,
                                double inv_n, __half* __restrict__ hi, __half* __restrict__ lo, __half* __restrict__ h4,
                                float* __restrict__ nh, int* __restrict__ bad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    double s = 0.0;
    bool ok = true;
    __half h[16], l[16];
#pragma unroll
    for (int c = 0; c < 16; c++) {
        double v = (c < d && i < n) ? A[(size_t)i * d + c] - colsum[c] * inv_n : 0.0;
        if (!(fabs(v) <= UM_MAXABS)) ok = false;
        __half hv = __double2half(v);
        h[c] = hv;
        l[c] = __double2half(v - (double)__half2float(hv));
        s = fma(v, v, s);
    }
    // accumulator seed -|x|^2 (1 - eps) / 2, fed to the tensor core as 2048 s0 + s1 + 2^-11 s2 with
    // three float16 (each step removes 11 bits of the remainder, so neither large norms nor
    // float16 underflow of tiny ones leave an absolute error); norms beyond 2048 * 65000 do not fit
    float seed = (i < n) ? (ok ? -0.5f * ((float)s * (1.0f - UM_EPS)) : 0.f) : -INFINITY;
    if (i < n && !(fabsf(seed) <= 2048.f * 65000.f)) { ok = false; seed = 0.f; }
    __half s0, s1, s2;
    if (i < n) {
        s0 = __float2half(seed * (1.0f / 2048.f));
        const float r1 = seed - 2048.f * __half2float(s0);
        s1 = __float2half(r1);
        const float r2 = r1 - __half2float(s1);
        s2 = __float2half(r2 * 2048.f);
    } else {
        s0 = __float2half(-INFINITY); s1 = __float2half(0.f); s2 = __float2half(0.f);
    }
#pragma unroll
    for (int c = 0; c < 16; c++) {
        hi[um_off(i, c)] = ok ? h[c] : __float2half(0.f);
        lo[um_off(i, c)] = ok ? l[c] : __float2half(0.f);
        h4[um_off(i, c)] = (c == 0) ? s0 : (c == 1) ? s1 : (c == 2) ? s2 : __float2half(0.f);
    }
    nh[i] = seed;
    if (!ok) atomicExch(bad, 1);
}

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t um_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void um_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void um_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void um_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void um_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void um_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major, no swizzle, LBO = 128 B, SBO = 256 B, sm_100 version bit
__device__ __forceinline__ uint64_t um_desc(uint32_t saddr) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(128u >> 4) << 16;
    d |= (uint64_t)(256u >> 4) << 32;
    d |= 1ull << 46;
    return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = f16, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t um_idesc() {
    return (1u << 4) | ((uint32_t)(UM_N >> 3) << 17) | ((uint32_t)(UM_M >> 4) << 24);
}
__device__ __forceinline__ void um_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void um_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void um_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void um_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// ------------------------------------------------------------------ filter
// NB = query blocks (of 128 rows) per CTA.  Every CTA streams the whole reference set through
// its shared-memory ring, so the L2 -> SM traffic is (nq / (128 NB)) x 96 B x n: two blocks per
// CTA halve it and give the SM eight epilogue warps (two per scheduler).
template <int NB>
struct UmSmem {
    __half a_hi[NB][UM_M * 16], a_lo[NB][UM_M * 16], a_4[UM_M * 16];      // 4 KB each
    __half b[UM_STAGES][3][UM_N * 16];                                     // stages x {hi, lo, h4} x 4 KB
    float scratch[NB * 128][9];                                           // rare-path rows (odd stride: no bank conflicts)
    unsigned long long full[UM_STAGES], empty[UM_STAGES], tfull[NB][2], tempty[NB][2], afull;
    uint32_t tmem_base;
};

// tcgen05.ld of 32 lanes x 32 columns WITHOUT the wait: the registers are valid only after
// um_ld_wait(); nothing may read them in between.
__device__ __forceinline__ void um_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// The registers are in/out operands of the wait, so no use of them can be scheduled above it.
__device__ __forceinline__ void um_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}

template <int K, int NB, bool PROBE>
__global__ void __launch_bounds__(64 + 128 * NB) k_knn_umma_filter(
    const __half* __restrict__ Ah, const __half* __restrict__ Al, const __half* __restrict__ A4, const float* __restrict__ Anh,
    int n, int n_pad, const __half* __restrict__ Qh, const __half* __restrict__ Ql, const float* __restrict__ Qnh, int nq,
    uint32_t* __restrict__ cand, int* __restrict__ cand_cnt, const double* __restrict__ A64, const double* __restrict__ Q64,
    int d, float* __restrict__ probe_out, const double* __restrict__ colsum, double inv_n) {
    constexpr int THREADS = 64 + 128 * NB;
    constexpr uint32_t TCOLS = NB * 2 * UM_N;            // 256 or 512 TMEM columns
    extern __shared__ __align__(128) unsigned char um_raw[];
    UmSmem<NB>& S = *reinterpret_cast<UmSmem<NB>*>(um_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_tiles = n_pad / UM_N;

    if (!PROBE && cand_cnt[nq] != 0) return;      // data outside the float16 range: no filter (uniform exit)

    // constant A operand [2048 1 2^-11 0 ... 0] in the core-matrix layout (generic-proxy writes)
    for (int i = tid; i < UM_M * 16; i += THREADS) {
        const int k = ((i >> 6) & 1) * 8 + (i & 7);
        S.a_4[i] = __float2half(k == 0 ? 2048.f : k == 1 ? 1.f : k == 2 ? (1.f / 2048.f) : 0.f);
    }
    if (tid == 0) {
        for (int s = 0; s < UM_STAGES; s++) { um_mbar_init(um_smem(&S.full[s]), 1); um_mbar_init(um_smem(&S.empty[s]), 1); }
        for (int x = 0; x < NB; x++)
            for (int b = 0; b < 2; b++) { um_mbar_init(um_smem(&S.tfull[x][b]), 1); um_mbar_init(um_smem(&S.tempty[x][b]), 4); }
        um_mbar_init(um_smem(&S.afull), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // a_4 must be visible to the tensor core
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem(&S.tmem_base)), "r"(TCOLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    um_fence_before();
    __syncthreads();
    um_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ===== producer: the CTA's queries once, then the reference tiles through the ring
        // (the whole warp walks the loop so that it reaches the final barrier converged; lane 0 issues)
        if (lane == 0) {
            um_mbar_expect_tx(um_smem(&S.afull), NB * 2 * UM_TILE_BYTES);
            for (int x = 0; x < NB; x++) {
                const size_t qoff = ((size_t)blockIdx.x * NB + x) * UM_M * 16;
                um_bulk_g2s(um_smem(S.a_hi[x]), Qh + qoff, UM_TILE_BYTES, um_smem(&S.afull));
                um_bulk_g2s(um_smem(S.a_lo[x]), Ql + qoff, UM_TILE_BYTES, um_smem(&S.afull));
            }
        }
        __syncwarp();
        for (int t = 0; t < n_tiles; t++) {
            const int s = t % UM_STAGES;
            const uint32_t ph = (uint32_t)(t / UM_STAGES) & 1u;
            um_mbar_wait(um_smem(&S.empty[s]), ph ^ 1u);
            if (lane == 0) {
                const uint32_t bar = um_smem(&S.full[s]);
                const size_t off = (size_t)t * UM_N * 16;
                um_mbar_expect_tx(bar, 3 * UM_TILE_BYTES);
                um_bulk_g2s(um_smem(S.b[s][0]), Ah + off, UM_TILE_BYTES, bar);
                um_bulk_g2s(um_smem(S.b[s][1]), Al + off, UM_TILE_BYTES, bar);
                um_bulk_g2s(um_smem(S.b[s][2]), A4 + off, UM_TILE_BYTES, bar);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===== MMA issuer (lane 0 issues; the warp stays converged)
        const uint32_t idesc = um_idesc();
        const uint64_t da_4 = um_desc(um_smem(S.a_4));
        um_mbar_wait(um_smem(&S.afull), 0);
        for (int t = 0; t < n_tiles; t++) {
            const int s = t % UM_STAGES, b = t & 1;
            const uint32_t ph = (uint32_t)(t / UM_STAGES) & 1u, bph = (uint32_t)(t >> 1) & 1u;
            um_mbar_wait(um_smem(&S.full[s]), ph);
            const uint64_t db_hi = um_desc(um_smem(S.b[s][0])), db_lo = um_desc(um_smem(S.b[s][1])), db_4 = um_desc(um_smem(S.b[s][2]));
#pragma unroll
            for (int x = 0; x < NB; x++) {
                um_mbar_wait(um_smem(&S.tempty[x][b]), bph ^ 1u);
                um_fence_after();
                if (lane == 0) {
                    const uint32_t dt = tmem + (uint32_t)(x * 2 + b) * UM_N;
                    const uint64_t da_hi = um_desc(um_smem(S.a_hi[x])), da_lo = um_desc(um_smem(S.a_lo[x]));
                    um_mma(dt, da_lo, db_hi, idesc, 0u);      // small terms first
                    um_mma(dt, da_hi, db_lo, idesc, 1u);
                    um_mma(dt, da_4, db_4, idesc, 1u);
                    um_mma(dt, da_hi, db_hi, idesc, 1u);
                    um_commit(um_smem(&S.tfull[x][b]));       // accumulator ready for block x's epilogue
                }
                __syncwarp();
            }
            if (lane == 0) um_commit(um_smem(&S.empty[s]));   // smem slot free once these MMAs retire
            __syncwarp();
        }
    } else {
        // ===== epilogue: block x = (warp - 2) / 4, TMEM lane quarter warp % 4, one query row per thread
        const int x = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int q = (blockIdx.x * NB + x) * UM_M + row;
        const bool qok = q < nq;
        const float nq2 = qok ? -2.f * Qnh[q] : 0.f;       // |q|^2 (1 - eps)
        const float nqp = nq2 - UM_SLACK;
        float thr = qok ? -INFINITY : INFINITY;
        float ub = INFINITY;
        float tk[K];
#pragma unroll
        for (int m = 0; m < K; m++) tk[m] = INFINITY;
        int cnt = 0;
        float perr = 0.f;

        // one 32-column chunk: max of the chunk against the threshold, the rare hits one by one
        auto process = [&](const uint32_t (&r)[32], int j0) {
            if (PROBE) {
                for (int i = 0; i < 32; i++) {
                    const int j = j0 + i;
                    if (qok && j < n) {
                        double s = 0.0, na = 0.0, nb = 0.0;
                        for (int c = 0; c < d; c++) {
                            const double mu = colsum[c] * inv_n;
                            double a = Q64[(size_t)q * d + c] - mu, bb = A64[(size_t)j * d + c] - mu;
                            s = fma(a - bb, a - bb, s); na = fma(a, a, na); nb = fma(bb, bb, nb);
                        }
                        const double nrp = -2.0 * (double)Anh[j];
                        double est = -2.0 * (double)__uint_as_float(r[i]) - nrp + (nrp + (double)nq2) / (1.0 - (double)UM_EPS);
                        perr = fmaxf(perr, (float)(fabs(est - s) / (na + nb + 1e-3)));
                    }
                }
                return;
            }
            // maxima of the four groups of 8 columns (independent chains), then of the chunk
            float g[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float a0 = fmaxf(fmaxf(__uint_as_float(r[8 * u]), __uint_as_float(r[8 * u + 1])), __uint_as_float(r[8 * u + 2]));
                const float a1 = fmaxf(fmaxf(__uint_as_float(r[8 * u + 3]), __uint_as_float(r[8 * u + 4])), __uint_as_float(r[8 * u + 5]));
                g[u] = fmaxf(fmaxf(a0, a1), fmaxf(__uint_as_float(r[8 * u + 6]), __uint_as_float(r[8 * u + 7])));
            }
            if (fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])) >= thr) {
                // rare path, kept SMALL (its fully unrolled form overflowed the instruction cache and
                // dominated the kernel): the 8 values of a group that holds a hit are parked in the
                // thread's private shared-memory row and walked by a rolled loop
                float* mine = S.scratch[(warp - 2) * 32 + lane];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (g[u] >= thr) {
#pragma unroll
                        for (int i = 0; i < 8; i++) mine[i] = __uint_as_float(r[8 * u + i]);
#pragma unroll 1
                        for (int i = 0; i < 8; i++) {
                            const float v = mine[i];
                            const int j = j0 + 8 * u + i;
                            if (v >= thr && j < n) {
                                const float l = -2.f * v;                    // nr' - 2 q.x
                                const float nrv = -2.f * Anh[j];             // nr'
                                const float uu = (l + nqp) + 2.2f * UM_EPS * (nrv + nq2) + 4.f * UM_SLACK;
                                if (uu < tk[K - 1]) {
                                    tk[K - 1] = uu;
#pragma unroll
                                    for (int m = K - 1; m > 0; m--) {
                                        if (tk[m] < tk[m - 1]) { float y = tk[m]; tk[m] = tk[m - 1]; tk[m - 1] = y; }
                                    }
                                    ub = fminf(ub, tk[K - 1]);
                                    thr = 0.5f * (nqp - ub);
                                }
                                if (cnt < UM_CAP) cand[(size_t)q * UM_CAP + cnt] = (uint32_t)j;
                                cnt++;
                            }
                        }
                    }
                }
            }
            __syncwarp();           // the tcgen05.ld / wait that follow are warp-collective
        };

        for (int t = 0; t < n_tiles; t++) {
            const int b = t & 1;
            const uint32_t bph = (uint32_t)(t >> 1) & 1u;
            um_mbar_wait(um_smem(&S.tfull[x][b]), bph);
            um_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(x * 2 + b) * UM_N;
            const int j0 = t * UM_N;
            // two register sets: the load of the next chunk is in flight while this one is scanned
            uint32_t ra[32], rb[32];
            um_ld32_issue(taddr, ra);
            um_ld_wait(ra);
            um_ld32_issue(taddr + 32, rb);
            process(ra, j0);
            um_ld_wait(rb);
            um_ld32_issue(taddr + 64, ra);
            process(rb, j0 + 32);
            um_ld_wait(ra);
            um_ld32_issue(taddr + 96, rb);
            process(ra, j0 + 64);
            um_ld_wait(rb);
            process(rb, j0 + 96);
            um_fence_before();
            __syncwarp();
            if (lane == 0) um_mbar_arrive(um_smem(&S.tempty[x][b]));
        }
        if (PROBE) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) perr = fmaxf(perr, __shfl_xor_sync(F16_FULL, perr, off));
            if (lane == 0) atomicMax(reinterpret_cast<int*>(probe_out), __float_as_int(perr));
        } else if (qok) {
            cand_cnt[q] = cnt;
        }
    }
    um_fence_before();
    __syncthreads();
    if (warp == 1) {
        um_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
    }
}

// ------------------------------------------------------------------ host
struct UmBuffers {
    __half *ah = nullptr, *al = nullptr, *a4 = nullptr, *qh = nullptr, *ql = nullptr, *q4 = nullptr;
    float *anh = nullptr, *qnh = nullptr;
    uint32_t* cand = nullptr;
    int* cnt = nullptr;
    double* colsum = nullptr;
};

static void um_free(UmBuffers& b, cudaStream_t st) {
    void* p[] = {b.ah, b.al, b.a4, b.qh, b.ql, b.q4, b.anh, b.qnh, b.cand, b.cnt, b.colsum};
    for (void* x : p) if (x) cudaFreeAsync(x, st);
}

static int um_alloc(UmBuffers& b, int n_pad, int nq, int nq_pad, bool same, cudaStream_t st) {
    const size_t hb = sizeof(__half) * 16 * (size_t)n_pad;
    CUDA_TRY(f16_malloc_async((void**)&b.ah, hb, st));
    CUDA_TRY(f16_malloc_async((void**)&b.al, hb, st));
    CUDA_TRY(f16_malloc_async((void**)&b.a4, hb, st));
    CUDA_TRY(f16_malloc_async((void**)&b.anh, sizeof(float) * (size_t)n_pad, st));
    if (!same) {
        const size_t qb = sizeof(__half) * 16 * (size_t)nq_pad;
        CUDA_TRY(f16_malloc_async((void**)&b.qh, qb, st));
        CUDA_TRY(f16_malloc_async((void**)&b.ql, qb, st));
        CUDA_TRY(f16_malloc_async((void**)&b.q4, qb, st));
        CUDA_TRY(f16_malloc_async((void**)&b.qnh, sizeof(float) * (size_t)nq_pad, st));
    }
    CUDA_TRY(f16_malloc_async((void**)&b.cand, sizeof(uint32_t) * UM_CAP * (size_t)nq, st));
    CUDA_TRY(f16_malloc_async((void**)&b.cnt, sizeof(int) * ((size_t)nq + 1), st));
    CUDA_TRY(cudaMemsetAsync(b.cnt + nq, 0, sizeof(int), st));
    CUDA_TRY(f16_malloc_async((void**)&b.colsum, sizeof(double) * F16_MAX_D, st));
    CUDA_TRY(cudaMemsetAsync(b.colsum, 0, sizeof(double) * F16_MAX_D, st));
    return F16_OK;
}

template <int K, bool PROBE>
static cudaError_t um_launch_filter(const UmBuffers& b, int n, int n_pad, int nq, int nq_pad, bool same, const double* A64,
                                    const double* Q64, int d, float* perr, cudaStream_t st) {
    constexpr int NB = UM_NB;
    const size_t smem = sizeof(UmSmem<NB>) + 128;
    cudaError_t e = cudaFuncSetAttribute(k_knn_umma_filter<K, NB, PROBE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_knn_umma_filter<K, NB, PROBE><<<nq_pad / (UM_M * NB), 64 + 128 * NB, smem, st>>>(
        b.ah, b.al, b.a4, b.anh, n, n_pad, same ? b.ah : b.qh, same ? b.al : b.ql, same ? b.anh : b.qnh, nq, b.cand, b.cnt,
        A64, Q64, d, perr, b.colsum, 1.0 / (double)n);
    return cudaGetLastError();
}

static int um_run(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out, float* probe_host,
                  cudaStream_t st) {
    if (k < 1 || k > 8 || d < 1 || d > F16_MAX_D || f16_knn_tc_cap() != UM_CAP) return F16_ERR_INVALID;
    const bool same = (A == Q) && (n == nq) && !probe_host;
    const int n_pad = (n + UM_N * UM_NB - 1) / (UM_N * UM_NB) * (UM_N * UM_NB);     // also a multiple of the query block
    const int nq_pad = (nq + UM_M * UM_NB - 1) / (UM_M * UM_NB) * (UM_M * UM_NB);
    UmBuffers b;
    int rc = um_alloc(b, n_pad, nq, nq_pad, same, st);
    if (rc != F16_OK) { um_free(b, st); return rc; }
    f16_knn_tc_colsum_launch(A, n, d, b.colsum, st);
    const double inv_n = 1.0 / (double)n;
    k_knn_umma_prep<<<(n_pad + 127) / 128, 128, 0, st>>>(A, n, n_pad, d, b.colsum, inv_n, b.ah, b.al, b.a4, b.anh, b.cnt + nq);
    if (!same) k_knn_umma_prep<<<(nq_pad + 127) / 128, 128, 0, st>>>(Q, nq, nq_pad, d, b.colsum, inv_n, b.qh, b.ql, b.q4, b.qnh, b.cnt + nq);
    cudaError_t e = cudaSuccess;
    if (probe_host) {
        float* perr = nullptr;
        CUDA_TRY(f16_malloc_async((void**)&perr, sizeof(float), st));
        CUDA_TRY(cudaMemsetAsync(perr, 0, sizeof(float), st));
        e = um_launch_filter<4, true>(b, n, n_pad, nq, nq_pad, same, A, Q, d, perr, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(probe_host, perr, sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        cudaFreeAsync(perr, st);
    } else {
#define UM_LAUNCH(KK) case KK: e = um_launch_filter<KK, false>(b, n, n_pad, nq, nq_pad, same, nullptr, nullptr, d, nullptr, st); break;
        switch (k) { UM_LAUNCH(1) UM_LAUNCH(2) UM_LAUNCH(3) UM_LAUNCH(4) UM_LAUNCH(5) UM_LAUNCH(6) UM_LAUNCH(7) UM_LAUNCH(8) }
#undef UM_LAUNCH
        if (e == cudaSuccess) {
            rc = f16_knn_tc_select_launch(A, n, Q, nq, d, k, perm, b.cand, b.cnt, out, st);
            if (rc == F16_OK) e = cudaGetLastError();
        }
        f16_count_launch(same ? 4 : 5);
    }
    um_free(b, st);
    if (e != cudaSuccess) { f16_set_error("f16_knn (tcgen05 filter): %s", cudaGetErrorString(e)); return F16_ERR_CUDA; }
    return rc;
}

int f16_knn_umma_launch(const double* A, int n, const double* Q, int nq, int d, int k, const int* perm, int32_t* out,
                        cudaStream_t st) {
    return um_run(A, n, Q, nq, d, k, perm, out, nullptr, st);
}

// Test hook, like f16_knn_tc_probe but for the tcgen05 filter.
extern "C" int f16_knn_umma_probe(const double* A_dev, int64_t n, const double* Q_dev, int64_t nq, int32_t d, float* err_host,
                                  void* stream) {
    if (!A_dev || !Q_dev || !err_host || n < 1 || nq < 1 || d < 1 || d > F16_MAX_D || n > (1 << 22) || nq > (1 << 22)) {
        f16_set_error("f16_knn_umma_probe: bad arguments"); return F16_ERR_INVALID;
    }
    int perm[F16_MAX_D];
    for (int c = 0; c < F16_MAX_D; c++) perm[c] = c;
    return um_run(A_dev, (int)n, Q_dev, (int)nq, d, 4, perm, nullptr, err_host, (cudaStream_t)stream);
}
