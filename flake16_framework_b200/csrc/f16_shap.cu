// Path-dependent TreeSHAP over the fitted forests.
//   reference: get_shap / write_shap, experiment.py:504-530 -
//              shap.TreeExplainer(model).shap_values(features)[0]   (shap 0.40.0, third party)
//   algorithm: Lundberg, Erion, Lee 2018, Algorithm 2 (the game v(S) = E[f(x) | x_S] with the tree's
//              covers as the conditional distribution); restated and pinned in oracle/treeshap_np.py.
//
// The recursive EXTEND / UNWIND bookkeeping of the CPU algorithm is replaced by a closed form that
// suits the GPU.  For one root-to-leaf path, merge the splits on the same feature: element j has
// a zero fraction z_j (product of cover(child) / cover(parent) over its splits) and a one fraction
// o_j in {0, 1} (does the row satisfy every split of that feature on the path).  The Shapley
// permutation weight |S|! (m - |S| - 1)! / m! is the Beta integral of t^|S| (1 - t)^(m-1-|S|), so
//     phi_j = v_leaf (o_j - z_j) * integral_0^1  prod_{i != j} (z_i (1 - t) + o_i t)  dt ,
// a polynomial of degree m - 1 <= 15 (m <= d <= 16 distinct features), integrated EXACTLY by the
// 8-point Gauss-Legendre rule.  Per (row, path) that is 8 m multiply-adds for the full products and
// 8 m for the leave-one-out sums (a multiplication by the reciprocal of the factor, which takes one
// of two path constants) - no recursion, no data-dependent control flow, no divisions in the inner
// loop.  It equals the EXTEND / UNWIND result to 1e-15 relative (tests/test_explain_gpu.py).
//
// Kernels: k_shap_leaves (leaf ids of every tree, in node order), k_shap_paths (one thread per leaf
// walks down from the root - pre-order numbering makes "is the leaf in the left subtree" the test
// leaf < right child - and writes the merged path record), k_shap_main (a CTA owns 256 rows in
// shared memory and streams a chunk of the forest's paths: constants of 8 paths are staged
// cooperatively, then every thread evaluates its two rows), k_shap_reduce (sums the path chunks in
// order and divides by the number of trees: deterministic).
#include "f16_tree.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>

extern "C" void f16_set_error(const char* fmt, ...);
extern "C" cudaError_t f16_malloc_async(void** p, size_t bytes, cudaStream_t st);

#define CUDA_TRY(x)                                                                     \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            f16_set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            rc = F16_ERR_CUDA;                                                          \
            goto done;                                                                  \
        }                                                                               \
    } while (0)

#define SH_Q 8            // Gauss-Legendre points: exact up to degree 15
#define SH_PC 8           // paths whose constants are staged together
#define SH_NT 128
#define SH_R 2            // rows per thread
#define SH_ROWS (SH_NT * SH_R)

// nodes t_q and weights w_q of the 8-point rule mapped to [0, 1]
__constant__ double c_sh_t[SH_Q] = {0.019855071751231912, 0.10166676129318664, 0.2372337950418355, 0.4082826787521751,
                                    0.5917173212478248, 0.7627662049581645, 0.8983332387068134, 0.9801449282487681};
__constant__ double c_sh_w[SH_Q] = {0.05061426814518853, 0.11119051722668721, 0.15685332293894344, 0.18134189168918083,
                                    0.18134189168918083, 0.15685332293894344, 0.11119051722668721, 0.05061426814518853};

struct __align__(16) ShapPath {
    double v;             // leaf value of the requested class (class fraction)
    int32_t m;            // distinct features on the path
    int32_t pad;
    double z[F16_MAX_D];  // zero fractions
    float lo[F16_MAX_D];  // the row satisfies element j  <=>  lo < x[f] <= hi
    float hi[F16_MAX_D];
    int8_t f[F16_MAX_D];
};

// one warp per tree: leaf node ids in increasing order
__global__ void k_shap_leaves(const F16Node* __restrict__ nodes, int node_cap, const int32_t* __restrict__ node_count,
                              int32_t* __restrict__ leaf_ids, int leaf_stride, int32_t* __restrict__ n_leaves) {
    const int t = blockIdx.x, lane = threadIdx.x;
    const F16Node* tn = nodes + (size_t)t * node_cap;
    const int nc = min(node_count[t], node_cap);
    int32_t* out = leaf_ids + (size_t)t * leaf_stride;
    int run = 0;
    for (int base = 0; base < nc; base += 32) {
        const int i = base + lane;
        const bool leaf = i < nc && tn[i].feature < 0;
        const unsigned bal = __ballot_sync(0xffffffffu, leaf);
        if (leaf) out[run + __popc(bal & ((1u << lane) - 1u))] = i;
        run += __popc(bal);
    }
    if (lane == 0) n_leaves[t] = run;
}

__global__ void k_shap_paths(const F16Node* __restrict__ nodes, int node_cap, const int32_t* __restrict__ leaf_ids,
                             int leaf_stride, const int32_t* __restrict__ n_leaves, const long long* __restrict__ path_off,
                             int klass, ShapPath* __restrict__ paths) {
    const int t = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_leaves[t]) return;
    const F16Node* tn = nodes + (size_t)t * node_cap;
    const int leaf = leaf_ids[(size_t)t * leaf_stride + k];
    double z[F16_MAX_D];
    float lo[F16_MAX_D], hi[F16_MAX_D];
    unsigned present = 0;
#pragma unroll
    for (int f = 0; f < F16_MAX_D; f++) { z[f] = 1.0; lo[f] = -INFINITY; hi[f] = INFINITY; }
    int i = 0;
    while (i != leaf) {
        const F16Node nd = tn[i];
        const bool left = leaf < nd.right;                 // pre-order: the left subtree is (i, right)
        const int child = left ? i + 1 : nd.right;
        const F16Node ch = tn[child];
        const double ratio = ((double)ch.c0 + (double)ch.c1) / ((double)nd.c0 + (double)nd.c1);
        const int f = nd.feature;
        // sklearn goes left iff float32 x <= float64 thr, i.e. iff x <= thr rounded down to float32
        const float tf = __double2float_rd(nd.thr);
        z[f] *= ratio;
        if (left) hi[f] = fminf(hi[f], tf); else lo[f] = fmaxf(lo[f], tf);
        present |= 1u << f;
        i = child;
    }
    ShapPath* P = paths + path_off[t] + k;
    const F16Node lf = tn[leaf];
    const double cover = (double)lf.c0 + (double)lf.c1;
    P->v = (klass == 0 ? (double)lf.c0 : (double)lf.c1) / cover;
    int m = 0;
    for (int f = 0; f < F16_MAX_D; f++) {
        if ((present >> f) & 1u) { P->z[m] = z[f]; P->lo[m] = lo[f]; P->hi[m] = hi[f]; P->f[m] = (int8_t)f; m++; }
    }
    P->m = m; P->pad = 0;
}

__global__ void __launch_bounds__(SH_NT) k_shap_main(const ShapPath* __restrict__ paths, long long n_paths,
                                                     long long paths_per_chunk, const float* __restrict__ X, int n,
                                                     int dp, int d, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char sh_raw[];
    float* xs = reinterpret_cast<float*>(sh_raw);                           // [16][ROWS]
    double* phi_s = reinterpret_cast<double*>(xs + F16_MAX_D * SH_ROWS);    // [16][ROWS]
    double* cA = phi_s + F16_MAX_D * SH_ROWS;                               // [PC][16][Q]  z (1 - t)
    double* cR = cA + SH_PC * F16_MAX_D * SH_Q;                             // 1 / A
    double* cD = cR + SH_PC * F16_MAX_D * SH_Q;                             // 1 / (t + A) - 1 / A
    double* cZ = cD + SH_PC * F16_MAX_D * SH_Q;                             // [PC][16]
    double* cV = cZ + SH_PC * F16_MAX_D;                                    // [PC]
    float* cLo = reinterpret_cast<float*>(cV + SH_PC);                      // [PC][16]
    float* cHi = cLo + SH_PC * F16_MAX_D;
    int* cF = reinterpret_cast<int*>(cHi + SH_PC * F16_MAX_D);
    int* cM = cF + SH_PC * F16_MAX_D;                                       // [PC]
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * SH_ROWS;
    for (int i = tid; i < SH_ROWS * dp; i += SH_NT) {
        const int r = i / dp, f = i - r * dp;
        xs[f * SH_ROWS + r] = (row0 + r < n) ? X[(size_t)(row0 + r) * dp + f] : 0.f;
    }
    for (int i = tid; i < F16_MAX_D * SH_ROWS; i += SH_NT) phi_s[i] = 0.0;
    const long long p_begin = (long long)blockIdx.y * paths_per_chunk;
    const long long p_end = min(n_paths, p_begin + paths_per_chunk);
    for (long long pb = p_begin; pb < p_end; pb += SH_PC) {
        const int npc = (int)min((long long)SH_PC, p_end - pb);
        __syncthreads();                                   // the previous group's constants are consumed
        for (int i = tid; i < npc * F16_MAX_D; i += SH_NT) {
            const int pi = i / F16_MAX_D, e = i % F16_MAX_D;
            const ShapPath* P = paths + pb + pi;
            if (e < P->m) { cZ[i] = P->z[e]; cLo[i] = P->lo[e]; cHi[i] = P->hi[e]; cF[i] = P->f[e]; }
            if (e == 0) { cV[pi] = P->v; cM[pi] = P->m; }
        }
        for (int i = tid; i < npc * F16_MAX_D * SH_Q; i += SH_NT) {
            const int pi = i / (F16_MAX_D * SH_Q), e = (i / SH_Q) % F16_MAX_D, q = i % SH_Q;
            const ShapPath* P = paths + pb + pi;
            if (e < P->m) {
                const double A = P->z[e] * (1.0 - c_sh_t[q]);
                const double rA = 1.0 / A, rB = 1.0 / (c_sh_t[q] + A);
                cA[i] = A; cR[i] = rA; cD[i] = rB - rA;
            }
        }
        __syncthreads();
        for (int pi = 0; pi < npc; pi++) {
            const int m = cM[pi];
            double G[SH_R][SH_Q];
            unsigned mask[SH_R];
#pragma unroll
            for (int r = 0; r < SH_R; r++) {
                mask[r] = 0;
#pragma unroll
                for (int q = 0; q < SH_Q; q++) G[r][q] = c_sh_w[q];
            }
            for (int e = 0; e < m; e++) {
                const int ce = pi * F16_MAX_D + e;
                const int f = cF[ce];
                const float lo = cLo[ce], hi = cHi[ce];
                const double* a = cA + ce * SH_Q;
#pragma unroll
                for (int r = 0; r < SH_R; r++) {
                    const float x = xs[f * SH_ROWS + tid + r * SH_NT];
                    const bool one = (x > lo) && (x <= hi);
                    mask[r] |= (one ? 1u : 0u) << e;
                    const double od = one ? 1.0 : 0.0;
#pragma unroll
                    for (int q = 0; q < SH_Q; q++) G[r][q] *= fma(od, c_sh_t[q], a[q]);
                }
            }
            const double v = cV[pi];
            for (int e = 0; e < m; e++) {
                const int ce = pi * F16_MAX_D + e;
                const int f = cF[ce];
                const double z = cZ[ce];
                const double* rr = cR + ce * SH_Q;
                const double* dd = cD + ce * SH_Q;
#pragma unroll
                for (int r = 0; r < SH_R; r++) {
                    const double od = ((mask[r] >> e) & 1u) ? 1.0 : 0.0;
                    double S = 0.0;
#pragma unroll
                    for (int q = 0; q < SH_Q; q++) S = fma(G[r][q], fma(od, dd[q], rr[q]), S);
                    phi_s[f * SH_ROWS + tid + r * SH_NT] += S * (od - z) * v;
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < SH_ROWS * d; i += SH_NT) {
        const int r = i / d, f = i - r * d;
        if (row0 + r < n) partial[((size_t)blockIdx.y * n + (size_t)(row0 + r)) * d + f] = phi_s[f * SH_ROWS + r];
    }
}

__global__ void k_shap_reduce(const double* __restrict__ partial, int n_chunks, long long nd, double inv_trees,
                              double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nd) return;
    double s = 0.0;
    for (int c = 0; c < n_chunks; c++) s += partial[(size_t)c * nd + i];
    out[i] = s * inv_trees;
}

static size_t shap_smem_bytes() {
    return sizeof(float) * F16_MAX_D * SH_ROWS + sizeof(double) * F16_MAX_D * SH_ROWS +
           sizeof(double) * (3 * SH_PC * F16_MAX_D * SH_Q + SH_PC * F16_MAX_D + SH_PC) +
           sizeof(float) * 2 * SH_PC * F16_MAX_D + sizeof(int) * (SH_PC * F16_MAX_D + SH_PC);
}

// TreeExplainer(model).shap_values(X)[klass] for the fitted forest: phi_dev float64 [n][d].
// X_dev: float32 [n][dp] rows (f16_gather_rows_f32).  Synchronises the stream once (path count).
extern "C" int f16_forest_shap(const f16_forest* F, const float* X_dev, int64_t n, int32_t klass, double* phi_dev,
                               void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!F || !X_dev || !phi_dev || n < 0 || n > 0x7fffffff || (klass != 0 && klass != 1)) {
        f16_set_error("f16_forest_shap: bad arguments");
        return F16_ERR_INVALID;
    }
    if (n == 0) return F16_OK;
    int rc = F16_OK;
    const int T = F->n_trees, d = F->d, dp = F->dp;
    const int leaf_stride = F->node_cap / 2 + 1;
    int32_t *leaf_ids = nullptr, *n_leaves = nullptr;
    long long* path_off = nullptr;
    ShapPath* paths = nullptr;
    double* partial = nullptr;
    std::vector<int32_t> nl(T);
    std::vector<long long> off(T + 1, 0);
    long long n_paths = 0, ppc = 0;
    int n_chunks = 1, max_leaves = 0;
    int32_t fit_err = 0;
    size_t smem = shap_smem_bytes();
    CUDA_TRY(f16_malloc_async((void**)&leaf_ids, sizeof(int32_t) * (size_t)T * leaf_stride, st));
    CUDA_TRY(f16_malloc_async((void**)&n_leaves, sizeof(int32_t) * T, st));
    k_shap_leaves<<<T, 32, 0, st>>>(F->nodes, F->node_cap, F->node_count, leaf_ids, leaf_stride, n_leaves);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(nl.data(), n_leaves, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(&fit_err, F->err, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (fit_err != 0) {        // a forest whose fit was aborted holds dangling links: refuse to walk it
        f16_set_error("f16_forest_shap: the forest's fit failed on the device (code %d)", fit_err);
        rc = fit_err;
        goto done;
    }
    for (int t = 0; t < T; t++) { off[t + 1] = off[t] + nl[t]; if (nl[t] > max_leaves) max_leaves = nl[t]; }
    n_paths = off[T];
    CUDA_TRY(f16_malloc_async((void**)&path_off, sizeof(long long) * (T + 1), st));
    CUDA_TRY(cudaMemcpyAsync(path_off, off.data(), sizeof(long long) * (T + 1), cudaMemcpyHostToDevice, st));
    CUDA_TRY(f16_malloc_async((void**)&paths, sizeof(ShapPath) * (size_t)(n_paths > 0 ? n_paths : 1), st));
    if (max_leaves > 0) {
        dim3 g((unsigned)((max_leaves + 127) / 128), (unsigned)T);
        k_shap_paths<<<g, 128, 0, st>>>(F->nodes, F->node_cap, leaf_ids, leaf_stride, n_leaves, path_off, klass, paths);
        CUDA_TRY(cudaGetLastError());
    }
    {
        // enough (row tile, path chunk) CTAs to fill the chip several times over, chunks of whole
        // staging groups
        const long long row_tiles = (n + SH_ROWS - 1) / SH_ROWS;
        long long want = (148 * 3 * 4 + row_tiles - 1) / row_tiles;
        if (want < 1) want = 1;
        if (want > 64) want = 64;
        ppc = (n_paths + want - 1) / want;
        ppc = (ppc + SH_PC - 1) / SH_PC * SH_PC;
        if (ppc < SH_PC) ppc = SH_PC;
        n_chunks = (int)((n_paths + ppc - 1) / ppc);
        if (n_chunks < 1) n_chunks = 1;
        CUDA_TRY(f16_malloc_async((void**)&partial, sizeof(double) * (size_t)n_chunks * (size_t)n * d, st));
        CUDA_TRY(cudaFuncSetAttribute(k_shap_main, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 g((unsigned)row_tiles, (unsigned)n_chunks);
        k_shap_main<<<g, SH_NT, smem, st>>>(paths, n_paths, ppc, X_dev, (int)n, dp, d, partial);
        CUDA_TRY(cudaGetLastError());
        const long long nd = (long long)n * d;
        k_shap_reduce<<<(unsigned)((nd + 255) / 256), 256, 0, st>>>(partial, n_chunks, nd, 1.0 / (double)T, phi_dev);
        CUDA_TRY(cudaGetLastError());
        f16_count_launch(4);
    }
done:
    if (leaf_ids) cudaFreeAsync(leaf_ids, st);
    if (n_leaves) cudaFreeAsync(n_leaves, st);
    if (path_off) cudaFreeAsync(path_off, st);
    if (paths) cudaFreeAsync(paths, st);
    if (partial) cudaFreeAsync(partial, st);
    return rc;
}
