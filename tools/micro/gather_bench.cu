// Random 64-byte-row gather throughput on B200 by mechanism (what bounds the ExtraTrees GLOBAL-regime
// sweeps): rows of a 90 000 x 16 float32 matrix (5.8 MB, L2 resident) fetched through per-"tree" id
// lists, reduced to a min/max so the loads stay live.
//   mode 0: LDG.128, 4 threads per row, U loads in flight per thread   (what k_build_random does, U = 4)
//   mode 1: cp.async 16 B into a shared-memory tile (U per thread in flight), then LDS
//   mode 2: cp.async.bulk 64 B per row (one copy per row, mbarrier complete_tx), then LDS
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>
#include <random>

#define DP 16
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int U>
__global__ void __launch_bounds__(256) k_ldg(const float* __restrict__ X, const uint32_t* __restrict__ ids, int n, float* out) {
    const uint32_t* src = ids + (size_t)blockIdx.x * n;
    const int q = threadIdx.x & 3, sl = threadIdx.x >> 2;
    float mn = 1e30f, mx = -1e30f;
    for (int i0 = sl; i0 < n; i0 += 64 * U) {
        uint32_t id[U]; float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) { int i = i0 + u * 64; id[u] = src[i < n ? i : i0]; }
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)id[u] * DP) + q);
#pragma unroll
        for (int u = 0; u < U; u++) { mn = fminf(mn, fminf(fminf(v[u].x, v[u].y), fminf(v[u].z, v[u].w))); mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w))); }
    }
    if (mn > mx) out[blockIdx.x * 256 + threadIdx.x] = mn;
    else if (mx == 12345.f) out[0] = mx;
}

// mode 1 / 2: every WARP owns its own id list segment and a tile of TR rows in shared memory
template <int TR, int MODE>
__global__ void __launch_bounds__(256) k_tile(const float* __restrict__ X, const uint32_t* __restrict__ ids, int n, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    float* tile = reinterpret_cast<float*>(sm) + (size_t)warp * TR * DP;
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<float*>(sm) + (size_t)nw * TR * DP);
    const uint32_t bar = smem_u32(bars + warp);
    if (MODE == 2) {
        if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
    // the block's list is split over its warps
    const int per = (n + nw - 1) / nw;
    const int beg = warp * per, end = min(n, beg + per);
    const uint32_t* src = ids + (size_t)blockIdx.x * n;
    float mn = 1e30f, mx = -1e30f;
    uint32_t parity = 0;
    for (int base = beg; base < end; base += TR) {
        const int cnt = min(TR, end - base);
        if (MODE == 1) {
            // 4 lanes per row (16 B each): 8 rows per step
            for (int r = lane >> 2; r < cnt; r += 8) {
                const uint32_t id = src[base + r];
                const float* g = X + (size_t)id * DP + (lane & 3) * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + r * DP + (lane & 3) * 4)), "l"(g));
            }
            asm volatile("cp.async.commit_group;" ::);
            asm volatile("cp.async.wait_group 0;" ::);
            __syncwarp();
        } else {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(cnt * DP * 4) : "memory");
            __syncwarp();
            for (int r = lane; r < cnt; r += 32) {
                const uint32_t id = src[base + r];
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(tile + r * DP)), "l"(X + (size_t)id * DP), "r"(DP * 4), "r"(bar) : "memory");
            }
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            } while (!done);
            parity ^= 1;
        }
        // min/max of all 16 features: lane = (feature, row parity)
        for (int r = lane >> 4; r < cnt; r += 2) { float v = tile[r * DP + (lane & 15)]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        __syncwarp();
    }
    if (mn > mx) out[blockIdx.x * 256 + threadIdx.x] = mn;
    else if (mx == 12345.f) out[0] = mx;
}

// mode 3: what a tree CTA could do in its idle 32 KB: every warp owns two 32-row tiles (double buffer),
// loads the NEXT tile's ids and issues its 16-byte cp.async gathers before it reduces the current tile.
__global__ void __launch_bounds__(256, 4) k_pipe(const float* __restrict__ X, const uint32_t* __restrict__ ids, int n, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* tile = reinterpret_cast<float*>(sm) + (size_t)warp * 2 * 32 * DP;      // [2][32][16]
    const uint32_t* src = ids + (size_t)blockIdx.x * n;
    const int ntiles = (n + 31) / 32;
    float mn = 1e30f, mx = -1e30f;
    auto issue = [&](int t, int buf, uint32_t myid) {
        // lane l holds the id of row 32 t + l; lane handles (row = j * 8 + l / 4, quarter = l % 4), j = 0..3
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int r = j * 8 + (lane >> 2);
            const uint32_t id = __shfl_sync(0xffffffffu, myid, r);
            if (t * 32 + r < n) {
                const float* g = X + (size_t)id * DP + (lane & 3) * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + (buf * 32 + r) * DP + (lane & 3) * 4)), "l"(g));
            }
        }
        asm volatile("cp.async.commit_group;" ::);
    };
    int t = warp;                       // tiles warp, warp + 8, ...
    uint32_t idc = (t < ntiles && t * 32 + lane < n) ? src[t * 32 + lane] : 0u;
    int buf = 0;
    if (t < ntiles) issue(t, 0, idc);
    uint32_t idn = (t + 8 < ntiles && (t + 8) * 32 + lane < n) ? src[(t + 8) * 32 + lane] : 0u;
    for (; t < ntiles; t += 8, buf ^= 1) {
        if (t + 8 < ntiles) issue(t + 8, buf ^ 1, idn);
        const int t2 = t + 16;
        idn = (t2 < ntiles && t2 * 32 + lane < n) ? src[t2 * 32 + lane] : 0u;
        if (t + 8 < ntiles) asm volatile("cp.async.wait_group 1;" ::); else asm volatile("cp.async.wait_group 0;" ::);
        __syncwarp();
        const int cnt = min(32, n - t * 32);
        for (int r = lane >> 4; r < cnt; r += 2) { float v = tile[(buf * 32 + r) * DP + (lane & 15)]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        __syncwarp();
    }
    if (mn > mx) out[blockIdx.x * 256 + threadIdx.x] = mn;
    else if (mx == 12345.f) out[0] = mx;
}

template <int U>
__global__ void __launch_bounds__(256, 4) k_ldg_occ4(const float* __restrict__ X, const uint32_t* __restrict__ ids, int n, float* out) {
    extern __shared__ __align__(128) unsigned char sm[];      // 40 KB of padding: 4 CTAs per SM like the tree kernel
    const uint32_t* src = ids + (size_t)blockIdx.x * n;
    const int q = threadIdx.x & 3, sl = threadIdx.x >> 2;
    float mn = 1e30f, mx = -1e30f;
    uint32_t idn[U];
#pragma unroll
    for (int u = 0; u < U; u++) { int i = sl + u * 64; idn[u] = src[i < n ? i : sl]; }
    for (int i0 = sl; i0 < n; i0 += 64 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = __ldg(reinterpret_cast<const float4*>(X + (size_t)idn[u] * DP) + q);
        const int i1 = i0 + 64 * U;
        if (i1 < n) {
#pragma unroll
            for (int u = 0; u < U; u++) { int i = i1 + u * 64; idn[u] = src[i < n ? i : i1]; }
        }
#pragma unroll
        for (int u = 0; u < U; u++) { mn = fminf(mn, fminf(fminf(v[u].x, v[u].y), fminf(v[u].z, v[u].w))); mx = fmaxf(mx, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w))); }
    }
    if (mn > mx) { out[blockIdx.x * 256 + threadIdx.x] = mn; sm[threadIdx.x] = 1; }
    else if (mx == 12345.f) out[0] = mx;
}

int main(int argc, char** argv) {
    const int n = 90000, trees = argc > 1 ? atoi(argv[1]) : 1184;
    std::vector<float> hX((size_t)n * DP);
    std::mt19937 rng(1);
    for (auto& v : hX) v = (float)(rng() % 1000);
    std::vector<uint32_t> hid((size_t)trees * n);
    for (int t = 0; t < trees; t++) {
        uint32_t* p = hid.data() + (size_t)t * n;
        for (int i = 0; i < n; i++) p[i] = i;
        std::shuffle(p, p + n, rng);
    }
    float *X, *out; uint32_t* ids;
    cudaMalloc(&X, hX.size() * 4); cudaMalloc(&ids, hid.size() * 4); cudaMalloc(&out, (size_t)trees * 256 * 4);
    cudaMemcpy(X, hX.data(), hX.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ids, hid.data(), hid.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t e = cudaGetLastError();
        printf("%-44s %8.3f ms  %7.1f G rows/s  %6.2f TB/s (64 B rows)%s\n", name, ms, (double)trees * n / ms / 1e6, (double)trees * n * 64 / ms / 1e9,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run("LDG.128 x4/row, 256 thr, 4 in flight", [&] { k_ldg<4><<<trees, 256>>>(X, ids, n, out); });
    run("LDG.128 x4/row, 256 thr, 8 in flight", [&] { k_ldg<8><<<trees, 256>>>(X, ids, n, out); });
    run("LDG.128 x4/row, 256 thr, 16 in flight", [&] { k_ldg<16><<<trees, 256>>>(X, ids, n, out); });
    run("LDG.128 x4/row, 128 thr, 8 in flight", [&] { k_ldg<8><<<trees, 128>>>(X, ids, n, out); });
    {
        auto smem = [](int nw, int tr) { return (size_t)nw * tr * DP * 4 + 64; };
        cudaFuncSetAttribute(k_tile<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(8, 256));
        cudaFuncSetAttribute(k_tile<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(8, 256));
        cudaFuncSetAttribute(k_tile<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(8, 128));
        cudaFuncSetAttribute(k_tile<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem(8, 128));
        run("cp.async 16 B, tile 256 rows/warp, 4 warps", [&] { k_tile<256, 1><<<trees, 128, smem(4, 256)>>>(X, ids, n, out); });
        run("cp.async 16 B, tile 128 rows/warp, 8 warps", [&] { k_tile<128, 1><<<trees, 256, smem(8, 128)>>>(X, ids, n, out); });
        run("bulk 64 B/row, tile 256 rows/warp, 4 warps", [&] { k_tile<256, 2><<<trees, 128, smem(4, 256)>>>(X, ids, n, out); });
        run("bulk 64 B/row, tile 128 rows/warp, 8 warps", [&] { k_tile<128, 2><<<trees, 256, smem(8, 128)>>>(X, ids, n, out); });
        run("bulk 64 B/row, tile 256 rows/warp, 1 warp", [&] { k_tile<256, 2><<<trees, 32, smem(1, 256)>>>(X, ids, n, out); });
        run("bulk 64 B/row, tile 128 rows/warp, 2 warps", [&] { k_tile<128, 2><<<trees, 64, smem(2, 128)>>>(X, ids, n, out); });
    }
    {
        cudaFuncSetAttribute(k_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
        cudaFuncSetAttribute(k_ldg_occ4<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
        cudaFuncSetAttribute(k_ldg_occ4<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
        run("LDG x4 in flight + id prefetch, 4 CTAs/SM (tree kernel today)", [&] { k_ldg_occ4<4><<<trees, 256, 40 * 1024>>>(X, ids, n, out); });
        run("LDG x8 in flight + id prefetch, 4 CTAs/SM", [&] { k_ldg_occ4<8><<<trees, 256, 40 * 1024>>>(X, ids, n, out); });
        run("cp.async 32-row tiles, double buffered per warp, 4 CTAs/SM", [&] { k_pipe<<<trees, 256, 40 * 1024>>>(X, ids, n, out); });
    }
    return 0;
}
