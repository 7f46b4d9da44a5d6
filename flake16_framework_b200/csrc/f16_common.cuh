// Common helpers for the flake16 `scores` hot path kernels (sm_100a).
//
// RNG contracts restated here:
//   * numpy legacy RandomState / MT19937: init_genrand seeding, 32-bit output tempering,
//     masked-rejection bounded integers (numpy/random/src/distributions/distributions.c:
//     buffered_bounded_masked_uint32) - used by sklearn for the per-tree seeds
//     (sklearn/ensemble/_base.py:_set_random_states), the bootstrap draw
//     (sklearn/ensemble/_forest.py:_generate_sample_indices) and the splitter seed
//     (sklearn/tree/_splitter.pyx:155).
//   * sklearn `our_rand_r` xorshift32 (sklearn/utils/_random.pxd:20-34), `rand_int`,
//     `rand_uniform` (sklearn/tree/_utils.pyx:51-61).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
extern "C" void f16_count_launch(int n);
extern "C" long long f16_launch_count(int reset);

#define F16_OK 0
#define F16_ERR_INVALID -1
#define F16_ERR_CUDA -2
#define F16_ERR_OVERFLOW -3   // device-side capacity / weight overflow flagged by a kernel
#define F16_ERR_NOMEM -4

#define F16_ID_BITS 24
#define F16_ID_MASK 0x00FFFFFFu
#define F16_MAX_ROWS (1 << F16_ID_BITS)
#define F16_MAX_W 127
#define F16_MAX_D 16

// packed sample entry: bits 0..23 row id, bits 24..30 bootstrap weight (1..127), bit 31 label
__host__ __device__ __forceinline__ uint32_t f16_pack(uint32_t id, uint32_t w, uint32_t y) {
    return id | (w << F16_ID_BITS) | (y << 31);
}
__host__ __device__ __forceinline__ uint32_t f16_id(uint32_t e) { return e & F16_ID_MASK; }
__host__ __device__ __forceinline__ uint32_t f16_w(uint32_t e) { return (e >> F16_ID_BITS) & 0x7Fu; }
__host__ __device__ __forceinline__ uint32_t f16_y(uint32_t e) { return e >> 31; }

// ------------------------------------------------------------------ xorshift (sklearn our_rand_r)
__host__ __device__ __forceinline__ uint32_t f16_rand_r(uint32_t* s) {
    if (*s == 0) *s = 1;                       // DEFAULT_SEED
    *s ^= (uint32_t)(*s << 13);
    *s ^= (uint32_t)(*s >> 17);
    *s ^= (uint32_t)(*s << 5);
    return *s % 0x80000000u;                   // % (RAND_R_MAX + 1)
}
__host__ __device__ __forceinline__ int f16_rand_int(int lo, int hi, uint32_t* s) {
    return lo + (int)(f16_rand_r(s) % (uint32_t)(hi - lo));
}
#ifdef __CUDACC__
// rand_int for ranges up to 16 (the feature draws, d <= 16): r % m by one multiply-high with
// floor(2^32 / m) and a single correction - the estimate is never more than one below r / m
// (r < 2^32), so the result equals the `%` above for every r.
static __constant__ uint32_t f16_magic16[17] = {0x00000000u, 0xffffffffu, 0x80000000u, 0x55555555u, 0x40000000u, 0x33333333u, 0x2aaaaaaau, 0x24924924u, 0x20000000u, 0x1c71c71cu, 0x19999999u, 0x1745d174u, 0x15555555u, 0x13b13b13u, 0x12492492u, 0x11111111u, 0x10000000u};
__device__ __forceinline__ int f16_rand_int_small(int lo, int hi, uint32_t* s) {
    const uint32_t r = f16_rand_r(s);
    const uint32_t m = (uint32_t)(hi - lo);                       // 1 .. 16
    const uint32_t magic = f16_magic16[m];                        // min(2^32 - 1, floor(2^32 / m))
    uint32_t rem = r - __umulhi(r, magic) * m;
    if (rem >= m) rem -= m;
    return lo + (int)rem;
}
#endif
// ((high - low) * r / RAND_R_MAX) + low, float64, this exact operation order, no FMA.
__device__ __forceinline__ double f16_rand_uniform(double lo, double hi, uint32_t* s) {
    double r = (double)f16_rand_r(s);
    return __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(hi, lo), r), 2147483647.0), lo);
}

// ------------------------------------------------------------------ MT19937
struct F16MT {
    uint32_t mt[624];
    int idx;
};
__host__ __device__ inline void f16_mt_seed(F16MT* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; i++)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
__host__ __device__ __forceinline__ uint32_t f16_mt_twist(uint32_t a, uint32_t b) {
    uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__host__ __device__ __forceinline__ uint32_t f16_mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__host__ __device__ inline uint32_t f16_mt_next(F16MT* s) {
    if (s->idx >= 624) {
        int i;
        for (i = 0; i < 624 - 397; i++) s->mt[i] = s->mt[i + 397] ^ f16_mt_twist(s->mt[i], s->mt[i + 1]);
        for (; i < 623; i++) s->mt[i] = s->mt[i + (397 - 624)] ^ f16_mt_twist(s->mt[i], s->mt[i + 1]);
        s->mt[623] = s->mt[396] ^ f16_mt_twist(s->mt[623], s->mt[0]);
        s->idx = 0;
    }
    return f16_mt_temper(s->mt[s->idx++]);
}
__host__ __device__ __forceinline__ uint32_t f16_gen_mask(uint32_t max) {
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    return mask;
}
// RandomState.randint(0, hi): rng = hi - 1; draw & mask, reject > rng.
__host__ __device__ inline uint32_t f16_mt_randint(F16MT* s, uint32_t hi) {
    uint32_t rng = hi - 1u;
    if (rng == 0) return 0;
    uint32_t mask = f16_gen_mask(rng), v;
    while ((v = (f16_mt_next(s) & mask)) > rng) {}
    return v;
}

// ------------------------------------------------------------------ warp / block helpers
#define F16_FULL 0xffffffffu

__device__ __forceinline__ int f16_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int f16_warp() { return threadIdx.x >> 5; }

__device__ __forceinline__ unsigned long long f16_shfl_up_u64(unsigned long long v, int d) {
    return __shfl_up_sync(F16_FULL, v, d);
}
__device__ __forceinline__ unsigned long long f16_warp_incl_scan_u64(unsigned long long v) {
    int lane = f16_lane();
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(F16_FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ unsigned long long f16_warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(F16_FULL, v, d);
    return v;
}
