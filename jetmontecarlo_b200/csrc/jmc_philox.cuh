// Counter-based RNG of the fused path: Philox-4x32-10 (Salmon et al., SC'11).
//
// Replaces the reference's per-coordinate `random.random()` (MT19937) calls at
// jetmontecarlo/utils/montecarlo_utils.py:28,53.  A counter RNG needs no state
// in memory: sample i, draw block b is philox(counter=(i_lo, i_hi, b, 0),
// key=(seed_lo, seed_hi)), so samples live only in registers, any GPU can
// produce any index range, and the union of samples is the same for every
// number of GPUs.
//
// Uniform conventions (mirrored by oracle/jmc_oracle.py: philox_uniforms53 /
// philox_words):
//   f64: block b yields uniforms 2b, 2b+1 = ((w[2j]>>5)*2^26 + (w[2j+1]>>6)) / 2^53
//        -- the same 53-bit construction CPython's random.random() uses;
//   f32: block b yields uniforms 4b..4b+3 = ((w[j]>>8) + 0.5) * 2^-24  in (0,1).
#pragma once
#include <cstdint>
#ifndef JMC_PHILOX_SPLIT
#define JMC_PHILOX_SPLIT 0       // 1: IMAD.HI + IMAD per product instead of one IMAD.WIDE (experiment switch)
#endif

namespace jmc {

struct PhiloxKey { uint32_t k0, k1; };

__host__ __device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2,
                                                      uint32_t& c3, uint32_t k0, uint32_t k1) {
#if defined(__CUDA_ARCH__) && !JMC_PHILOX_SPLIT
    // one IMAD.WIDE.U32 per product (hi and lo together) instead of IMAD.HI + IMAD: halves the issue slots
    // of the generator, which is the largest single cost of a sample
    uint32_t hi0, lo0, hi1, lo1;
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
    asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
        : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
#else
    const uint32_t hi0 = (uint32_t)(((uint64_t)0xD2511F53u * c0) >> 32);
    const uint32_t hi1 = (uint32_t)(((uint64_t)0xCD9E8D57u * c2) >> 32);
    const uint32_t lo0 = 0xD2511F53u * c0;
    const uint32_t lo1 = 0xCD9E8D57u * c2;
#endif
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
}

// four 32-bit words for (index, block) under key `seed`
__host__ __device__ __forceinline__ void philox4x32_10(uint64_t index, uint32_t block, uint64_t seed,
                                                       uint32_t w[4]) {
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = block, c3 = 0u;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

// The ten round keys of a launch (key schedule k += W per round).  A kernel that calls the generator in a hot
// loop computes them once; pin() makes them opaque to the compiler, which otherwise prefers to recompute the
// schedule with uniform-datapath adds in every iteration (10 issue slots per sample).
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
    __device__ __forceinline__ explicit PhiloxKeys(uint64_t seed) {
        uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            k0[r] = a; k1[r] = b;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
    }
    __device__ __forceinline__ void pin() {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int r = 0; r < 10; ++r) asm volatile("" : "+r"(k0[r]), "+r"(k1[r]));
#endif
    }
};

// index given as its two halves: a launch that keeps the high half uniform (no 2^32 crossing inside it) gets the
// first round's c1 ^ k0 and the second round's product of it for free (loop-invariant, hoisted by the compiler)
__device__ __forceinline__ void philox4x32_10(uint32_t index_lo, uint32_t index_hi, uint32_t block,
                                              const PhiloxKeys& K, uint32_t w[4]) {
    uint32_t c0 = index_lo, c1 = index_hi, c2 = block, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c0, c1, c2, c3, K.k0[r], K.k1[r]);
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}
__device__ __forceinline__ void philox4x32_10(uint64_t index, uint32_t block, const PhiloxKeys& K, uint32_t w[4]) {
    philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), block, K, w);
}

// two 53-bit uniforms in [0,1) from one block
__host__ __device__ __forceinline__ void philox_uniform53x2(uint64_t index, uint32_t block, uint64_t seed,
                                                            double& u0, double& u1) {
    uint32_t w[4];
    philox4x32_10(index, block, seed, w);
    u0 = ((double)(w[0] >> 5) * 67108864.0 + (double)(w[1] >> 6)) / 9007199254740992.0;
    u1 = ((double)(w[2] >> 5) * 67108864.0 + (double)(w[3] >> 6)) / 9007199254740992.0;
}

// 24-bit uniform strictly inside (0,1)
__host__ __device__ __forceinline__ float u32_to_unit_float(uint32_t w) {
    return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-08f;
}

}  // namespace jmc
