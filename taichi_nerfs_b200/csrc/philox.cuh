// philox.cuh — Philox4x32-10 counter-based RNG (Salmon et al., SC'11; pinned on the Random123 known-answer vectors by
// tests/test_oracle.py::test_philox_known_answers through the oracle's restatement).  Used by the ray-batch sampler
// (rays.cu) and the occupancy-grid cell sampler (grid.cu): keyed draws make both launches replayable and identical on
// every rank.
#pragma once
#include <stdint.h>

struct Philox4 {
    uint32_t v[4];
};

// counter = (c0,c1,c2,c3), key = (k0,k1)
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{{c0, c1, c2, c3}};
}


// multiply-shift range reduction: floor(r * n / 2^32) in [0, n)
__device__ __forceinline__ uint32_t philox_below(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * (uint64_t)n) >> 32); }
// 24 random bits -> [0, 1)
__device__ __forceinline__ float philox_unit(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-8f; }
