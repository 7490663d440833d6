// Counter-based normal variates for the device-side noise draws (cloak eps, class-balance augmentation).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sept {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (ctr_lo, ctr_hi, 0, 0), key = seed --------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
    uint32_t c2 = 0u, c3 = 0u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)

__device__ __forceinline__ float4 normal4(uint64_t seed, uint64_t offset, uint32_t i4, float std_) {
    const uint64_t ctr = offset + i4;
    const uint4 r = philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
    const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
    float s0, c0, s1, c1;
    sincosf(6.28318530717958648f * u01(r.y), &s0, &c0);
    sincosf(6.28318530717958648f * u01(r.w), &s1, &c1);
    return make_float4(std_ * r0 * c0, std_ * r0 * s0, std_ * r1 * c1, std_ * r1 * s1);
}

}  // namespace sept
