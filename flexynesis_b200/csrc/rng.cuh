// Counter-based RNG shared by the dropout (bn_act.cu) and Gaussian-noise (vae.cu) kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fxn {

// ---- counter-based RNG (Philox4x32-10) for dropout masks: backward regenerates the mask from (seed, index) ----
__device__ __forceinline__ uint4 philox4x32(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1) {
  uint32_t c2 = 0x5bd1e995u, c3 = 0x1b873593u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
// keep-flags of the 8 consecutive elements starting at linear index idx8*8
__device__ __forceinline__ uint32_t dropout_keep8(unsigned long long seed, unsigned long long idx8, float p) {
  const uint32_t thr = static_cast<uint32_t>(p * 65536.0f);   // 16-bit uniforms: keep iff u16 >= p*2^16
  const uint4 r = philox4x32(static_cast<uint32_t>(idx8), static_cast<uint32_t>(idx8 >> 32),
                             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  uint32_t bits = 0;
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bits |= ((w[j] & 0xFFFFu) >= thr ? 1u : 0u) << (2 * j);
    bits |= ((w[j] >> 16) >= thr ? 1u : 0u) << (2 * j + 1);
  }
  return bits;
}

// per-step seed: the host seed mixed with a device-resident step counter, so a replayed CUDA graph still draws a
// fresh mask every step while forward and backward of the same step agree
__device__ __forceinline__ unsigned long long step_seed(unsigned long long seed, const long long* seed_dev) {
  return seed_dev ? seed + 0x9E3779B97F4A7C15ull * static_cast<unsigned long long>(*seed_dev + 1) : seed;
}

// two standard normals from two 32-bit uniforms (Box-Muller; u1 in (0, 1])
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  const float u2 = static_cast<float>(b >> 8) * (1.0f / 16777216.0f);
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

}  // namespace fxn
