// polar_warp.cuh -- warp-level bit-packed polar transform and Philox counter RNG (device only).
#pragma once
#include "polar_common.cuh"

namespace polar {

// In-word stages for a row shorter than / equal to 32 positions (m = log2 n stages, n <= 32) or the
// 5 in-word stages of a longer row.
__device__ __forceinline__ uint32_t ptransform_rt(uint32_t x, int m) {
  if (m > 0) x ^= (x >> 1) & 0x55555555u;
  if (m > 1) x ^= (x >> 2) & 0x33333333u;
  if (m > 2) x ^= (x >> 4) & 0x0F0F0F0Fu;
  if (m > 3) x ^= (x >> 8) & 0x00FF00FFu;
  if (m > 4) x ^= (x >> 16) & 0x0000FFFFu;
  return x;
}

// One codeword per warp, R words per lane: word index wd = r*32 + lane (R = max(1, n/1024)).
// x = u.G over GF(2): stage s pairs position d with d + 2^s (bit s of d clear)
// (my_sn/fec/polar/enc.py:70-74 / (c @ G) % 2 of x_run_sn_polar/polar/enc.py:42).
// Lanes >= nw (n < 1024) carry zeros and are ignored by the caller.
template <int R>
__device__ __forceinline__ void warp_polar_transform(uint32_t (&x)[R], int m, int nw) {
#pragma unroll
  for (int r = 0; r < R; ++r) x[r] = ptransform_rt(x[r], m);
  const int lane = threadIdx.x & 31;
  // cross-word stages inside the warp: word distance d = 1,2,4,8,16 (only while d < nw)
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    if (d < nw) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x[r], d);
        if (!(lane & d)) x[r] ^= y;
      }
    }
  }
  // cross-register stages: word distance 32*d
#pragma unroll
  for (int d = 1; d < R; d <<= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (!(r & d)) x[r] ^= x[r + d];
  }
}

// Philox4x32-10 (Salmon et al., SC'11): counter-based, stateless.
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      c0 = hi1 ^ c1 ^ a; c1 = lo1; c2 = hi0 ^ c3 ^ b; c3 = lo0;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
  }
};

// Box-Muller: two uint32 -> two independent N(0,1).
// Special-function-unit version: lg2 / sqrt / sin / cos are one MUFU instruction each (lg2.approx: 2^-22 relative,
// sin/cos.approx: 2^-20.9 absolute on [0, 2pi]) instead of ~55 instructions for logf + sqrtf + sincospif -- the front end
// was bound by instruction issue (1.7 TB/s of logits), and errors of 1e-6 in a noise sample are far below anything a
// Monte-Carlo BER estimate resolves (tests/test_gpu_link.py checks mean, variance, kurtosis, tail mass, I/Q correlation).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]; float rounding may give 1.0 -> r=0, fine
  const float th = (float)b * (2.3283064365386963e-10f * 6.283185307179586f);   // [0, 2pi]
  float l2, r, s, c;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaxf(l2 * -1.3862943611198906f, 0.0f)));   // -2 ln u1
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(th));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(th));
  return make_float2(r * c, r * s);
}

}  // namespace polar
