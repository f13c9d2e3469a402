// polar_common.cuh -- shared device/host helpers for the sm_100a polar kernels.
// Semantics follow SURVEY.md Appendix A (validated restatement of the reference decoders).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PHD __host__ __device__ __forceinline__
#define PDEV __device__ __forceinline__
#else
#define PHD inline
#define PDEV inline
#endif

namespace polar {

constexpr float kLlrMax = 30.0f;  // polar_sc.py:21 / polar_scl.py:35

PHD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; memcpy(&u, &x, 4); return u;
#endif
}
PHD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; memcpy(&x, &u, 4); return x;
#endif
}

// f: clip both inputs to +-30, then sign.sign.min(|a|,|b|)   (polar_sc.py:35-36,46).
// |clip(x)| = min(|x|,30), so the magnitude is min(|a|,|b|,30); the sign is the xor of the sign
// bits.  When an input is +-0 the reference yields +-0 (sign(0)=0); so does this (sign of a zero
// LLR is never observable: the leaf rule is llr<=0 -> 1 and g adds it).
PHD float f_minsum(float a, float b) {
  float mag = fminf(fminf(fabsf(a), fabsf(b)), kLlrMax);
  uint32_t sgn = (f2u(a) ^ f2u(b)) & 0x80000000u;
  return u2f(f2u(mag) | sgn);
}
// same, for inputs already known to lie in [-30, 30] (outputs of f): the clip is the identity.
PHD float f_minsum_noclip(float a, float b) {
  float mag = fminf(fabsf(a), fabsf(b));
  uint32_t sgn = (f2u(a) ^ f2u(b)) & 0x80000000u;
  return u2f(f2u(mag) | sgn);
}
// g: (1-2u).a + b, unclipped, one rounding (polar_sc.py:49-53).  signmask = u ? 0x80000000 : 0.
PHD float g_minsum(float a, float b, uint32_t signmask) { return u2f(f2u(a) ^ signmask) + b; }

// GF(2) polar transform of the low 2^T bits of x: x[d] ^= x[d + 2^s] for every d with bit s clear
// (my_sn/fec/polar/enc.py:70-74).  It is an involution: u = T(beta), beta = T(u).
template <int T>
PHD uint32_t ptransform(uint32_t x) {
  if (T > 0) x ^= (x >> 1) & 0x55555555u;
  if (T > 1) x ^= (x >> 2) & 0x33333333u;
  if (T > 2) x ^= (x >> 4) & 0x0F0F0F0Fu;
  if (T > 3) x ^= (x >> 8) & 0x00FF00FFu;
  if (T > 4) x ^= (x >> 16) & 0x0000FFFFu;
  return x;
}

// SC decode of a 2^T-leaf subtree held entirely in registers by ONE thread (compile-time indices
// only).  x: node LLRs; fm: frozen bits of the 2^T leaves (bit j = leaf j); returns the partial sums
// beta (bit j) and the decisions u.  Restates polar_sc.py:54-98 (f -> left -> g -> right -> combine).
// Exact shortcuts (uniform over the warp, the frozen pattern is shared):
//  rate-0 (all frozen): u = beta = 0.
//  rate-1 (none frozen): beta = hard decisions of x, u = T(beta) -- identical to the recursion
//   unless some x is exactly 0 (reference tie rule llr==0 -> 1), which falls through.
template <int T, bool CLIPPED = false>
struct SubTree {
  static constexpr int N = 1 << T, H = N / 2;
  static constexpr uint32_t FULL = (N == 32) ? 0xFFFFFFFFu : ((1u << N) - 1u);
  static constexpr uint32_t HALF = (1u << H) - 1u;
  PHD static uint32_t run(const float (&x)[N], uint32_t fm, uint32_t &u) {
    fm &= FULL;
    if (fm == FULL) { u = 0; return 0; }
    if (T >= 3 && fm == 0) {
      uint32_t hd = 0; bool zero = false;
#pragma unroll
      for (int j = 0; j < N; ++j) { hd |= (uint32_t)(x[j] < 0.0f) << j; zero |= (x[j] == 0.0f); }
      if (!zero) { u = ptransform<T>(hd); return hd; }
    }
    float y[H];
#pragma unroll
    for (int j = 0; j < H; ++j) y[j] = CLIPPED ? f_minsum_noclip(x[j], x[j + H]) : f_minsum(x[j], x[j + H]);
    uint32_t ul, ur;
    uint32_t bl = SubTree<T - 1, true>::run(y, fm & HALF, ul);
#pragma unroll
    for (int j = 0; j < H; ++j) y[j] = g_minsum(x[j], x[j + H], (bl << (31 - j)) & 0x80000000u);
    uint32_t br = SubTree<T - 1, false>::run(y, fm >> H, ur);
    u = ul | (ur << H);
    return (bl ^ br) | (br << H);
  }
};
template <bool CLIPPED>
struct SubTree<0, CLIPPED> {
  PHD static uint32_t run(const float (&x)[1], uint32_t fm, uint32_t &u) {
    // polar_sc.py:90-98: frozen -> 0; else u = [llr <= 0] (exact 0 -> 1)
    uint32_t bit = ((fm & 1u) == 0u && x[0] <= 0.0f) ? 1u : 0u;
    u = bit; return bit;
  }
};

PHD int ilog2(int n) { int m = 0; while ((1 << m) < n) ++m; return m; }

}  // namespace polar
