// polar_common.cuh -- shared device/host helpers for the sm_100a polar kernels.
// Semantics follow SURVEY.md Appendix A (validated restatement of the reference decoders).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

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
#if defined(POLAR_F_BOXPLUS)
// Exact boxplus f of the Sionna-style decoders (my_sn/fec/polar/dec.py:33-46), fp32 like the
// reference's torch ops: clip both inputs to +-30, ln(1+e^(x+y)) - ln(e^x + e^y).  Selected per translation unit by
// polar_bp_wrap.cu, which compiles the SC kernels a second time into namespace polar_bp.  expf/logf are the
// accurate CUDA versions (<= 2 ulp); the CPU reference uses its own libm, so parity for this mode is statistical
// (decisions differ only where the boxplus difference cancels to rounding noise).
PHD float f_minsum(float a, float b) {
  const float x = fminf(fmaxf(a, -kLlrMax), kLlrMax), y = fminf(fmaxf(b, -kLlrMax), kLlrMax);
  return logf(1.0f + expf(x + y)) - logf(expf(x) + expf(y));
}
PHD float f_minsum_noclip(float a, float b) { return f_minsum(a, b); }
#else
// One instruction per min on sm_100a: `min.xorsign.abs.f32 d, a, b` = (sign(a) xor sign(b)) . min(|a|, |b|) (SASS
// FMNMX.XORSIGN).  Clipping is the same instruction against +30 (positive: the sign stays a's), so f is two dependent
// ALU-pipe instructions and the unclipped form ONE -- round 1 used FMNMX3 + FMUL (sign via the product) + LOP3.
#if defined(__CUDA_ARCH__)
PDEV float xorsign_min(float a, float b) {
  float d;
  asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
#endif
PHD float f_minsum(float a, float b) {
#if defined(__CUDA_ARCH__)
  return xorsign_min(xorsign_min(a, kLlrMax), b);
#else
  float mag = fminf(fminf(fabsf(a), fabsf(b)), kLlrMax);
  uint32_t sgn = (f2u(a) ^ f2u(b)) & 0x80000000u;
  return u2f(f2u(mag) | sgn);
#endif
}
// same, for inputs already known to lie in [-30, 30] (outputs of f): the clip is the identity.
PHD float f_minsum_noclip(float a, float b) {
#if defined(__CUDA_ARCH__)
  return xorsign_min(a, b);
#else
  float mag = fminf(fabsf(a), fabsf(b));
  uint32_t sgn = (f2u(a) ^ f2u(b)) & 0x80000000u;
  return u2f(f2u(mag) | sgn);
#endif
}
#endif
// f on LOGITS (LLR = -logit, polar_sc.py:122): for min-sum the negation cancels exactly (sign.sign, |.|); the
// boxplus variant negates explicitly so that its exp/log see the same arguments as the reference's.
#if defined(POLAR_F_BOXPLUS)
PHD float f_minsum_neg(float a, float b) { return f_minsum(-a, -b); }
#else
PHD float f_minsum_neg(float a, float b) { return f_minsum(a, b); }
#endif
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
//  left child rate-0: its f values are never used; the right child's LLRs are g(a,b,0) = a + b
//   (a REP node -- all frozen but the last leaf -- thus reduces to the reference's own add tree).
template <int T, bool CLIPPED = false>
struct SubTree {
  static constexpr int N = 1 << T, H = N / 2;
  static constexpr uint32_t FULL = (N == 32) ? 0xFFFFFFFFu : ((1u << N) - 1u);
  static constexpr uint32_t HALF = (1u << H) - 1u;
  PHD static uint32_t run(const float (&x)[N], uint32_t fm, uint32_t &u) {
    fm &= FULL;
    if (fm == FULL) { u = 0; return 0; }
    if (fm == 0) {
      uint32_t hd = 0; bool zero = false;
#pragma unroll
      for (int j = 0; j < N; ++j) { hd |= (f2u(x[j]) >> 31) << j; zero |= (x[j] == 0.0f); }
      if (!zero) { u = ptransform<T>(hd); return hd; }
    }
    float y[H];
    uint32_t ul = 0, ur, bl = 0;
    if ((fm & HALF) != HALF) {
#pragma unroll
      for (int j = 0; j < H; ++j) y[j] = CLIPPED ? f_minsum_noclip(x[j], x[j + H]) : f_minsum(x[j], x[j + H]);
      bl = SubTree<T - 1, true>::run(y, fm & HALF, ul);
#pragma unroll
      for (int j = 0; j < H; ++j) y[j] = g_minsum(x[j], x[j + H], (bl << (31 - j)) & 0x80000000u);
    } else {
      // left child is rate-0: its decisions and partial sums are 0, so g = (1-0).a + b
#pragma unroll
      for (int j = 0; j < H; ++j) y[j] = x[j] + x[j + H];
    }
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


// Rate-1 node: beta = hard decisions (sign bits) of x, u = T(beta).  Exact unless some x is exactly 0
// (returns false then; the caller falls back to the recursion).  Balanced trees instead of chains: the
// caller is a single lane on the latency-critical path.
template <int T>
PDEV bool rate1_decide(const float (&x)[1 << T], uint32_t &beta, uint32_t &u) {
  constexpr int N = 1 << T, G = N < 8 ? N : 8;   // groups of 8 bound the number of live temporaries
  uint32_t hd = 0;
  float mn = 1.0f;
#pragma unroll
  for (int g0 = 0; g0 < N; g0 += G) {
    uint32_t t[G];
    float a[G];
#pragma unroll
    for (int j = 0; j < G; ++j) { t[j] = (f2u(x[g0 + j]) >> 31) << (g0 + j); a[j] = fabsf(x[g0 + j]); }
#pragma unroll
    for (int st = G / 2; st >= 1; st >>= 1) {
#pragma unroll
      for (int j = 0; j < st; ++j) { t[j] |= t[j + st]; a[j] = fminf(a[j], a[j + st]); }
    }
    hd |= t[0]; mn = fminf(mn, a[0]);
  }
  beta = hd; u = ptransform<T>(hd);
  return mn != 0.0f;
}

// Same contract as SubTree<T>::run, but every level T >= 3 is a ROLLED two-iteration loop over its
// children (h = 0: f then left child, h = 1: g then right child) with the child inlined once.  Stage s
// of the subtree always lives in the same registers (one live node per stage), so indices stay
// compile-time while the code is ~0.5 K instructions instead of ~2 K for the fully unrolled recursion --
// small enough to stay in the instruction cache when several CTAs run different phases on one SM.
template <int T>
struct RollTree {
  static constexpr int N = 1 << T, H = N / 2;
  static constexpr uint32_t FULL = (N == 32) ? 0xFFFFFFFFu : ((1u << N) - 1u);
  static constexpr uint32_t HALF = (1u << H) - 1u;
  PDEV static uint32_t run(const float (&x)[N], uint32_t fm, uint32_t &u) {
    fm &= FULL;
    if (fm == FULL) { u = 0; return 0; }
    if (fm == 0) {
      uint32_t b;
      if (rate1_decide<T>(x, b, u)) return b;
    }
    float y[H];
    uint32_t bl = 0, ul = 0, bc = 0, uc = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const uint32_t fmc = (h == 0) ? (fm & HALF) : (fm >> H);
      if (fmc == HALF) { bc = 0; uc = 0; continue; }      // rate-0 child: decisions and partial sums are 0
      if (h == 0) {
#pragma unroll
        for (int j = 0; j < H; ++j) y[j] = f_minsum(x[j], x[j + H]);
      } else {
#pragma unroll
        for (int j = 0; j < H; ++j) y[j] = g_minsum(x[j], x[j + H], (bl << (31 - j)) & 0x80000000u);
      }
      bc = RollTree<T - 1>::run(y, fmc, uc);
      if (h == 0) { bl = bc; ul = uc; }
    }
    u = ul | (uc << H);
    return (bl ^ bc) | (bc << H);
  }
};
// 4-leaf subtree, branch-free.  g is evaluated speculatively: a+b (u=0) and b-a (u=1) are formed while the
// decision they depend on is still in flight, then selected -- (1-2u).a + b is exactly one of the two
// (polar_sc.py:49-53), so this is bit-identical while the dependent chain per leaf drops from
// ~7 to ~3 instructions.  Frozen leaves are forced to 0 through the (warp-uniform) mask bits.
PDEV uint32_t leaf4(const float (&x)[4], uint32_t fm, uint32_t &u) {
  const bool n0 = !(fm & 1u), n1 = !(fm & 2u), n2 = !(fm & 4u), n3 = !(fm & 8u);
  const float a0 = f_minsum(x[0], x[2]), a1 = f_minsum(x[1], x[3]);
  const float s0 = x[0] + x[2], d0 = x[2] - x[0], s1 = x[1] + x[3], d1 = x[3] - x[1];
  const float l0 = f_minsum_noclip(a0, a1);
  const float sa = a0 + a1, da = a1 - a0;
  const bool u0 = n0 & (l0 <= 0.0f);
  const float l1 = u0 ? da : sa;
  const bool u1 = n1 & (l1 <= 0.0f);
  const bool p0 = u0 != u1;                      // partial sums of the left pair: (u0^u1, u1)
  const float b0 = p0 ? d0 : s0, b1 = u1 ? d1 : s1;
  const float l2 = f_minsum_noclip(b0, b1);      // a leaf LLR is only compared with 0: the +-30 clip cannot change that
  const float sb = b0 + b1, db = b1 - b0;
  const bool u2 = n2 & (l2 <= 0.0f);
  const float l3 = u2 ? db : sb;
  const bool u3 = n3 & (l3 <= 0.0f);
  const bool p2 = u2 != u3;
  u = (uint32_t)u0 | ((uint32_t)u1 << 1) | ((uint32_t)u2 << 2) | ((uint32_t)u3 << 3);
  return (uint32_t)(p0 != p2) | ((uint32_t)(u1 != u3) << 1) | ((uint32_t)p2 << 2) | ((uint32_t)u3 << 3);
}
template <>
struct RollTree<2> {
  PDEV static uint32_t run(const float (&x)[4], uint32_t fm, uint32_t &u) {
    fm &= 0xFu;
    if (fm == 0xFu) { u = 0; return 0; }
    return leaf4(x, fm, u);
  }
};
// 8-leaf subtree: two leaf4 with the same speculative g between them; branch-free apart from the rate-0 skip
// (the uniform tests cost more on this serial path than the arithmetic they would save).
template <>
struct RollTree<3> {
  PDEV static uint32_t run(const float (&x)[8], uint32_t fm, uint32_t &u) {
    fm &= 0xFFu;
    if (fm == 0xFFu) { u = 0; return 0; }
    float y[4], sd[4], df[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { y[j] = f_minsum(x[j], x[j + 4]); sd[j] = x[j] + x[j + 4]; df[j] = x[j + 4] - x[j]; }
    uint32_t ul, ur;
    const uint32_t bl = leaf4(y, fm & 0xFu, ul);
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = ((bl >> j) & 1u) ? df[j] : sd[j];
    const uint32_t br = leaf4(y, fm >> 4, ur);
    u = ul | (ur << 4);
    return (bl ^ br) | (br << 4);
  }
};

// ---- partial-sum-only subtrees (polar_sc3.cu) -------------------------------------------------------
// The decisions of a node are the polar transform of its partial sums (u = T(beta), T an involution), so
// the serial per-lane path only has to produce beta; u is recovered once per codeword at the end.
// CL: the node's LLRs are outputs of an f (a left child), i.e. already inside [-30, 30] -- the clip of the first f level is
// then the identity and costs nothing (one FMNMX.XORSIGN instead of two, on the serial path).
template <bool CL = false>
PDEV uint32_t leaf4_beta(const float (&x)[4], uint32_t fm) {
  const bool n0 = !(fm & 1u), n1 = !(fm & 2u), n2 = !(fm & 4u), n3 = !(fm & 8u);
  const float a0 = CL ? f_minsum_noclip(x[0], x[2]) : f_minsum(x[0], x[2]);
  const float a1 = CL ? f_minsum_noclip(x[1], x[3]) : f_minsum(x[1], x[3]);
  const float s0 = x[0] + x[2], d0 = x[2] - x[0], s1 = x[1] + x[3], d1 = x[3] - x[1];
  const float l0 = f_minsum_noclip(a0, a1);
  const float sa = a0 + a1, da = a1 - a0;
  const bool u0 = n0 & (l0 <= 0.0f);
  const float l1 = u0 ? da : sa;
  const bool u1 = n1 & (l1 <= 0.0f);
  const bool p0 = u0 != u1;                      // partial sums of the left pair: (u0^u1, u1)
  const float b0 = p0 ? d0 : s0, b1 = u1 ? d1 : s1;
  const float l2 = f_minsum_noclip(b0, b1);      // a leaf LLR is only compared with 0: the +-30 clip cannot change that
  const float sb = b0 + b1, db = b1 - b0;
  const bool u2 = n2 & (l2 <= 0.0f);
  const float l3 = u2 ? db : sb;
  const bool u3 = n3 & (l3 <= 0.0f);
  const bool p2 = u2 != u3;
  return (uint32_t)(p0 != p2) | ((uint32_t)(u1 != u3) << 1) | ((uint32_t)p2 << 2) | ((uint32_t)u3 << 3);
}
template <int T>
PDEV bool rate1_beta(const float (&x)[1 << T], uint32_t &beta) {
  uint32_t u;
  return rate1_decide<T>(x, beta, u);      // the transform of u is dead code here and is removed
}
// CL: the node's LLRs are outputs of an f (the node is a left child), already inside [-30, 30]: the first f level below it
// needs no clip.  Known at compile time wherever the two children of a node are separate calls (16 leaves and below);
// the rolled levels above pass CL = false.  (A warp-uniform run-time flag for those levels was measured: the extra uniform
// branches on the serial path cost more than the 8 .. 64 FMNMX they save, 2.16 -> 2.27 ms at n = 1024.)
template <int T, bool CL = false>
struct BetaTree;
template <bool CL>
struct BetaTree<3, CL> {
  PDEV static uint32_t run(const float (&x)[8], uint32_t fm) {
    fm &= 0xFFu;
    if (fm == 0xFFu) return 0;
    float y[4], sd[4], df[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      y[j] = CL ? f_minsum_noclip(x[j], x[j + 4]) : f_minsum(x[j], x[j + 4]);
      sd[j] = x[j] + x[j + 4]; df[j] = x[j + 4] - x[j];
    }
    const uint32_t bl = leaf4_beta<true>(y, fm & 0xFu);
#pragma unroll
    for (int j = 0; j < 4; ++j) y[j] = ((bl >> j) & 1u) ? df[j] : sd[j];
    const uint32_t br = leaf4_beta<false>(y, fm >> 4);
    return (bl ^ br) | (br << 4);
  }
};
// 16 leaves: the two children are separate calls, so each knows whether its LLRs come from an f (left) or a g (right)
template <bool CL>
struct BetaTree<4, CL> {
  PDEV static uint32_t run(const float (&x)[16], uint32_t fm) {
    fm &= 0xFFFFu;
    if (fm == 0xFFFFu) return 0;
    if (fm == 0) {
      uint32_t b;
      if (rate1_beta<4>(x, b)) return b;
    }
    float y[8];
    uint32_t bl = 0, br = 0;
    if ((fm & 0xFFu) != 0xFFu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = CL ? f_minsum_noclip(x[j], x[j + 8]) : f_minsum(x[j], x[j + 8]);
      bl = BetaTree<3, true>::run(y, fm & 0xFFu);
    }
    if ((fm >> 8) != 0xFFu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = g_minsum(x[j], x[j + 8], (bl << (31 - j)) & 0x80000000u);
      br = BetaTree<3, false>::run(y, fm >> 8);
    }
    return (bl ^ br) | (br << 8);
  }
};
#if defined(POLAR_BT5_UNROLL)
template <bool CL>
struct BetaTree<5, CL> {
  PDEV static uint32_t run(const float (&x)[32], uint32_t fm) {
    if (fm == 0xFFFFFFFFu) return 0;
    if (fm == 0) {
      uint32_t b;
      if (rate1_beta<5>(x, b)) return b;
    }
    float y[16];
    uint32_t bl = 0, br = 0;
    if ((fm & 0xFFFFu) != 0xFFFFu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = CL ? f_minsum_noclip(x[j], x[j + 16]) : f_minsum(x[j], x[j + 16]);
      bl = BetaTree<4, true>::run(y, fm & 0xFFFFu);
    }
    if ((fm >> 16) != 0xFFFFu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = g_minsum(x[j], x[j + 16], (bl << (31 - j)) & 0x80000000u);
      br = BetaTree<4, false>::run(y, fm >> 16);
    }
    return (bl ^ br) | (br << 16);
  }
};
#endif
template <int T, bool CL>
struct BetaTree {   // rolled two-iteration loop per level, like RollTree
  static constexpr int N = 1 << T, H = N / 2;
  static constexpr uint32_t FULL = (N == 32) ? 0xFFFFFFFFu : ((1u << N) - 1u);
  static constexpr uint32_t HALF = (1u << H) - 1u;
  PDEV static uint32_t run(const float (&x)[N], uint32_t fm) {
    fm &= FULL;
    if (fm == FULL) return 0;
    if (fm == 0) {
      uint32_t b;
      if (rate1_beta<T>(x, b)) return b;
    }
    float y[H];
    uint32_t bl = 0, bc = 0;
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const uint32_t fmc = (h == 0) ? (fm & HALF) : (fm >> H);
      if (fmc == HALF) { bc = 0; continue; }
      if (h == 0) {
#pragma unroll
        for (int j = 0; j < H; ++j) y[j] = CL ? f_minsum_noclip(x[j], x[j + H]) : f_minsum(x[j], x[j + H]);
      } else {
#pragma unroll
        for (int j = 0; j < H; ++j) y[j] = g_minsum(x[j], x[j + H], (bl << (31 - j)) & 0x80000000u);
      }
      bc = BetaTree<T - 1>::run(y, fmc);
      if (h == 0) bl = bc;
    }
    return (bl ^ bc) | (bc << H);
  }
};

PHD int ilog2(int n) { int m = 0; while ((1 << m) < n) ++m; return m; }

}  // namespace polar
