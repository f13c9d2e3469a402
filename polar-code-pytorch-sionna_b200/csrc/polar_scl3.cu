// polar_scl3.cu -- SCL decoder, third mapping (default for n >= 64, L >= 2): lane = (codeword, path) like
// scl2_kernel (polar_scl.cu), but with a compile-time tree and the LLR tree kept on chip.
//
// Same semantics as polar_scl.cu (x_run_sn_polar/polar/polar_scl.py:49-234, SURVEY.md Appendix A, L-survivor
// formulation; fp64 min-sum f with +-30 clip, g = (1-2u)a+b, exact softplus path metric, stable (pm, index)
// ranking, optional CRC-aided selection my_sn/fec/polar/dec.py:507-527) and the same lazy-copy bookkeeping
// (per-stage slot pointers packed 5 bits per stage, permuted by shuffles on a fork; nothing is copied).
// What changed, and why (ncu of scl2_kernel<8>, n = 1024: 65 % of the warp time was long-scoreboard stalls in
// the tree loops on a 1.2 GB workspace, 42 % of the instructions were the info-leaf code):
//   * The two top stages are VIRTUAL: stage m-3 is computed straight from the channel row (8 logits -> 4+2+1
//     f/g), eight passes per codeword.  f of fp32 inputs is exact in fp32, so the first half of the passes runs
//     mostly in fp32; only g converts to fp64.  The per-path LLR state drops from 8 KB to 2 KB (n = 1024) and the
//     channel row -- shared by the L paths of a codeword -- is the only thing read from global memory.
//   * Stages < SS live in shared memory ([element][32 lanes] doubles), the (at most two) stages between SS and
//     m-3 in a small L2-resident workspace; every stage has its own compile-time code path (LDS/LDG, unrolled).
//   * Leaf: log(1+exp(.)) stays literal (ties between path metrics are decided by their last bit, see
//     polar_softplus.cuh) but runs the math library's exp / log operation sequences without their out-of-domain
//     branches and with the coefficients in constant memory: ~65 instead of 163 instructions per penalty.
// CTAs of 4 independent warps (16 warps per SM); the warps of a CTA meet at a barrier every few leaves so that they run
// the same code at the same time: the kernel is ~70 KB of mostly straight-line code, and 16-32 unsynchronised
// warps missed the SM instruction cache 10-22 % of the time (ncu: GPC instruction-fetch path 76-86 % busy).
#include <math.h>
#include <type_traits>

#include "polar_internal.h"
#include "polar_warp.cuh"
#include "polar_softplus.cuh"

namespace polar {

#ifndef POLAR_SCL3_RANK
#define POLAR_SCL3_RANK 1
#endif
#ifndef POLAR_SCL3_VUNROLL
#define POLAR_SCL3_VUNROLL 1
#endif
#ifndef POLAR_SCL3_DISCARD
#define POLAR_SCL3_DISCARD 1
#endif
#ifndef POLAR_SCL3_RANK_MAXL
#define POLAR_SCL3_RANK_MAXL 16
#endif

namespace scl3 {

constexpr int kVUnroll = POLAR_SCL3_VUNROLL;   // channel-load groups in flight in the virtual passes
constexpr double kLlrMaxD = 30.0;
constexpr unsigned FULL = 0xFFFFFFFFu;

struct Params {
  const float *logit; const uint32_t *fmask; int64_t B;
  uint32_t *best; float *u_info; const int32_t *info_pos; int k;
  double *pm_out; uint32_t *list; const uint32_t *crc_rows; int crc_len;
  unsigned char *ws; size_t ws_bytes_per_warp; int sync_mask;
};

// ---- f / g on fp32 (exact for f) and fp64 ----------------------------------------------------------
#if defined(POLAR_F_BOXPLUS)
// exact boxplus of the Sionna-style list decoder (my_sn/fec/polar/dec.py:331-340, numpy float64): clip to +-30,
// ln(1+e^(x+y)) - ln(e^x+e^y), always in fp64 (polar_bp_wrap.cu compiles this unit a second time into namespace polar_bp)
PDEV double fop(double a, double b) {
  const double x = fmax(fmin(a, kLlrMaxD), -kLlrMaxD), y = fmax(fmin(b, kLlrMaxD), -kLlrMaxD);
  return log(1.0 + exp(x + y)) - log(exp(x) + exp(y));
}
PDEV double fop(float a, float b) { return fop((double)a, (double)b); }
PDEV double fop_nc(double a, double b) { return fop(a, b); }
constexpr bool kEvenLeafInside = false;     // the boxplus of two clipped values stays inside the clip, but only mathematically
#else
PDEV float fop(float a, float b) {          // polar_scl.py:93-106 on fp32-representable values: exact
  const float mag = fminf(fminf(fabsf(a), fabsf(b)), 30.0f);
  return __uint_as_float(__float_as_uint(mag) | ((__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u));
}
PDEV double fop(double a, double b) {
  const double aa = fabs(a), ab = fabs(b);
  double mag = (aa < ab) ? aa : ab;
  mag = (mag < kLlrMaxD) ? mag : kLlrMaxD;
  const int hi = (__double2hiint(a) ^ __double2hiint(b)) & 0x80000000;
  return __hiloint2double(__double2hiint(mag) | hi, __double2loint(mag));
}
// same for inputs that are outputs of an f themselves (inside [-30, 30]): the clip is the identity
PDEV double fop_nc(double a, double b) {
  const double aa = fabs(a), ab = fabs(b);
  const double mag = (aa < ab) ? aa : ab;
  const int hi = (__double2hiint(a) ^ __double2hiint(b)) & 0x80000000;
  return __hiloint2double(__double2hiint(mag) | hi, __double2loint(mag));
}
constexpr bool kEvenLeafInside = true;      // an even leaf's LLR is the output of an f: polar_scl.py:81's clip cannot change it
#endif
PDEV double gop(double a, double b, unsigned u) {   // polar_scl.py:107-108; u in {0,1}
  const double sa = __hiloint2double(__double2hiint(a) ^ (int)(u << 31), __double2loint(a));
  return __dadd_rn(sa, b);
}
PDEV double gop(float a, float b, unsigned u) { return gop((double)a, (double)b, u); }
template <bool G, class T>
PDEV auto op(T a, T b, unsigned u) {
  if constexpr (G) return gop(a, b, u);
  else return fop(a, b);
}

template <int L> struct Log2 { static constexpr int v = 1 + Log2<L / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

// M = log2 n (6..12), L = list size (2..32), SS = number of LLR stages held in shared memory.
template <int M, int L, int SS_>
struct Cfg {
  static constexpr int N = 1 << M, NW = N / 32;
  static constexpr int TOP = M - 3;                       // highest stored LLR stage
  static constexpr int SS = (SS_ > TOP + 1) ? TOP + 1 : SS_;
  static constexpr int HT = 1 << TOP;                     // elements per pass
  static constexpr int CPW = 32 / L;
  static constexpr int LOGL = Log2<L>::v;
  static constexpr size_t llr_smem_bytes = (size_t)32 * 8 * ((1u << SS) - 1u);
  // ranking scratch (L <= 8): candidate keys [group][L+1] x 16 B (one pad entry per group: conflict-free broadcast
  // reads), then the survivors [group][L+1] x 16 B
  static constexpr bool RANK = (L <= POLAR_SCL3_RANK_MAXL) && (POLAR_SCL3_RANK != 0);
  static constexpr int GSTRIDE = L + 1;                        // 16-byte entries per lane group
  static constexpr size_t rank_bytes = RANK ? (size_t)2 * CPW * GSTRIDE * 16 : 0;
  static constexpr size_t smem_bytes = llr_smem_bytes + rank_bytes;
  static constexpr size_t gl_doubles = (size_t)32 * ((1u << (TOP + 1)) - (1u << SS));
  static constexpr size_t word_count = (size_t)32 * 2 * NW;
  static constexpr size_t ws_bytes = ((gl_doubles * 8 + word_count * 4 + 255) / 256) * 256;
};

// WPC warps per CTA, CPS CTAs per SM (register budget).  The warps of a CTA are independent decoders; they only meet at
// a barrier every `sync_mask + 1` leaves, which keeps them on the same few KB of code (see the header comment).
template <int M, int L, int SS_, int WPC, int CPS>
__global__ void __launch_bounds__(32 * WPC, CPS) scl3_kernel(const Params P) {
  using C = Cfg<M, L, SS_>;
  constexpr int N = C::N, NW = C::NW, TOP = C::TOP, SS = C::SS, HT = C::HT, CPW = C::CPW, LOGL = C::LOGL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *const llr_s = reinterpret_cast<double *>(smem_raw + (size_t)warp * C::smem_bytes);
  const int p = lane & (L - 1), gbase = lane & ~(L - 1), cwl = lane >> LOGL;
  unsigned char *const wsb = P.ws + ((size_t)blockIdx.x * WPC + warp) * P.ws_bytes_per_warp;
  double *const llr_g = reinterpret_cast<double *>(wsb);
  uint32_t *const bl = reinterpret_cast<uint32_t *>(wsb + C::gl_doubles * 8);   // stage s>=5 at (2^(s-5)-1)*32
  uint32_t *const rootw = bl + (size_t)32 * NW;                                  // [NW][32]

  const unsigned long long idrow = (unsigned long long)p * 0x0084210842108421ull;   // field s (5 bits) = p
  const int64_t nbatch = (P.B + CPW - 1) / CPW;

  // ---- per-batch state (lambdas below capture by reference) ----
  unsigned long long rowL = idrow, rowB = idrow;
  uint32_t small = 0u;
  const float *ch = P.logit;

  auto stage_sm = [&](auto sc) -> double * { constexpr int S = decltype(sc)::value; return llr_s + 32u * ((1u << S) - 1u); };
  auto stage_gl = [&](auto sc) -> double * { constexpr int S = decltype(sc)::value; return llr_g + (32u * ((1u << S) - 1u) - 32u * ((1u << SS) - 1u)); };

  // f step: stage S -> S-1, own slot (lane-private)
  auto fstep = [&](auto sc) {
    constexpr int S = decltype(sc)::value;
    constexpr int h = 1 << (S - 1);
    const double *src;
    double *dst;
    if constexpr (S < SS) src = stage_sm(sc) + lane; else src = stage_gl(sc) + lane;
    if constexpr (S - 1 < SS) dst = stage_sm(std::integral_constant<int, S - 1>{}) + lane;
    else dst = stage_gl(std::integral_constant<int, S - 1>{}) + lane;
    constexpr int U = h < 8 ? h : 8;
#pragma unroll 1
    for (int e0 = 0; e0 < h; e0 += U) {
      double a[U], b[U];
#pragma unroll
      for (int r = 0; r < U; ++r) { a[r] = src[(e0 + r) * 32]; b[r] = src[(e0 + r + h) * 32]; }
#pragma unroll
      for (int r = 0; r < U; ++r) dst[(e0 + r) * 32] = fop(a[r], b[r]);
    }
  };
  // g step: stage T+1 (slot of the ancestor that wrote it) -> T (own slot); beta = left sibling's partial sums
  auto gstep = [&](auto tc) {
    constexpr int T = decltype(tc)::value;
    constexpr int h = 1 << T;
    const unsigned q = gbase + (unsigned)((rowL >> (5 * (T + 1))) & 31u);
    const double *src;
    double *dst;
    if constexpr (T + 1 < SS) src = stage_sm(std::integral_constant<int, T + 1>{}) + q;
    else src = stage_gl(std::integral_constant<int, T + 1>{}) + q;
    if constexpr (T < SS) dst = stage_sm(tc) + lane; else dst = stage_gl(tc) + lane;
    if constexpr (T < 5) {
      const uint32_t ub = small >> ((1u << T) - 1u);
      double a[h], b[h];
#pragma unroll
      for (int r = 0; r < h; ++r) { a[r] = src[r * 32]; b[r] = src[(r + h) * 32]; }
#pragma unroll
      for (int r = 0; r < h; ++r) dst[r * 32] = gop(a[r], b[r], (ub >> r) & 1u);
    } else {
      const uint32_t *bw = bl + ((1u << (T - 5)) - 1u) * 32 + gbase + (unsigned)((rowB >> (5 * T)) & 31u);
#pragma unroll 1
      for (int w = 0; w < (h >> 5); ++w) {
        const uint32_t ub = bw[w * 32];
#pragma unroll 1
        for (int e0 = 0; e0 < 32; e0 += 8) {
          double a[8], b[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) { a[r] = src[(w * 32 + e0 + r) * 32]; b[r] = src[(w * 32 + e0 + r + h) * 32]; }
#pragma unroll
          for (int r = 0; r < 8; ++r) dst[(w * 32 + e0 + r) * 32] = gop(a[r], b[r], (ub >> (e0 + r)) & 1u);
        }
      }
    }
    if constexpr (T + 1 >= SS && POLAR_SCL3_DISCARD) {
      // every path has now consumed stage T+1 for the last time (it is rewritten before it is read again): drop the
      // array from the L2 instead of letting its dirty lines be written back to DRAM
      __syncwarp();
      const char *dead = reinterpret_cast<const char *>(stage_gl(std::integral_constant<int, T + 1>{}));
#pragma unroll 1
      for (int ln = lane; ln < (2 << (T + 1)); ln += 32)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead + (size_t)ln * 128) : "memory");
    }
  };
  // one pass of the virtual top: stage TOP of this path from the channel row.  Q = (i >> TOP): bit 2 / 1 / 0
  // set <=> the node at stage m-1 / m-2 / m-3 is a right child (g with the left sibling's partial sums).
  auto vpass = [&](auto qc) {
    constexpr int Q = decltype(qc)::value;
    constexpr bool G9 = (Q & 4) != 0, G8 = (Q & 2) != 0, G7 = (Q & 1) != 0;
    if constexpr (HT < 32) {
      // n = 64 / 128: a pass is 8 / 16 elements; the partial sums of the three stages come from the word arrays
      // (stage >= 5) or from the `small` register (stage < 5), four bits per float4 group
      auto bits4 = [&](auto sc, int e0) -> uint32_t {
        constexpr int S = decltype(sc)::value;
        if constexpr (S >= 5) {
          const uint32_t *w = bl + ((1u << (S - 5)) - 1u) * 32 + gbase + (unsigned)((rowB >> (5 * S)) & 31u);
          return (w[(e0 >> 5) * 32] >> (e0 & 31)) & 0xFu;
        } else {
          return (small >> (((1u << S) - 1u) + (unsigned)e0)) & 0xFu;
        }
      };
      double *dst = stage_sm(std::integral_constant<int, TOP>{}) + lane;
#pragma unroll
      for (int e = 0; e < HT; e += 4) {
        float4 c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = __ldg(reinterpret_cast<const float4 *>(ch + e + HT * j));
        uint32_t u9[4] = {0u, 0u, 0u, 0u}, u8[2] = {0u, 0u}, u7 = 0u;
        if constexpr (G9) {
#pragma unroll
          for (int j = 0; j < 4; ++j) u9[j] = bits4(std::integral_constant<int, M - 1>{}, e + HT * j);
        }
        if constexpr (G8) {
#pragma unroll
          for (int j = 0; j < 2; ++j) u8[j] = bits4(std::integral_constant<int, M - 2>{}, e + HT * j);
        }
        if constexpr (G7) u7 = bits4(std::integral_constant<int, M - 3>{}, e);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {   // LLR = -logit (polar_scl.py:219)
            const float v = (r == 0) ? c[j].x : (r == 1) ? c[j].y : (r == 2) ? c[j].z : c[j].w;
            x[j] = -v;
          }
          const auto s90 = op<G9>(x[0], x[4], (u9[0] >> r) & 1u);
          const auto s91 = op<G9>(x[1], x[5], (u9[1] >> r) & 1u);
          const auto s92 = op<G9>(x[2], x[6], (u9[2] >> r) & 1u);
          const auto s93 = op<G9>(x[3], x[7], (u9[3] >> r) & 1u);
          const auto s80 = op<G8>(s90, s92, (u8[0] >> r) & 1u);
          const auto s81 = op<G8>(s91, s93, (u8[1] >> r) & 1u);
          const auto s7 = op<G7>(s80, s81, (u7 >> r) & 1u);
          dst[(e + r) * 32] = (double)s7;
        }
      }
    } else {
      constexpr int WB = HT / 32;                                     // 32-element blocks per pass
      const uint32_t *w9 = bl + ((1u << (M - 1 - 5)) - 1u) * 32 + gbase + (unsigned)((rowB >> (5 * (M - 1))) & 31u);
      const uint32_t *w8 = bl + ((1u << (M - 2 - 5)) - 1u) * 32 + gbase + (unsigned)((rowB >> (5 * (M - 2))) & 31u);
      const uint32_t *w7 = bl + ((1u << (M - 3 - 5)) - 1u) * 32 + gbase + (unsigned)((rowB >> (5 * (M - 3))) & 31u);
      double *dst;
      if constexpr (TOP < SS) dst = stage_sm(std::integral_constant<int, TOP>{}) + lane;
      else dst = stage_gl(std::integral_constant<int, TOP>{}) + lane;
  #pragma unroll 1
      for (int wb = 0; wb < WB; ++wb) {
        uint32_t u9[4] = {0u, 0u, 0u, 0u}, u8[2] = {0u, 0u}, u7 = 0u;
        if constexpr (G9) {
  #pragma unroll
          for (int j = 0; j < 4; ++j) u9[j] = w9[(wb + WB * j) * 32];
        }
        if constexpr (G8) {
  #pragma unroll
          for (int j = 0; j < 2; ++j) u8[j] = w8[(wb + WB * j) * 32];
        }
        if constexpr (G7) u7 = w7[wb * 32];
  #pragma unroll kVUnroll
        for (int e4 = 0; e4 < 8; ++e4) {
          const int e = wb * 32 + e4 * 4;
          float4 c[8];
  #pragma unroll
          for (int j = 0; j < 8; ++j) c[j] = __ldg(reinterpret_cast<const float4 *>(ch + e + HT * j));
  #pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int bp = e4 * 4 + r;
            float x[8];
  #pragma unroll
            for (int j = 0; j < 8; ++j) {   // LLR = -logit (polar_scl.py:219)
              const float v = (r == 0) ? c[j].x : (r == 1) ? c[j].y : (r == 2) ? c[j].z : c[j].w;
              x[j] = -v;
            }
            const auto s90 = op<G9>(x[0], x[4], (u9[0] >> bp) & 1u);
            const auto s91 = op<G9>(x[1], x[5], (u9[1] >> bp) & 1u);
            const auto s92 = op<G9>(x[2], x[6], (u9[2] >> bp) & 1u);
            const auto s93 = op<G9>(x[3], x[7], (u9[3] >> bp) & 1u);
            const auto s80 = op<G8>(s90, s92, (u8[0] >> bp) & 1u);
            const auto s81 = op<G8>(s91, s93, (u8[1] >> bp) & 1u);
            const auto s7 = op<G7>(s80, s81, (u7 >> bp) & 1u);
            dst[(e + r) * 32] = (double)s7;
          }
        }
      }
    }
  };
  // every warp of the CTA runs the same number of batches (the barriers below need all of them); a batch past the end
  // decodes the last codeword again and stores nothing
  const int64_t bstride = (int64_t)gridDim.x * WPC;
  const int64_t rounds = (nbatch + bstride - 1) / bstride;
  for (int64_t rd = 0; rd < rounds; ++rd) {
    const int64_t bb = rd * bstride + (int64_t)blockIdx.x * WPC + warp;
    const int64_t b = bb * CPW + cwl;
    const bool valid = bb < nbatch && b < P.B;
    ch = P.logit + (valid ? b : (P.B - 1)) * (int64_t)N;
    double pm = (p == 0) ? 0.0 : kLlrMaxD;                       // polar_scl.py:192-194
    rowL = idrow; rowB = idrow; small = 0u;
    uint32_t fword = 0u;

    // One iteration = one PAIR of leaves (i0, i0+1) under a stage-1 node.  The stage-0 LLRs never leave registers, the
    // odd leaf needs no dispatch at all, and a frozen-frozen pair (no fork possible) evaluates its two penalties as
    // independent instruction streams.
#pragma unroll 1
    for (int pr = 0; pr < N / 2; ++pr) {
      const int i0 = 2 * pr;
      if constexpr (WPC > 1) {
        if ((pr & P.sync_mask) == 0) __syncthreads();
      }
      __syncwarp();
      if ((i0 & 31) == 0) fword = __ldg(P.fmask + (i0 >> 5));
      const unsigned fz = (fword >> (i0 & 31)) & 3u;             // bit 0 / 1: leaf i0 / i0+1 frozen
      // ------------------------------------------------------------------ descent to the stage-1 node
      const int t = (i0 == 0) ? M : (__ffs(i0) - 1);             // >= 1
      double l1a, l1b, x0;                                         // x0: LLR of the even leaf = f(stage 1)
      if (t == 1) {
        // g: stage 2 (slot of the ancestor that wrote it) -> stage 1, beta = the left pair's partial sums
        const double *src = llr_s + 32 * 3 + gbase + (unsigned)((rowL >> 10) & 31u);
        const uint32_t ub = small >> 1;
        l1a = gop(src[0], src[2 * 32], ub & 1u);
        l1b = gop(src[1 * 32], src[3 * 32], (ub >> 1) & 1u);
        x0 = fop(l1a, l1b);
      } else {
        // every f / g step exists exactly once in the binary: a g (or virtual-pass) switch, then ONE fall-through f
        // cascade down to stage 2.
        if (t >= TOP) {
          switch (i0 >> TOP) {
            case 0: vpass(std::integral_constant<int, 0>{}); break;
            case 1: vpass(std::integral_constant<int, 1>{}); break;
            case 2: vpass(std::integral_constant<int, 2>{}); break;
            case 3: vpass(std::integral_constant<int, 3>{}); break;
            case 4: vpass(std::integral_constant<int, 4>{}); break;
            case 5: vpass(std::integral_constant<int, 5>{}); break;
            case 6: vpass(std::integral_constant<int, 6>{}); break;
            default: vpass(std::integral_constant<int, 7>{}); break;
          }
        } else {
#define POLAR_SCL3_G(T) case T: if constexpr (T < TOP) gstep(std::integral_constant<int, T>{}); break;
          switch (t) {
            POLAR_SCL3_G(2) POLAR_SCL3_G(3) POLAR_SCL3_G(4)
            POLAR_SCL3_G(5) POLAR_SCL3_G(6) POLAR_SCL3_G(7) POLAR_SCL3_G(8)
            default: break;
          }
#undef POLAR_SCL3_G
        }
#define POLAR_SCL3_F(S) case S: if constexpr (S <= TOP) fstep(std::integral_constant<int, S>{}); [[fallthrough]];
        switch (t < TOP ? t : TOP) {
          POLAR_SCL3_F(9) POLAR_SCL3_F(8) POLAR_SCL3_F(7) POLAR_SCL3_F(6) POLAR_SCL3_F(5)
          POLAR_SCL3_F(4) POLAR_SCL3_F(3)
          default: break;
        }
#undef POLAR_SCL3_F
        const double *src = llr_s + 32 * 3 + lane;               // stage 2, own slot
        l1a = fop(src[0], src[2 * 32]);
        l1b = fop(src[1 * 32], src[3 * 32]);
        x0 = fop_nc(l1a, l1b);                                   // both inputs are outputs of an f: nothing to clip
      }
      {  // stages 1..min(t, TOP) now belong to this path (own slot); stage 0 lives in registers only
        const int top = (t < TOP ? t : TOP);
        const unsigned long long msk = ((1ull << (5 * (top + 1))) - 1ull) & ~31ull;
        rowL = (rowL & ~msk) | (idrow & msk);
      }
      // info leaf: fork every path, rank the 2L candidates, keep L.  Candidate E = u*L + p (reference slot order
      // [u=0 paths | u=1 paths], polar_scl.py:49-68); two candidates per lane, bitonic sort inside the lane group.
      auto info_leaf = [&](double x, auto inside) -> unsigned {   // inside: x is the output of an f, i.e. inside the clip already
        const double xc = (decltype(inside)::value && kEvenLeafInside) ? x : fmax(fmin(x, kLlrMaxD), -kLlrMaxD);     // polar_scl.py:81
        double k0 = pm + sp::softplus_literal(-xc), k1 = pm + sp::softplus_literal(xc);   // u = 0 / u = 1
        int s0 = p, s1 = L + p;
        if constexpr (C::RANK) {
          // rank by counting: every candidate is compared with all 2L candidates of its codeword (independent
          // compares instead of the 3 log2(2L) dependent shuffle layers of a sorting network); ties go to the lower
          // candidate index, like the stable sort.  rank < L survives and moves to lane `rank` of the group.
          double2 *kb = reinterpret_cast<double2 *>(reinterpret_cast<unsigned char *>(llr_s) + C::llr_smem_bytes) + cwl * C::GSTRIDE;
          kb[p] = make_double2(k0, k1);
          __syncwarp();
          int r0 = 0, r1 = 0;
#pragma unroll
          for (int j = 0; j < L; ++j) {
            const double2 o = kb[j];
            const bool jl = j < p;
            r0 += (int)((o.x < k0) || (o.x == k0 && jl));
            r0 += (int)(o.y < k0);
            r1 += (int)(o.x <= k1);
            r1 += (int)((o.y < k1) || (o.y == k1 && jl));
          }
          double2 *sb = kb + CPW * C::GSTRIDE;                     // survivors of this group
          if (r0 < L) sb[r0] = make_double2(k0, __hiloint2double(0, s0));
          if (r1 < L) sb[r1] = make_double2(k1, __hiloint2double(0, s1));
          __syncwarp();
          const double2 w = sb[p];
          k0 = w.x; s0 = __double2loint(w.y);
        } else {
#pragma unroll
        for (int k = 2; k <= 2 * L; k <<= 1) {
#pragma unroll
          for (int d = k >> 1; d > 0; d >>= 1) {
            if (d == L) {   // partner is the other register; k == 2L: ascending
              const bool less10 = (k1 < k0) || (k1 == k0 && s1 < s0);
              if (less10) { const double tk = k0; k0 = k1; k1 = tk; const int ts = s0; s0 = s1; s1 = ts; }
            } else {
              const double pk0 = __shfl_xor_sync(FULL, k0, d), pk1 = __shfl_xor_sync(FULL, k1, d);
              const int ps0 = __shfl_xor_sync(FULL, s0, d), ps1 = __shfl_xor_sync(FULL, s1, d);
              const bool lower = ((p & d) == 0);
              const bool up0 = (k == 2 * L) ? true : (k == L) ? true : ((p & k) == 0);
              const bool up1 = (k == 2 * L) ? true : (k == L) ? false : ((p & k) == 0);
              const bool less0 = (pk0 < k0) || (pk0 == k0 && ps0 < s0);
              const bool less1 = (pk1 < k1) || (pk1 == k1 && ps1 < s1);
              if ((lower == up0) == less0) { k0 = pk0; s0 = ps0; }
              if ((lower == up1) == less1) { k1 = pk1; s1 = ps1; }
            }
          }
        }
        }
        const int parent = gbase + (s0 & (L - 1));
        pm = k0;
        rowL = __shfl_sync(FULL, rowL, parent);
        rowB = __shfl_sync(FULL, rowB, parent);
        small = __shfl_sync(FULL, small, parent);
        __syncwarp();   // forked paths read their parents' slots from here on
        return (unsigned)(s0 >> LOGL) & 1u;
      };
      auto frozen_pen = [&](double x, auto inside) -> double {   // polar_scl.py:82-83 with u = 0
        return sp::softplus_literal(-((decltype(inside)::value && kEvenLeafInside) ? x : fmax(fmin(x, kLlrMaxD), -kLlrMaxD)));
      };
      constexpr std::true_type kInside{};                        // the even leaf (f output)
      constexpr std::false_type kAnywhere{};                     // the odd leaf (g output)
      // ------------------------------------------------------------------ the two leaves
      uint32_t cur;                                                // partial sums of the pair: (u0 ^ u1, u1)
      if (fz == 3u) {
        const double pen0 = frozen_pen(x0, kInside), pen1 = frozen_pen(__dadd_rn(l1a, l1b), kAnywhere);   // g with u0 = 0
        pm += pen0;
        pm += pen1;
        cur = 0u;
      } else {
        double x1;
        if (fz & 1u) {
          pm += frozen_pen(x0, kInside);
          x1 = __dadd_rn(l1a, l1b);
          small &= ~1u;
        } else {
          double *st1 = llr_s + 32 * 1 + lane;                    // stage 1, own slot: the children read it after the fork
          st1[0] = l1a; st1[32] = l1b;
          const unsigned u0 = info_leaf(x0, kInside);
          small = (small & ~1u) | u0;                              // travels with the path through the next fork
          const double *q1 = llr_s + 32 * 1 + gbase + (unsigned)((rowL >> 5) & 31u);
          x1 = gop(q1[0], q1[32], u0);
        }
        unsigned u1 = 0u;
        if (fz & 2u) pm += frozen_pen(x1, kAnywhere);
        else u1 = info_leaf(x1, kAnywhere);
        cur = ((small & 1u) ^ u1) | (u1 << 1);
      }
      // ------------------------------------------------------------------ partial-sum cascade
      // z = number of completed right children above leaf i0+1  (polar_scl.py:147-153, [bl ^ br, br]); z >= 1
      const int z = (i0 + 1 == N - 1) ? M : (__ffs(~(i0 + 1)) - 1);
      const int zs = z < 5 ? z : 5;
      for (int sN = 1; sN < zs; ++sN) {
        const uint32_t w = 1u << sN;
        const uint32_t field = (small >> (w - 1u)) & ((1u << w) - 1u);
        cur = (field ^ cur) | (cur << w);
      }
      if (z < 5) {
        const uint32_t w = 1u << z, off = w - 1u, msk = ((1u << w) - 1u) << off;
        small = (small & ~msk) | (cur << off);
      } else {
        const int nwz = 1 << (z - 5);
        uint32_t *dest = ((z < M) ? (bl + (size_t)(nwz - 1) * 32) : rootw) + lane;
        dest[(nwz - 1) * 32] = cur;
        for (int sN = 5; sN < z; ++sN) {
          const int hw = 1 << (sN - 5);
          const uint32_t *bls = bl + (size_t)(hw - 1) * 32 + gbase + (unsigned)((rowB >> (5 * sN)) & 31u);
          for (int w = 0; w < hw; ++w) dest[(nwz - 2 * hw + w) * 32] = bls[w * 32] ^ dest[(nwz - hw + w) * 32];
        }
        if (z < M) rowB = (rowB & ~(31ull << (5 * z))) | ((unsigned long long)p << (5 * z));
        __syncwarp();
      }
    }  // leaf pairs

    // ---------------------------------------------------------------------- epilogue
    // root partial sums = codeword estimate x_hat; u_hat = T(x_hat) (involution); lane-private words
    for (int w = 0; w < NW; ++w) rootw[w * 32 + lane] = ptransform_rt(rootw[w * 32 + lane], 5);
    for (int d = 1; d < NW; d <<= 1)
      for (int w = 0; w < NW; ++w)
        if (!(w & d)) rootw[w * 32 + lane] ^= rootw[(w + d) * 32 + lane];
    __syncwarp();
    // final sort by path metric inside each codeword group (polar_scl.py:204)
    double key = pm;
    int src = p;
#pragma unroll
    for (int k = 2; k <= L; k <<= 1) {
#pragma unroll
      for (int d = k >> 1; d > 0; d >>= 1) {
        const double pk = __shfl_xor_sync(FULL, key, d);
        const int ps = __shfl_xor_sync(FULL, src, d);
        const bool take_min = (((p & d) == 0) == ((p & k) == 0 || k == L));
        const bool partner_less = (pk < key) || (pk == key && ps < src);
        if (take_min == partner_less) { key = pk; src = ps; }
      }
    }
    // lane p of each group now holds rank p: (pm ascending, source path)
    if (P.pm_out && valid) P.pm_out[b * L + p] = key;
    const int slot = gbase + src;                                  // slot holding the decisions of rank p
    if (P.list && valid)
      for (int w = 0; w < NW; ++w) P.list[((size_t)b * L + p) * NW + w] = rootw[w * 32 + slot];
    // CRC-aided selection (my_sn/fec/polar/dec.py:507-520): pm += 30*k for candidates failing the CRC
    double pen = key;
    if (P.crc_len > 0 && P.crc_rows) {
      uint32_t syn = 0u;
      for (int w = 0; w < NW; ++w) {
        uint32_t uw = rootw[w * 32 + slot];
        while (uw) {
          const int bpos = __ffs(uw) - 1;
          uw &= uw - 1u;
          syn ^= __ldg(P.crc_rows + w * 32 + bpos);
        }
      }
      if (syn != 0u) pen = key + kLlrMaxD * (double)P.k;
    }
    // argmin over the L sorted candidates, first minimum wins (np.argmin, dec.py:520)
    double bk = pen;
    int bi = p;
#pragma unroll
    for (int d = L >> 1; d > 0; d >>= 1) {
      const double ok = __shfl_xor_sync(FULL, bk, d);
      const int oi = __shfl_xor_sync(FULL, bi, d);
      if (ok < bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
    }
    const int best_slot = gbase + __shfl_sync(FULL, src, gbase + bi);
    if (valid) {
      if (P.best)
        for (int w = p; w < NW; w += L) P.best[(size_t)b * NW + w] = rootw[w * 32 + best_slot];
      if (P.u_info) {
        float *row = P.u_info + b * (int64_t)P.k;
        for (int tt = p; tt < P.k; tt += L) {
          const int pos = __ldg(P.info_pos + tt);
          row[tt] = (float)((rootw[(pos >> 5) * 32 + best_slot] >> (pos & 31)) & 1u);
        }
      }
    }
    __syncwarp();
  }
}

template <int M, int L, int SS_, int WPC, int CPS>
static int launch_one(const Params &P0, int64_t cta_cap_per_sm, cudaStream_t st, Scl3Plan *plan_only, int64_t B) {
  using C = Cfg<M, L, SS_>;
  constexpr size_t smem_cta = C::smem_bytes * WPC;
  int ctas_per_sm = (int)((size_t)(227 * 1024) / (smem_cta + 1024));
  if (ctas_per_sm > CPS) ctas_per_sm = CPS;
  if (cta_cap_per_sm > 0 && ctas_per_sm > cta_cap_per_sm) ctas_per_sm = (int)cta_cap_per_sm;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int64_t nbatch = (B + C::CPW - 1) / C::CPW;
  int64_t grid = (int64_t)device_sm_count() * ctas_per_sm;
  if (grid > (nbatch + WPC - 1) / WPC) grid = (nbatch + WPC - 1) / WPC;
  if (grid < 1) grid = 1;
  if (plan_only) {
    plan_only->grid = grid * WPC;                 // in warps
    plan_only->ws_bytes_per_warp = C::ws_bytes;
    return POLAR_OK;
  }
  Params P = P0;
  P.ws_bytes_per_warp = C::ws_bytes;
  auto kern = scl3_kernel<M, L, SS_, WPC, CPS>;
  if (smem_cta > (size_t)device_max_smem_optin()) return set_error(POLAR_ENOMEM, "scl3: needs %zu B shared memory per CTA", smem_cta);
  POLAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
  kern<<<(unsigned)grid, 32 * WPC, smem_cta, st>>>(P);
  count_launch();
  POLAR_CHECK_LAUNCH("scl3_kernel");
  return POLAR_OK;
}

// bitwise comparison of exp_nb / log_nb / softplus_literal with the CUDA math library on the domain the decoder uses
__global__ void math_selftest_kernel(uint64_t count, unsigned long long *mismatch) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long bad_exp = 0, bad_log = 0, bad_sp = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    uint64_t h = i * 0x9E3779B97F4A7C15ull + 0x1234567ull;     // splitmix64
    h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull; h = (h ^ (h >> 27)) * 0x94D049BB133111EBull; h ^= h >> 31;
    double v = ((double)(h >> 11) * (1.0 / 9007199254740992.0)) * 60.0 - 30.0;   // [-30, 30)
    if ((i & 15) == 0) v = (i & 16) ? 30.0 : -30.0;            // the clip values themselves
    if ((i & 15) == 1) v = (double)(float)v;                   // fp32-representable LLRs
    const double e_ref = exp(v), e_got = sp::exp_nb(v);
    bad_exp += (__double_as_longlong(e_ref) != __double_as_longlong(e_got));
    const double w = 1.0 + e_ref;
    bad_log += (__double_as_longlong(log(w)) != __double_as_longlong(sp::log_nb(w)));
    bad_sp += (__double_as_longlong(log(1.0 + exp(v))) != __double_as_longlong(sp::softplus_literal(v)));
  }
  if (bad_exp) atomicAdd(mismatch + 0, bad_exp);
  if (bad_log) atomicAdd(mismatch + 1, bad_log);
  if (bad_sp) atomicAdd(mismatch + 2, bad_sp);
}

}  // namespace scl3

bool scl3_supported(int n, int L) {
  const int m = ilog2(n);
#if defined(POLAR_SCL3_DEV)
  return (m == 10 && L == 8) || ((m == 6 || m == 7) && L == 4);
#else
  return m >= 6 && m <= 12 && L >= 2 && L <= 32;
#endif
}

// plan_only != nullptr: fill the plan (grid, workspace bytes per warp) and return; else launch.
int launch_scl3(const float *logit, const uint32_t *fmask, int n, int L, int64_t B, uint32_t *best, float *u_info,
                const int32_t *info_pos, int k, double *pm_out, uint32_t *list, const uint32_t *crc_rows, int crc_len,
                void *ws, cudaStream_t st, Scl3Plan *plan_only) {
  scl3::Params P;
  P.logit = logit; P.fmask = fmask; P.B = B; P.best = best; P.u_info = u_info; P.info_pos = info_pos; P.k = k;
  P.pm_out = pm_out; P.list = list; P.crc_rows = crc_rows; P.crc_len = crc_len;
  P.ws = (unsigned char *)ws; P.ws_bytes_per_warp = 0;
  const int m = ilog2(n);
  const int ss = env_int("POLAR_SCL3_SS", 54);
  P.sync_mask = env_int("POLAR_SCL3_SYNC", 4) - 1;
  const int64_t cap = env_int("POLAR_SCL3_CTAS", 0);
#if defined(POLAR_SCL3_DEV)
#define POLAR_SCL3_L(MM, LL)                                                                          \
  if (L == LL) {                                                                                       \
    if (ss == 51) return scl3::launch_one<MM, LL, 5, 1, 16>(P, cap, st, plan_only, B);                 \
    if (ss == 54) return scl3::launch_one<MM, LL, 5, 4, 4>(P, cap, st, plan_only, B);                  \
    if (ss == 58) return scl3::launch_one<MM, LL, 5, 8, 2>(P, cap, st, plan_only, B);                  \
    if (ss == 516) return scl3::launch_one<MM, LL, 5, 16, 1>(P, cap, st, plan_only, B);                \
    if (ss == 583) return scl3::launch_one<MM, LL, 5, 8, 3>(P, cap, st, plan_only, B);                 \
    if (ss == 5122) return scl3::launch_one<MM, LL, 5, 12, 2>(P, cap, st, plan_only, B);               \
    if (ss == 574) return scl3::launch_one<MM, LL, 5, 7, 4>(P, cap, st, plan_only, B);                 \
    if (ss == 643) return scl3::launch_one<MM, LL, 6, 4, 3>(P, cap, st, plan_only, B);                 \
    if (ss == 652) return scl3::launch_one<MM, LL, 6, 5, 2>(P, cap, st, plan_only, B);                 \
    if (ss == 642) return scl3::launch_one<MM, LL, 6, 4, 2>(P, cap, st, plan_only, B);                 \
    if (ss == 543) return scl3::launch_one<MM, LL, 5, 4, 3>(P, cap, st, plan_only, B);                 \
    if (ss == 545) return scl3::launch_one<MM, LL, 5, 4, 5>(P, cap, st, plan_only, B);                 \
    if (ss == 546) return scl3::launch_one<MM, LL, 5, 4, 6>(P, cap, st, plan_only, B);                 \
    if (ss == 547) return scl3::launch_one<MM, LL, 5, 4, 7>(P, cap, st, plan_only, B);                 \
    if (ss == 48) return scl3::launch_one<MM, LL, 4, 8, 4>(P, cap, st, plan_only, B);                  \
    if (ss == 416) return scl3::launch_one<MM, LL, 4, 16, 2>(P, cap, st, plan_only, B);                \
    if (ss == 412) return scl3::launch_one<MM, LL, 4, 12, 2>(P, cap, st, plan_only, B);                \
    if (ss == 68) return scl3::launch_one<MM, LL, 6, 6, 2>(P, cap, st, plan_only, B);                  \
    return scl3::launch_one<MM, LL, 5, 4, 4>(P, cap, st, plan_only, B);                                \
  }
#else
#define POLAR_SCL3_L(MM, LL)                                                                          \
  if (L == LL) return scl3::launch_one<MM, LL, 5, 4, 4>(P, cap, st, plan_only, B);
#endif
#if defined(POLAR_SCL3_DEV)
#define POLAR_SCL3_M(MM) if (m == MM) { POLAR_SCL3_L(MM, 8) }
  POLAR_SCL3_M(10)
#else
#define POLAR_SCL3_M(MM) if (m == MM) { POLAR_SCL3_L(MM, 2) POLAR_SCL3_L(MM, 4) POLAR_SCL3_L(MM, 8) POLAR_SCL3_L(MM, 16) POLAR_SCL3_L(MM, 32) }
  POLAR_SCL3_M(6) POLAR_SCL3_M(7) POLAR_SCL3_M(8) POLAR_SCL3_M(9) POLAR_SCL3_M(10) POLAR_SCL3_M(11) POLAR_SCL3_M(12)
#endif
#undef POLAR_SCL3_M
#undef POLAR_SCL3_L
  return set_error(POLAR_EINVAL, "scl3: unsupported n=%d L=%d", n, L);
}

}  // namespace polar

#if !defined(POLAR_F_BOXPLUS)
// Debug / test hook (not part of include/polar_b200.h): d_mismatch[3] += number of arguments (of `count` pseudo-random
// ones in [-30, 30]) on which exp_nb / log_nb / softplus_literal differ bitwise from the CUDA math library.
extern "C" int polar_scl3_math_selftest(uint64_t count, unsigned long long *d_mismatch, void *stream) {
  if (!d_mismatch) return polar::set_error(POLAR_EINVAL, "selftest: null counter");
  polar::scl3::math_selftest_kernel<<<592, 256, 0, (cudaStream_t)stream>>>(count, d_mismatch);
  polar::count_launch();
  POLAR_CHECK_LAUNCH("math_selftest_kernel");
  return POLAR_OK;
}
#endif
