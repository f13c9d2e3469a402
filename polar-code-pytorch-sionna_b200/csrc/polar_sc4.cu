// polar_sc4.cu -- SC decoder, warp-autonomous mapping (default for 128 <= n <= 2048).
//
// Same algorithm and exact semantics as polar_sc.cu / polar_sc3.cu (x_run_sn_polar/polar/polar_sc.py:54-133,
// SURVEY.md Appendix A).  Measurements of polar_sc3.cu (CTA per 32 codewords: one warp walks the 64-leaf
// subtrees while three helper warps wait at a barrier) showed that the decoder is bound by the latency of
// that serial walk and that co-resident CTAs do not slow each other down -- throughput is simply (serial
// chains in flight per SM) / (latency of a chain).  So here EVERY warp is a chain:
//   * a warp owns 32 codewords for the whole decode and never synchronises with another warp
//     (__syncwarp only); one persistent CTA per SM holds as many such warps as on-chip storage allows
//     (n=1024: 7 warps = 224 codewords in flight per SM instead of 96..128);
//   * wide stages are processed cooperatively by the 32 lanes (4 elements per lane and round), the 64-leaf
//     subtrees with one lane per codeword, entirely in registers (BetaTree, polar_common.cuh);
//   * storage per codeword, n = 1024: channel stage in global memory / L2; stage 9 virtual (recomputed from
//     the channel, see polar_sc3.cu); stage 8 in TENSOR MEMORY (tcgen05.st/ld 32x32b -- lane l of warp w owns
//     TMEM lane 32(w%4)+l, 256 columns per warp; the tensor cores are idle, their 256 KB per SM is free
//     storage); stages 7 and 6 plus the partial-sum words in shared memory (916 B per codeword).
//   * only partial sums are produced on the serial path; the decisions are recovered once per codeword as
//     u = T(x_hat).
#include <atomic>
#include <mutex>

#include "polar_common.cuh"
#include "polar_internal.h"

namespace polar {

// phase timeline of warp 0 of CTA 0 (cycles), filled only when POLAR_SC3_DBG=1 (tools/perf_probe.py):
// 0 virtual steps, 1 g steps, 2 f steps, 3 64-leaf subtrees, 4 merges, 5 outputs, 6 total, 7 batches
__device__ unsigned long long g_sc4_dbg[8];

namespace {

constexpr unsigned FULLMASK = 0xFFFFFFFFu;
#define SC4_T(slot)                                                                   \
  do {                                                                                \
    if (dbg && threadIdx.x == 0 && blockIdx.x == 0) {                                 \
      const long long t__ = clock64(); g_sc4_dbg[slot] += (unsigned long long)(t__ - tlast); tlast = t__; \
    }                                                                                 \
  } while (0)

constexpr size_t kSc4ScratchPerSm = (size_t)512 * 1024;   // 8 warps x 32 codewords x 512 floats (n=1024) = 4 x 32 x 1024 (n=2048)

struct Sc4Layout {
  int nw, nws, n64, top, stride;
  size_t nz_off, warp_off, per_warp, total;
};
__host__ __device__ inline Sc4Layout sc4_layout(int m, int top, int bot, int warps) {
  Sc4Layout l;
  const int n = 1 << m;
  l.nw = n >> 5; l.nws = l.nw + 1; l.n64 = n >> bot;    // n64: number of 2^bot-leaf blocks
  l.top = top;                                        // highest stage kept in shared memory (>= 6)
  l.stride = (2 << l.top) - (1 << bot) + 4;           // floats per codeword row (stages bot..top); stride/4 is odd
  l.nz_off = (size_t)((l.nw * 4 + 15) / 16) * 16;
  l.warp_off = l.nz_off + (size_t)((2 * l.n64 + 15) / 16) * 16 + 16;   // +16: tensor-memory base address slot
  l.per_warp = (((size_t)32 * l.stride * 4 + (size_t)32 * l.nws * 4) + 15) / 16 * 16;
  l.total = l.warp_off + (size_t)warps * l.per_warp;
  return l;
}

PDEV float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
// 128-bit read-only load with an L2 eviction policy (createpolicy) and no L1 allocation
PDEV float4 ldg4_hint(const float *p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
  return v;
}
PDEV uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
PDEV uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
PDEV uint64_t l2_policy_evict_normal() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
// coherent 128-bit load / store with an L2 eviction policy (the stage scratch is written and read by the same kernel)
PDEV float4 ldg4_coh_hint(const float *p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy) : "memory");
  return v;
}
PDEV void stg4_hint(float *p, const float4 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
               ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}
PDEV float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
PDEV void sts4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
PDEV float4 f4(const float4 a, const float4 b) {
  float4 o;
  o.x = f_minsum(a.x, b.x); o.y = f_minsum(a.y, b.y); o.z = f_minsum(a.z, b.z); o.w = f_minsum(a.w, b.w);
  return o;
}
PDEV float4 f4neg(const float4 a, const float4 b) {   // f on logits (see f_minsum_neg)
  float4 o;
  o.x = f_minsum_neg(a.x, b.x); o.y = f_minsum_neg(a.y, b.y); o.z = f_minsum_neg(a.z, b.z); o.w = f_minsum_neg(a.w, b.w);
  return o;
}
PDEV float4 g4(const float4 a, const float4 b, const uint32_t bits) {   // bit e of `bits` = partial sum of element e
  float4 o;
  o.x = g_minsum(a.x, b.x, (bits << 31) & 0x80000000u);
  o.y = g_minsum(a.y, b.y, (bits << 30) & 0x80000000u);
  o.z = g_minsum(a.z, b.z, (bits << 29) & 0x80000000u);
  o.w = g_minsum(a.w, b.w, (bits << 28) & 0x80000000u);
  return o;
}
// g(-a, -b, u): the operands are logits, the LLR is their negation (polar_sc.py:122)
PDEV float gneg(float a, float b, uint32_t signmask) { return u2f(f2u(a) ^ signmask ^ 0x80000000u) - b; }
PDEV float4 g4neg(const float4 a, const float4 b, const uint32_t bits) {
  float4 o;
  o.x = gneg(a.x, b.x, (bits << 31) & 0x80000000u);
  o.y = gneg(a.y, b.y, (bits << 30) & 0x80000000u);
  o.z = gneg(a.z, b.z, (bits << 29) & 0x80000000u);
  o.w = gneg(a.w, b.w, (bits << 28) & 0x80000000u);
  return o;
}

// ---- tensor memory as per-thread scratch (32x32b: lane l of the warp owns TMEM lane base+l) -----
PDEV void tmem_st8(uint32_t taddr, const float4 a, const float4 b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(f2u(a.x)), "r"(f2u(a.y)), "r"(f2u(a.z)), "r"(f2u(a.w)), "r"(f2u(b.x)), "r"(f2u(b.y)),
               "r"(f2u(b.z)), "r"(f2u(b.w)) : "memory");
}
PDEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
struct Tm8 { uint32_t r[8]; };
PDEV void tmem_ld8_issue(uint32_t taddr, Tm8 &v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7])
               : "r"(taddr) : "memory");
}
// the registers are defined only after the wait; tying them to it keeps every use behind it
PDEV void tmem_ld_wait(Tm8 &a, Tm8 &b) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a.r[0]), "+r"(a.r[1]), "+r"(a.r[2]), "+r"(a.r[3]), "+r"(a.r[4]), "+r"(a.r[5]), "+r"(a.r[6]), "+r"(a.r[7]),
                 "+r"(b.r[0]), "+r"(b.r[1]), "+r"(b.r[2]), "+r"(b.r[3]), "+r"(b.r[4]), "+r"(b.r[5]), "+r"(b.r[6]), "+r"(b.r[7])
               :: "memory");
}
PDEV float4 tm_lo(const Tm8 &v) { return make_float4(u2f(v.r[0]), u2f(v.r[1]), u2f(v.r[2]), u2f(v.r[3])); }
PDEV float4 tm_hi(const Tm8 &v) { return make_float4(u2f(v.r[4]), u2f(v.r[5]), u2f(v.r[6]), u2f(v.r[7])); }

// ---- cooperative steps of ONE warp over its 32 codewords ------------------------------------------
// stage S+1 -> S inside shared memory.  out[j] = f(a[j], a[j+H]) or g(a[j], a[j+H], beta_left[j]).
template <int S, bool IS_G, int BOT>
PDEV void step_smem(float *L, const uint32_t *beta, int stride, int nws, int lane, int left_word) {
  constexpr int H = 1 << S, HQ = H >> 2, ITEMS = 32 * HQ;
  float *dst = L + (H - (1 << BOT));
  const float *src = L + (2 * H - (1 << BOT));
#pragma unroll 4
  for (int it = lane; it < ITEMS; it += 32) {
    const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
    const float4 a = lds4(src + c * stride + j), b = lds4(src + c * stride + j + H);
    float4 o;
    if (IS_G) o = g4(a, b, beta[c * nws + left_word + (j >> 5)] >> (j & 31));
    else o = f4(a, b);
    sts4(dst + c * stride + j, o);
  }
}
template <int SMAX, bool IS_G, int BOT>
PDEV void step_smem_any(int s, float *L, const uint32_t *beta, int stride, int nws, int lane, int left_word) {
  if constexpr (SMAX >= BOT) {
    if (s == SMAX) step_smem<SMAX, IS_G, BOT>(L, beta, stride, nws, lane, left_word);
    else step_smem_any<SMAX - 1, IS_G, BOT>(s, L, beta, stride, nws, lane, left_word);
  }
}

// channel (global, stage M) -> stage M-1 in shared memory (n <= 512: everything fits in shared memory).
template <int M, bool IS_G, int BOT>
__device__ __noinline__ void step_glob(const float *__restrict__ logit, int64_t cw0, int nvalid, float *L,
                                       const uint32_t *beta, int stride, int nws, int lane) {
  constexpr int N = 1 << M, H = N >> 1, HQ = H >> 2, ITEMS = 32 * HQ;
  constexpr int U = 4;        // items per round: 8 independent 128-bit loads in flight per lane
  float *dst = L + (H - (1 << BOT));
#pragma unroll 1
  for (int it0 = lane; it0 < ITEMS; it0 += U * 32) {
    float4 a[U], b[U];
#pragma unroll
    for (int r = 0; r < U; ++r) {
      const int it = it0 + r * 32;
      const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
      const int cl = c < nvalid ? c : nvalid - 1;
      const float *row = logit + (cw0 + cl) * (int64_t)N + j;
      a[r] = ldg4(row); b[r] = ldg4(row + H);
    }
#pragma unroll
    for (int r = 0; r < U; ++r) {
      const int it = it0 + r * 32;
      const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
      float4 o;
      if (IS_G) o = g4neg(a[r], b[r], beta[c * nws + (j >> 5)] >> (j & 31));
      else o = f4neg(a[r], b[r]);                       // f(-a,-b) == f(a,b) for min-sum
      sts4(dst + c * stride + j, o);
    }
  }
}

// channel (global, stage M) -> stage M-2 in TENSOR MEMORY through the virtual stage M-1.
// kind = quarter of the codeword the target node covers: 0 LL, 1 LR, 2 RL, 3 RR (warp-uniform).
// Work item p = lane + 32k = (codeword c, pair q): the float4 at elements 4q and 4q + H/2 of the stage M-2
// node, i.e. exactly what one f/g of the next step consumes; it goes to TMEM columns 8k..8k+7 of the lane.
// scr (optional): per-warp global scratch [32 codewords][N/2 floats] that stays in the L2.  The passes that compute a
// LEFT stage M-2 node (kind 0 / 2) also store the stage M-1 node they had to form; the pass for its right sibling
// (kind 1 / 3, a quarter of a decode later) then reads those N/2 values instead of the whole channel row again and
// skips two of its three f/g per element.  scr_load is only set when the left pass of this batch really ran (it is
// skipped when the left node is rate-0).
template <int M, bool STORE>
__device__ __noinline__ void step_virt_tmem(const int kind, const float *__restrict__ logit, int64_t cw0, int nvalid,
                                            const uint32_t *beta, int nws, int lane, uint32_t tm_base, int hints,
                                            float *scr) {
  constexpr int N = 1 << M, H = N >> 2, PQ = H >> 3, HW = H >> 5;   // PQ pairs per codeword (multiple of 32)
  constexpr int KMAX = PQ;                                          // 32 codewords * PQ pairs / 32 lanes
  constexpr int VU = 2;                                             // pairs per round; rounds are double buffered
  // The row is read four times, a quarter of the decode apart, and the rows in flight (148 SMs x 8 warps x 32 x 4 KB)
  // exceed the L2.  hints = 1: keep every row until its last pass (evict_last x3, evict_first).  hints = 2: only protect
  // the pairs of passes that are a quarter apart (0->1 and 2->3), halving the protected set so that it fits.
  // With the stage scratch (STORE) the sibling pass does not come back to the row: stream it (evict_first) and leave
  // the L2 to the scratch.
  const uint64_t pol = STORE ? l2_policy_evict_first()
                     : !hints ? l2_policy_evict_normal()
                     : (hints == 2) ? ((kind & 1) ? l2_policy_evict_first() : l2_policy_evict_last())
                                    : ((kind == 3) ? l2_policy_evict_first() : l2_policy_evict_last());
  const bool right = kind >= 2, is_g = kind & 1;
  const int gw = (kind == 3) ? 2 * HW : 0;
  float4 c0[2][VU][2], c1[2][VU][2], c2[2][VU][2], c3[2][VU][2];      // [buffer][pair][element]
  constexpr bool scr_store = STORE;
  auto issue = [&](int buf, int k0) {
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
      const int cl = c < nvalid ? c : nvalid - 1;
      const float *row = logit + (cw0 + cl) * (int64_t)N + 4 * q;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float *rp = row + e * (H / 2);
        c0[buf][r][e] = ldg4_hint(rp, pol); c1[buf][r][e] = ldg4_hint(rp + H, pol);
        c2[buf][r][e] = ldg4_hint(rp + 2 * H, pol); c3[buf][r][e] = ldg4_hint(rp + 3 * H, pol);
      }
    }
  };
  auto compute = [&](int buf, int k0) {
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
      float4 o[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 4 * q + e * (H / 2);
        const uint32_t *bw = beta + c * nws + (j >> 5);
        const int sh = j & 31;
        float4 y0, y1;
        if (!right) {              // left half of the codeword: stage M-1 node = f(channel)
          y0 = f4neg(c0[buf][r][e], c2[buf][r][e]); y1 = f4neg(c1[buf][r][e], c3[buf][r][e]);
        } else {                   // right half: stage M-1 node = g(channel, beta of the left half)
          y0 = g4neg(c0[buf][r][e], c2[buf][r][e], bw[0] >> sh); y1 = g4neg(c1[buf][r][e], c3[buf][r][e], bw[HW] >> sh);
        }
        if constexpr (scr_store) {
          float *sp = scr + c * (2 * H) + j;
          const uint64_t pol_scr = l2_policy_evict_last();
          stg4_hint(sp, y0, pol_scr); stg4_hint(sp + H, y1, pol_scr);
        }
        if (!is_g) o[e] = f4(y0, y1);
        else o[e] = g4(y0, y1, bw[gw] >> sh);
      }
      tmem_st8(tm_base + 8 * (k0 + r), o[0], o[1]);
    }
  };
  // software pipeline: the loads of round r+1 are in flight while round r is computed (static buffer indices)
  issue(0, 0);
#pragma unroll 1
  for (int k0 = 0; k0 < KMAX; k0 += 2 * VU) {
    issue(1, k0 + VU);
    compute(0, k0);
    if (k0 + 2 * VU < KMAX) issue(0, k0 + 2 * VU);
    compute(1, k0 + VU);
  }
  tmem_wait_st();
}

// The right-sibling pass (kind 1 / 3) when the left pass of this batch stored the stage M-1 node: one g per element
// from N/2 scratch values per codeword instead of three f/g from the whole channel row.
template <int M>
__device__ __noinline__ void step_virt_scr(const int kind, const uint32_t *beta, int nws, int lane, uint32_t tm_base,
                                           const float *scr, const bool discard) {
  constexpr int N = 1 << M, H = N >> 2, PQ = H >> 3, HW = H >> 5;
  constexpr int KMAX = PQ, VU = 4;
  const uint64_t pol = l2_policy_evict_last();
  const int gw = (kind == 3) ? 2 * HW : 0;
#pragma unroll 1
  for (int k0 = 0; k0 < KMAX; k0 += VU) {
    float4 a[VU][2], b[VU][2];
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
      const float *sp = scr + c * (2 * H) + 4 * q;
#pragma unroll
      for (int e = 0; e < 2; ++e) { a[r][e] = ldg4_coh_hint(sp + e * (H / 2), pol); b[r][e] = ldg4_coh_hint(sp + e * (H / 2) + H, pol); }
    }
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
      float4 o[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 4 * q + e * (H / 2);
        o[e] = g4(a[r][e], b[r][e], beta[c * nws + (j >> 5) + gw] >> (j & 31));
      }
      tmem_st8(tm_base + 8 * (k0 + r), o[0], o[1]);
      if (discard && (lane & 7) == 0) {      // the 8 lanes of a 128-byte line have consumed it: drop it from the L2 unwritten
        const float *sp = scr + c * (2 * H) + 4 * q;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(sp + e * (H / 2)) : "memory");
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(sp + e * (H / 2) + H) : "memory");
        }
      }
    }
  }
  tmem_wait_st();
}

// channel (global, stage M) -> stage M-1 in TENSOR MEMORY (n = 512: no virtual stage needed).
// Work item p = lane + 32k = (codeword c, pair q): the float4 at elements 4q and 4q + H/2 of the stage M-1 node.
template <int M, bool IS_G>
__device__ __noinline__ void step_glob_tmem(const float *__restrict__ logit, int64_t cw0, int nvalid, const uint32_t *beta,
                                            int nws, int lane, uint32_t tm_base) {
  constexpr int N = 1 << M, H = N >> 1, PQ = H >> 3;   // PQ pairs per codeword (multiple of 32)
  constexpr int KMAX = PQ, VU = 4;
#pragma unroll 1
  for (int k0 = 0; k0 < KMAX; k0 += VU) {
    float4 a[VU][2], b[VU][2];
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
      const int cl = c < nvalid ? c : nvalid - 1;
      const float *row = logit + (cw0 + cl) * (int64_t)N + 4 * q;
#pragma unroll
      for (int e = 0; e < 2; ++e) { a[r][e] = ldg4(row + e * (H / 2)); b[r][e] = ldg4(row + e * (H / 2) + H); }
    }
#pragma unroll
    for (int r = 0; r < VU; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
      float4 o[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = 4 * q + e * (H / 2);
        if (IS_G) o[e] = g4neg(a[r][e], b[r][e], beta[c * nws + (j >> 5)] >> (j & 31));
        else o[e] = f4neg(a[r][e], b[r][e]);              // f(-a,-b) == f(a,b) for min-sum
      }
      tmem_st8(tm_base + 8 * (k0 + r), o[0], o[1]);
    }
  }
  tmem_wait_st();
}

// stage TS (tensor memory, lane-private pairs) -> stage TS-1 in shared memory.
template <int TS, bool IS_G, int BOT>   // TS = stage held in tensor memory; writes stage TS-1
PDEV void step_tmem(float *L, const uint32_t *beta, int stride, int nws, int lane, uint32_t tm_base, int left_word) {
  constexpr int H = 1 << (TS - 1), PQ = H >> 2;        // H outputs per codeword = PQ float4
  constexpr int KMAX = PQ;
  float *dst = L + (H - (1 << BOT));
#pragma unroll 2
  for (int k0 = 0; k0 < KMAX; k0 += 2) {
    Tm8 v0, v1;
    tmem_ld8_issue(tm_base + 8 * k0, v0);
    tmem_ld8_issue(tm_base + 8 * (k0 + 1), v1);
    tmem_ld_wait(v0, v1);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int p = lane + 32 * (k0 + r);
      const int c = (int)((unsigned)p / (unsigned)PQ), j = (int)((unsigned)p % (unsigned)PQ) << 2;
      const float4 a = tm_lo(r ? v1 : v0), b = tm_hi(r ? v1 : v0);
      float4 o;
      if (IS_G) o = g4(a, b, beta[c * nws + left_word + (j >> 5)] >> (j & 31));
      else o = f4(a, b);
      sts4(dst + c * stride + j, o);
    }
  }
}

// ask the L2 for the channel rows of the warp's next batch (one 4 KB bulk prefetch per lane and round)
PDEV void prefetch_rows_l2(const float *base, size_t bytes, int lane) {
  const char *p = reinterpret_cast<const char *>(base);
  for (size_t off = (size_t)lane * 4096; off < bytes; off += (size_t)32 * 4096) {
    const unsigned sz = (unsigned)((bytes - off) < 4096 ? (bytes - off) : 4096);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(sz & ~15u) : "memory");
  }
}

// ---- one lane per codeword: the 64-leaf subtree below the lane's stage-6 node (shared memory) -----
PDEV uint2 bottom64(const float *node, uint32_t fm0, uint32_t fm1) {
  uint32_t bl = 0, bc = 0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {      // rolled: one copy of the 32-leaf subtree code
    const uint32_t fmc = h ? fm1 : fm0;
    if (fmc == FULLMASK) { bc = 0; continue; }
    float x[32];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = lds4(node + 4 * q), b = lds4(node + 32 + 4 * q);
      const float4 o = h ? g4(a, b, bl >> (4 * q)) : f4(a, b);
      x[4 * q] = o.x; x[4 * q + 1] = o.y; x[4 * q + 2] = o.z; x[4 * q + 3] = o.w;
    }
    bc = BetaTree<5>::run(x, fmc);
    if (h == 0) bl = bc;
  }
  return make_uint2(bl ^ bc, bc);
}

// MODE 0: stages 6..M-1 in shared memory (n <= 256).  MODE 1: stage M-1 in tensor memory (n = 512).
// MODE 2: stage M-1 virtual, stage M-2 in tensor memory (n >= 1024).
// the 128-leaf subtree below the lane's stage-7 node: both 64-leaf halves in registers (x[64] + the BetaTree levels),
// so that shared memory only has to hold stage 7 (one more warp per SM) and stage 6 costs no LDS/STS round trip.
PDEV uint64_t tree64(const float (&x)[64], uint64_t fm) {
  uint32_t bl = 0, bc = 0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const uint32_t fmc = h ? (uint32_t)(fm >> 32) : (uint32_t)fm;
    if (fmc == FULLMASK) { bc = 0; continue; }
    float y[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
      y[j] = h ? g_minsum(x[j], x[j + 32], (bl << (31 - j)) & 0x80000000u) : f_minsum(x[j], x[j + 32]);
    bc = BetaTree<5>::run(y, fmc);
    if (h == 0) bl = bc;
  }
  return (uint64_t)(bl ^ bc) | ((uint64_t)bc << 32);
}
PDEV uint4 bottom128(const float *node, uint64_t fm0, uint64_t fm1) {
  uint64_t bl = 0, bc = 0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const uint64_t fmc = h ? fm1 : fm0;
    if (fmc == ~0ull) { bc = 0; continue; }
    float x[64];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float4 a = lds4(node + 4 * q), b = lds4(node + 64 + 4 * q);
      const float4 o = h ? g4(a, b, (uint32_t)(bl >> (4 * q))) : f4(a, b);
      x[4 * q] = o.x; x[4 * q + 1] = o.y; x[4 * q + 2] = o.z; x[4 * q + 3] = o.w;
    }
    bc = tree64(x, fmc);
    if (h == 0) bl = bc;
  }
  const uint64_t lo = bl ^ bc;
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)bc, (uint32_t)(bc >> 32));
}

template <int M, int MODE>
__global__ void __launch_bounds__(256, 1) sc4_kernel(const float *__restrict__ logit, const uint32_t *__restrict__ fmask_g,
                                                     int64_t B, int64_t nbatches, int l2_prefetch, int l2_hints, int dbg,
                                                     float *scratch, int scr_discard, uint32_t *__restrict__ u_packed, float *__restrict__ u_info,
                                                     const int32_t *__restrict__ info_pos, int k) {
  constexpr bool TM = MODE >= 1, VIRT = MODE == 2;
  constexpr int TS = VIRT ? M - 2 : M - 1;                  // stage held in tensor memory (TM only)
  static_assert(MODE == 0 ? (M >= 7) : (TS >= 8 && TS <= 9), "sc4: the bottom stage must exist in shared memory; a warp reaches 512 TMEM columns");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int BOT = (M >= 8) ? 7 : 6;                      // stage of the node the per-lane subtree starts from
  constexpr int WB = 1 << (BOT - 5);                         // 32-bit words per bottom block
  constexpr int N = 1 << M, NW = N >> 5, NWS = NW + 1, N64 = N >> BOT, TOP = TM ? TS - 1 : M - 1;
  constexpr int TM_COLS_WARP = TM ? (1 << TS) : 32;         // 32 codewords x 2^TS floats / 32 lanes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const Sc4Layout lay = sc4_layout(M, TOP, BOT, nwarps);
  constexpr int stride = (2 << TOP) - (1 << BOT) + 4;   // == lay.stride, compile-time so that row offsets fold into immediates
  uint32_t *fmask = reinterpret_cast<uint32_t *>(smem_raw);
  unsigned char *nz = smem_raw + lay.nz_off;    // nz[(N64 >> lv) + (i >> lv)]: node of 2^lv 64-blocks at block i is rate-0
  uint32_t *tm_slot = reinterpret_cast<uint32_t *>(smem_raw + lay.warp_off - 16);
  float *L = reinterpret_cast<float *>(smem_raw + lay.warp_off + (size_t)warp * lay.per_warp);
  uint32_t *beta = reinterpret_cast<uint32_t *>(L + 32 * stride);

  uint32_t tm_base = 0;
  if (TM) {
    if (warp == 0) {          // the CTA is alone on its SM (shared memory): take all 512 columns
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"((uint32_t)__cvta_generic_to_shared(tm_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  for (int i = tid; i < NW; i += blockDim.x) fmask[i] = __ldg(fmask_g + i);
  __syncthreads();
  if (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // warp w: TMEM lanes 32(w%4)..+31 (the only ones it can address), column block w/4
    tm_base = *tm_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * TM_COLS_WARP);
  }
  for (int i = tid; i < N64; i += blockDim.x) {
    uint32_t all = FULLMASK;
    for (int w = 0; w < WB; ++w) all &= fmask[WB * i + w];
    nz[N64 + i] = all == FULLMASK;
  }
  __syncthreads();
  if (tid == 0)
    for (int idx = N64 - 1; idx >= 1; --idx) nz[idx] = nz[2 * idx] & nz[2 * idx + 1];
  __syncthreads();

  // stage scratch of this warp: indexed by the PHYSICAL SM (only one CTA of this kernel fits on an SM, so concurrent
  // launches on other streams can never share a slot), kSc4ScratchPerSm bytes per SM.  Recomputed at each use: the
  // 128-leaf subtrees need every register, nothing extra may stay live across them.
  auto scr_ptr = [&]() -> float * {
    if (!VIRT || !scratch) return nullptr;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    return scratch + (size_t)smid * (kSc4ScratchPerSm / 4) + (size_t)(threadIdx.x >> 5) * (32 * (N / 2));
  };
  long long tlast = clock64();
  const long long tstart = tlast;
  const int64_t wstride = (int64_t)gridDim.x * nwarps;
  for (int64_t batch = (int64_t)warp * gridDim.x + blockIdx.x; batch < nbatches; batch += wstride) {
    const int64_t cw0 = batch * 32;
    const int nvalid = (int)((B - cw0) < 32 ? (B - cw0) : 32);
    int i = 0;                                   // current 64-leaf block
    while (i < N64) {
      // node entered at block i: the root, or the right child whose left sibling just finished
      const int S = (i == 0) ? M : BOT + (__ffs(i) - 1);
      int s = S;
      if (VIRT && l2_prefetch && ((i + 1) & ((l2_prefetch == 2 ? N64 / 2 : N64 / 4) - 1)) == 0) {
        // the block after this one starts with a pass over the channel rows (this batch's next quarter, or the next
        // batch's first): ask the L2 for them now, one 64-leaf.. block (several microseconds) ahead of the loads
        const int64_t pb = (i + 1 < N64) ? batch : batch + wstride;
        if (pb < nbatches) {
          const int64_t r = pb * 32 + lane;
          if (r < B) {
            const float *rp = logit + r * (int64_t)N;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rp), "r"(N * 4) : "memory");
          }
        }
      }
      bool zeroed = nz[(N64 >> (S - BOT)) + (i >> (S - BOT))] != 0;
      if (!zeroed && S < M && !(VIRT && S == M - 1)) {
        // g step into (S, i) from its parent at stage S+1; the left sibling's beta starts at word 2*(i - 2^(S-6))
        const int left_word = WB * (i - (1 << (S - BOT)));
        if (VIRT && S == M - 2) {
          // the left sibling's pass stored its stage M-1 node unless it was skipped (left quarter rate-0)
          const int kind = i < N64 / 2 ? 1 : 3;
          float *scr = scr_ptr();
          if (scr && !nz[4 + kind - 1]) step_virt_scr<M>(kind, beta, NWS, lane, tm_base, scr, scr_discard != 0);
          else step_virt_tmem<M, false>(kind, logit, cw0, nvalid, beta, NWS, lane, tm_base, l2_hints, nullptr);
        } else if (TM && !VIRT && S == M - 1) {
          step_glob_tmem<M, true>(logit, cw0, nvalid, beta, NWS, lane, tm_base);
        } else if (TM && S == TS - 1) {
          step_tmem<TS, true, BOT>(L, beta, stride, NWS, lane, tm_base, left_word);
        } else if (!TM && S == M - 1) {
          step_glob<M, true, BOT>(logit, cw0, nvalid, L, beta, stride, NWS, lane);
        } else {
          step_smem_any<TOP - 1, true, BOT>(S, L, beta, stride, NWS, lane, left_word);
        }
        __syncwarp();
        SC4_T((VIRT && S == M - 2) ? 0 : 1);
      }
      while (!zeroed && s > BOT) {
        if (nz[(N64 >> (s - 1 - BOT)) + (i >> (s - 1 - BOT))]) { zeroed = true; --s; break; }   // left child is rate-0
        if (VIRT && s == M) { --s; continue; }                                         // virtual stage: nothing stored
        if (VIRT && s == M - 1) {
          {
            float *scr = scr_ptr();
            if (scr) step_virt_tmem<M, true>(i < N64 / 2 ? 0 : 2, logit, cw0, nvalid, beta, NWS, lane, tm_base, l2_hints, scr);
            else step_virt_tmem<M, false>(i < N64 / 2 ? 0 : 2, logit, cw0, nvalid, beta, NWS, lane, tm_base, l2_hints, nullptr);
          }
        } else if (TM && !VIRT && s == M) {
          step_glob_tmem<M, false>(logit, cw0, nvalid, beta, NWS, lane, tm_base);
        } else if (TM && s == TS) {
          step_tmem<TS, false, BOT>(L, beta, stride, NWS, lane, tm_base, 0);
        } else if (!TM && s == M) {
          step_glob<M, false, BOT>(logit, cw0, nvalid, L, beta, stride, NWS, lane);
        } else {
          step_smem_any<TOP - 1, false, BOT>(s - 1, L, beta, stride, NWS, lane, 0);
        }
        __syncwarp();
        SC4_T((VIRT && s == M - 1) ? 0 : 2);
        --s;
      }
      const int lv0 = s - BOT;                   // the finished node covers 2^lv0 bottom blocks starting at i
      if (zeroed) {
        const int nwd = WB << lv0;
        for (int q = lane; q < 32 * nwd; q += 32) {
          const int c = q / nwd, w = q & (nwd - 1);
          beta[c * NWS + WB * i + w] = 0u;
        }
      } else if (BOT == 7) {
        const uint32_t *fmw = fmask + 4 * i;
        const uint4 b = bottom128(L + lane * stride, (uint64_t)fmw[0] | ((uint64_t)fmw[1] << 32),
                                  (uint64_t)fmw[2] | ((uint64_t)fmw[3] << 32));
        uint32_t *bp = beta + lane * NWS + 4 * i;
        bp[0] = b.x; bp[1] = b.y; bp[2] = b.z; bp[3] = b.w;
      } else {
        const uint2 b = bottom64(L + lane * stride, fmask[2 * i], fmask[2 * i + 1]);
        uint32_t *bp = beta + lane * NWS + 2 * i;
        bp[0] = b.x; bp[1] = b.y;
      }
      __syncwarp();
      SC4_T(3);
      {  // merge partial sums upward while the finished node is a right child: [bl ^ br, br] (polar_sc.py:83-89)
        int lv = lv0, a = i;
        while (lv < M - BOT && ((a >> lv) & 1)) {
          const int nwd = WB << lv, left = a - (1 << lv);
          for (int q = lane; q < 32 * nwd; q += 32) {
            const int c = q / nwd, w = q & (nwd - 1);
            beta[c * NWS + WB * left + w] ^= beta[c * NWS + WB * a + w];
          }
          __syncwarp();
          a = left; ++lv;
        }
      }
      SC4_T(4);
      i += 1 << lv0;
    }
    // beta now holds the re-encoded codeword x_hat of every codeword; the decisions are u = T(x_hat)
    // (my_sn/fec/polar/enc.py:85-96 is an involution): 5 stages inside each word, M-5 across words.
    for (int q = lane; q < 32 * NW; q += 32) {
      const int c = q / NW, w = q % NW;
      beta[c * NWS + w] = ptransform<5>(beta[c * NWS + w]);
    }
    __syncwarp();
#pragma unroll 1
    for (int st = 0; st < M - 5; ++st) {
      for (int q = lane; q < 32 * (NW / 2); q += 32) {
        const int c = q / (NW / 2), r = q % (NW / 2);
        const int w = ((r >> st) << (st + 1)) | (r & ((1 << st) - 1));     // word index with bit st clear
        beta[c * NWS + w] ^= beta[c * NWS + w + (1 << st)];
      }
      __syncwarp();
    }
    if (u_packed) {
      for (int q = lane; q < 32 * NW; q += 32) {
        const int c = q / NW, w = q % NW;
        if (c < nvalid) u_packed[(cw0 + c) * NW + w] = beta[c * NWS + w];
      }
    }
    if (u_info) {
      // the API tensor [B, k] fp32 (polar_sc.py:127-133): a codeword's row at a time, one float4 per lane and round
      // (bit -> 0.f / 1.f by masking the bit pattern of 1.0f: no int-to-float conversion on the slow pipe)
      if ((k & 3) == 0 && (reinterpret_cast<uintptr_t>(u_info) & 15) == 0 && (reinterpret_cast<uintptr_t>(info_pos) & 15) == 0) {
        const int k4 = k >> 2;
        const int4 *ip = reinterpret_cast<const int4 *>(info_pos);
        for (int c = 0; c < nvalid; ++c) {
          float4 *row = reinterpret_cast<float4 *>(u_info + (cw0 + c) * (int64_t)k);
          const uint32_t *bw = beta + c * NWS;
          for (int t4 = lane; t4 < k4; t4 += 32) {
            const int4 p = __ldg(ip + t4);
            float4 o;
            o.x = u2f((0u - ((bw[p.x >> 5] >> (p.x & 31)) & 1u)) & 0x3f800000u);
            o.y = u2f((0u - ((bw[p.y >> 5] >> (p.y & 31)) & 1u)) & 0x3f800000u);
            o.z = u2f((0u - ((bw[p.z >> 5] >> (p.z & 31)) & 1u)) & 0x3f800000u);
            o.w = u2f((0u - ((bw[p.w >> 5] >> (p.w & 31)) & 1u)) & 0x3f800000u);
            __stcs(row + t4, o);
          }
        }
      } else {
        for (int q = lane; q < nvalid * k; q += 32) {
          const int c = q / k, t = q - c * k;
          const int p = __ldg(info_pos + t);
          u_info[(cw0 + c) * (int64_t)k + t] = (float)((beta[c * NWS + (p >> 5)] >> (p & 31)) & 1u);
        }
      }
    }
    __syncwarp();
    SC4_T(5);
    if (dbg && tid == 0 && blockIdx.x == 0) g_sc4_dbg[7] += 1;
  }
  if (dbg && tid == 0 && blockIdx.x == 0) g_sc4_dbg[6] += (unsigned long long)(clock64() - tstart);
  if (TM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tm_slot), "n"(512) : "memory");
  }
}

__global__ void nsmid_kernel(unsigned *out) {
  unsigned v;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
  *out = v;
}
#if !defined(POLAR_F_BOXPLUS)
std::mutex g_scr_mu;
std::atomic<float *> g_scr_buf[64];
#endif

}  // namespace

#if !defined(POLAR_F_BOXPLUS)     // one scratch per device for the whole library: the boxplus unit uses ::polar's

// Per-device stage scratch of the virtual-stage kernels (n >= 1024): %nsmid slots of kSc4ScratchPerSm (76 MB on a
// 148-SM part) -- small enough to stay resident in the L2.  Allocated once by polar_init(device) (which may allocate and
// synchronise; the decode entry points never do) and kept for the life of the process.
int sc4_scratch_init(int device) {
  if (device < 0 || device >= 64) return set_error(POLAR_EINVAL, "init: bad device %d", device);
  std::lock_guard<std::mutex> lk(g_scr_mu);
  if (g_scr_buf[device].load(std::memory_order_acquire)) return POLAR_OK;
  unsigned *d_n = nullptr, h_n = 0;
  POLAR_CUDA(cudaMalloc(&d_n, sizeof(unsigned)));
  nsmid_kernel<<<1, 1>>>(d_n);
  const cudaError_t e = cudaMemcpy(&h_n, d_n, sizeof(unsigned), cudaMemcpyDeviceToHost);
  cudaFree(d_n);
  if (e != cudaSuccess) return set_error(POLAR_ECUDA, "init: %s", cudaGetErrorString(e));
  if (h_n == 0 || h_n > 1024) return set_error(POLAR_ECUDA, "init: implausible %%nsmid = %u", h_n);
  void *p = nullptr;
  if (cudaMalloc(&p, (size_t)h_n * kSc4ScratchPerSm) != cudaSuccess) {
    (void)cudaGetLastError();
    return set_error(POLAR_ENOMEM, "init: cudaMalloc of the %zu-byte SC stage scratch failed", (size_t)h_n * kSc4ScratchPerSm);
  }
  g_scr_buf[device].store((float *)p, std::memory_order_release);
  return POLAR_OK;
}
float *sc4_scratch() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_scr_buf[dev].load(std::memory_order_acquire);
}
#endif

namespace {

template <int M, int MODE>
int launch_sc4_t(const float *logit, const uint32_t *fmask, int64_t B, uint32_t *u_packed, float *u_info,
                 const int32_t *info_pos, int k, int warps, cudaStream_t st) {
  const int max_smem = device_max_smem_optin();
  constexpr bool TM = MODE >= 1;
  constexpr int TS = MODE == 2 ? M - 2 : M - 1;
  int wmax = 8;
  if (TM) wmax = 4 * (512 >> TS);               // TMEM columns: 2^TS per warp, 512 per lane quarter
  if (wmax > 8) wmax = 8;
  if (warps <= 0 || warps > wmax) warps = wmax;
  constexpr int TOP = TM ? TS - 1 : M - 1, BOT = (M >= 8) ? 7 : 6;
  while (warps > 1 && sc4_layout(M, TOP, BOT, warps).total > (size_t)max_smem) --warps;
  const Sc4Layout lay = sc4_layout(M, TOP, BOT, warps);
  if (lay.total > (size_t)max_smem) return set_error(POLAR_ENOMEM, "sc: n=%d needs %zu B shared memory per CTA", 1 << M, lay.total);
  auto kern = sc4_kernel<M, MODE>;
  // one persistent CTA per SM.  With tensor memory the CTA takes all 512 columns, so a second CTA must never
  // become resident on the same SM: pad the request above half of the SM's shared memory.
  size_t smem = lay.total;
  if (TM && smem < (size_t)116 * 1024) smem = (size_t)116 * 1024;
  POLAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = device_sm_count();
  const int64_t nbatches = (B + 31) / 32;
  int64_t grid = (nbatches + warps - 1) / warps;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  float *scratch = (MODE == 2 && env_int("POLAR_SC4_SCRATCH", 1)) ? sc4_scratch() : nullptr;
  if (MODE == 2 && env_int("POLAR_SC4_SCRATCH", 1) && !scratch)
    return set_error(POLAR_EINVAL, "sc: n=%d needs the per-device stage scratch -- call polar_init(device) once before decoding", 1 << M);
  kern<<<(unsigned)grid, warps * 32, smem, st>>>(logit, fmask, B, nbatches, env_int("POLAR_SC4_PREFETCH", 0),
                                                  env_int("POLAR_SC4_HINTS", 2), env_int("POLAR_SC3_DBG", 0), scratch, env_int("POLAR_SC4_DISCARD", 1), u_packed, u_info, info_pos, k);
  count_launch();
  POLAR_CHECK_LAUNCH("sc4_kernel");
  return POLAR_OK;
}

}  // namespace

// n in [128, 2048].  warps = autonomous warps per SM (0 = as many as shared memory / tensor memory hold).
int launch_sc4(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed, float *u_info,
               const int32_t *info_pos, int k, int warps, cudaStream_t st) {
  switch (ilog2(n)) {
    case 7: return launch_sc4_t<7, 0>(logit, fmask, B, u_packed, u_info, info_pos, k, warps, st);
    case 8: return launch_sc4_t<8, 0>(logit, fmask, B, u_packed, u_info, info_pos, k, warps, st);
    case 9: return launch_sc4_t<9, 1>(logit, fmask, B, u_packed, u_info, info_pos, k, warps, st);
    case 10: return launch_sc4_t<10, 2>(logit, fmask, B, u_packed, u_info, info_pos, k, warps, st);
    case 11: return launch_sc4_t<11, 2>(logit, fmask, B, u_packed, u_info, info_pos, k, warps, st);
    default: return set_error(POLAR_EINVAL, "sc4: n=%d not supported by this mapping", n);
  }
}

}  // namespace polar

// debug only (not part of include/polar_b200.h): read and clear the phase timeline of warp 0 of CTA 0
extern "C" int polar_sc4_debug_read(unsigned long long *h_out8) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(h_out8, polar::g_sc4_dbg, sizeof(z)) != cudaSuccess) return POLAR_ECUDA;
  if (cudaMemcpyToSymbol(polar::g_sc4_dbg, z, sizeof(z)) != cudaSuccess) return POLAR_ECUDA;
  return POLAR_OK;
}
