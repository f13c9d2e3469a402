// polar_sc3.cu -- SC decoder, default mapping for n >= 128: compile-time tree geometry, virtual top stage,
// tensor memory as thread-private scratch, 64-leaf per-lane subtrees.
//
// Same algorithm and exact semantics as polar_sc.cu (x_run_sn_polar/polar/polar_sc.py:54-133, SURVEY.md
// Appendix A); what changes is where the LLR tree lives and who walks it.  The decoder is latency bound (the
// SC schedule is a chain of 2n-2 dependent steps), so throughput = codewords in flight per SM / latency of a
// codeword; the layout below is chosen to maximise the first and the per-lane code to minimise the second.
//   * A CTA (128 threads = one warpgroup) owns `cw` <= 32 codewords and walks the SC schedule once for all
//     of them (the schedule depends only on the frozen pattern).
//   * Stage m (channel) stays in global memory (L2 after the first pass).
//   * n >= 1024: stage m-1 is VIRTUAL -- never stored; the four stage m-2 nodes are computed straight from
//     the channel row (LL = f(f(c0,c2),f(c1,c3)), LR = g(f,f,b), RL = f(g,g), RR = g(g,g,b)) at the price of
//     n extra f/g per codeword and two more passes over the row.
//   * n >= 1024: stage m-2 lives in TENSOR MEMORY (tcgen05.st / tcgen05.ld, 32x32b: one TMEM lane per thread,
//     2^(m-4) columns per CTA).  The tensor cores are idle in this kernel, so their 256 KB per SM is free
//     storage: each thread keeps the (j, j + n/8) element pairs it produced and later consumes them itself
//     for the f and g steps into stage m-3, so the data never has to be visible to another thread.
//   * Stages 5 .. m-3 (n >= 1024) or 5 .. m-1 live in shared memory, one live node per stage; the row stride
//     is an odd number of float4, which makes both the cooperative and the lane-per-codeword accesses
//     bank-conflict free.  n = 1024: 1044 B per codeword instead of 4.2 KB -> 6 CTAs = 192 codewords per SM.
//   * Every 64-leaf subtree is decoded by ONE lane per codeword (rolled two-iteration loops per level,
//     speculative g at the leaves, see BetaTree in polar_common.cuh).  Only partial sums are produced;
//     the decisions are recovered once per codeword as u = T(x_hat).
//   * The warp that runs the 64-leaf subtrees rotates with the CTA's slot on the SM so that co-resident
//     CTAs keep their serial phases on different SM sub-partitions.
#include "polar_common.cuh"
#include "polar_internal.h"

namespace polar {

// phase timeline of CTA 0 (cycles), filled only when POLAR_SC3_DBG=1: tools/perf_probe.py reads it through
// polar_sc3_debug_read().  0 virtual steps, 1 g steps, 2 f steps, 3 64-leaf subtrees, 4 merges, 5 outputs,
// 6 total, 7 batches
__device__ unsigned long long g_sc3_dbg[8];

namespace {

constexpr unsigned FULLMASK = 0xFFFFFFFFu;
constexpr int NT = 128;                 // threads per CTA: one warpgroup = the 128 lanes of tensor memory
#ifndef SC3_MINB
#define SC3_MINB 5   // CTAs per SM the register allocation is sized for
#endif
#define SC3_T(slot)                                                                   \
  do {                                                                                \
    if (dbg && tid == 0 && blockIdx.x == 0) {                                         \
      const long long t__ = clock64(); g_sc3_dbg[slot] += (unsigned long long)(t__ - tlast); tlast = t__; \
    }                                                                                 \
  } while (0)

struct Sc3Layout {
  int nw, nws, n64, top, stride;
  size_t nz_off, llr_off, beta_off, total;
};
__host__ __device__ inline Sc3Layout sc3_layout(int m, bool tm, int cw) {
  Sc3Layout l;
  const int n = 1 << m;
  l.nw = n >> 5; l.nws = l.nw + 1; l.n64 = n >> 6;
  l.top = tm ? m - 3 : m - 1;                         // highest stage kept in shared memory (>= 6)
  l.stride = (2 << l.top) - 32 + 4;                   // floats per codeword row (stages 5..top); stride/4 is odd
  l.nz_off = (size_t)((l.nw * 4 + 15) / 16) * 16;
  l.llr_off = l.nz_off + (size_t)((2 * l.n64 + 15) / 16) * 16 + 16;   // +16: tensor-memory base address slot
  l.beta_off = l.llr_off + (size_t)cw * l.stride * 4;
  l.total = ((l.beta_off + (size_t)cw * l.nws * 4 + 15) / 16) * 16;
  return l;
}

PDEV float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
PDEV float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
PDEV void sts4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

PDEV float4 f4(const float4 a, const float4 b) {
  float4 o;
  o.x = f_minsum(a.x, b.x); o.y = f_minsum(a.y, b.y); o.z = f_minsum(a.z, b.z); o.w = f_minsum(a.w, b.w);
  return o;
}
// g on four consecutive elements; bit e of `bits` is the partial sum of element e
PDEV float4 f4neg(const float4 a, const float4 b) {   // f on logits (see f_minsum_neg)
  float4 o;
  o.x = f_minsum_neg(a.x, b.x); o.y = f_minsum_neg(a.y, b.y); o.z = f_minsum_neg(a.z, b.z); o.w = f_minsum_neg(a.w, b.w);
  return o;
}
PDEV float4 g4(const float4 a, const float4 b, const uint32_t bits) {
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

// ---- tensor memory as per-thread scratch (32x32b shape: thread t of the warpgroup owns TMEM lane t) -----
PDEV void tmem_st8(uint32_t taddr, const float4 a, const float4 b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(f2u(a.x)), "r"(f2u(a.y)), "r"(f2u(a.z)), "r"(f2u(a.w)), "r"(f2u(b.x)), "r"(f2u(b.y)),
               "r"(f2u(b.z)), "r"(f2u(b.w)) : "memory");
}
PDEV void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
PDEV void tmem_ld8(uint32_t taddr, float4 &a, float4 &b) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr) : "memory");
  // the registers are defined only after the wait; tying them to it keeps every use behind it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) :: "memory");
  a.x = u2f(r0); a.y = u2f(r1); a.z = u2f(r2); a.w = u2f(r3);
  b.x = u2f(r4); b.y = u2f(r5); b.z = u2f(r6); b.w = u2f(r7);
}

// ---- cooperative steps (all threads of the CTA) --------------------------------------------------
// stage S+1 -> S inside shared memory.  out[j] = f(a[j], a[j+H]) or g(a[j], a[j+H], beta_left[j]).
template <int S, bool IS_G>
PDEV void step_smem(float *L, const uint32_t *beta, int stride, int nws, int cw, int tid, int left_word) {
  constexpr int H = 1 << S, HQ = H >> 2;
  float *dst = L + (H - 32);
  const float *src = L + (2 * H - 32);
  const int items = cw * HQ;
#pragma unroll 2
  for (int it = tid; it < items; it += NT) {
    const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
    const float4 a = lds4(src + c * stride + j), b = lds4(src + c * stride + j + H);
    float4 o;
    if (IS_G) o = g4(a, b, beta[c * nws + left_word + (j >> 5)] >> (j & 31));
    else o = f4(a, b);
    sts4(dst + c * stride + j, o);
  }
}
template <int SMAX, bool IS_G>
PDEV void step_smem_any(int s, float *L, const uint32_t *beta, int stride, int nws, int cw, int tid, int left_word) {
  if constexpr (SMAX >= 6) {
    if (s == SMAX) step_smem<SMAX, IS_G>(L, beta, stride, nws, cw, tid, left_word);
    else step_smem_any<SMAX - 1, IS_G>(s, L, beta, stride, nws, cw, tid, left_word);
  }
}

// channel (global, stage M) -> stage M-1 in shared memory (n <= 512: everything fits in shared memory).
template <int M, bool IS_G>
__device__ __noinline__ void step_glob(const float *__restrict__ logit, int64_t cw0, int nvalid, float *L,
                                       const uint32_t *beta, int stride, int nws, int cw, int tid) {
  constexpr int N = 1 << M, H = N >> 1, HQ = H >> 2;
  constexpr int U = 4;        // items per round: 8 independent 128-bit loads in flight per thread
  float *dst = L + (H - 32);
  const int items = cw * HQ;
#pragma unroll 1
  for (int it0 = tid; it0 < items; it0 += U * NT) {
    float4 a[U], b[U];
#pragma unroll
    for (int r = 0; r < U; ++r) {
      const int it = min(it0 + r * NT, items - 1);
      const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
      const int cl = c < nvalid ? c : nvalid - 1;
      const float *row = logit + (cw0 + cl) * (int64_t)N + j;
      a[r] = ldg4(row); b[r] = ldg4(row + H);
    }
#pragma unroll
    for (int r = 0; r < U; ++r) {
      const int it = it0 + r * NT;
      if (it < items) {
        const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
        float4 o;
        if (IS_G) o = g4neg(a[r], b[r], beta[c * nws + (j >> 5)] >> (j & 31));
        else o = f4neg(a[r], b[r]);                       // f(-a,-b) == f(a,b) for min-sum
        sts4(dst + c * stride + j, o);
      }
    }
  }
}

// channel (global, stage M) -> stage M-2 in TENSOR MEMORY through the virtual stage M-1.
// kind = quarter of the codeword the target node covers: 0 LL, 1 LR, 2 RL, 3 RR (CTA-uniform).
// Work item = (codeword c, q): the float4 pair at elements 4q and 4q + H/2 of the stage M-2 node, i.e. exactly
// what one f/g of the next step consumes.  Item p = tid + NT*k goes to TMEM columns 8k..8k+7 of the thread.
template <int M>
__device__ __noinline__ void step_virt_tmem(const int kind, const float *__restrict__ logit, int64_t cw0, int nvalid,
                                            const uint32_t *beta, int nws, int cw, int tid, uint32_t tm_lane_base) {
  constexpr int N = 1 << M, H = N >> 2, PQ = H >> 3, HW = H >> 5;   // PQ pairs per codeword
  constexpr int KMAX = 32 * PQ / NT;
  const int pairs = cw * PQ;
  const bool right = kind >= 2, is_g = kind & 1;
  const int gw = (kind == 3) ? 2 * HW : 0;
#pragma unroll 1
  for (int k = 0; k < KMAX; ++k) {
    const int p = min(tid + NT * k, pairs - 1);
    const int c = (int)((unsigned)p / (unsigned)PQ), q = (int)((unsigned)p % (unsigned)PQ);
    const int cl = c < nvalid ? c : nvalid - 1;
    const float *row = logit + (cw0 + cl) * (int64_t)N;
    float4 o[2];
    float4 c0[2], c1[2], c2[2], c3[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float *r = row + 4 * q + e * (H / 2);
      c0[e] = ldg4(r); c1[e] = ldg4(r + H); c2[e] = ldg4(r + 2 * H); c3[e] = ldg4(r + 3 * H);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 4 * q + e * (H / 2);
      const uint32_t *bw = beta + c * nws + (j >> 5);
      const int sh = j & 31;
      float4 y0, y1;
      if (!right) {              // left half of the codeword: stage M-1 node = f(channel)
        y0 = f4neg(c0[e], c2[e]); y1 = f4neg(c1[e], c3[e]);
      } else {                   // right half: stage M-1 node = g(channel, beta of the left half)
        y0 = g4neg(c0[e], c2[e], bw[0] >> sh); y1 = g4neg(c1[e], c3[e], bw[HW] >> sh);
      }
      if (!is_g) o[e] = f4(y0, y1);
      else o[e] = g4(y0, y1, bw[gw] >> sh);
    }
    tmem_st8(tm_lane_base + 8 * k, o[0], o[1]);
  }
  tmem_wait_st();
}

// stage M-2 (tensor memory, thread-private pairs) -> stage M-3 in shared memory.
template <int M, bool IS_G>
PDEV void step_tmem(float *L, const uint32_t *beta, int stride, int nws, int cw, int tid, uint32_t tm_lane_base,
                    int left_word) {
  constexpr int N = 1 << M, H = N >> 3, PQ = H >> 2;   // H outputs per codeword = PQ float4
  constexpr int KMAX = 32 * PQ / NT;
  float *dst = L + (H - 32);
  const int pairs = cw * PQ;
#pragma unroll 2
  for (int k = 0; k < KMAX; ++k) {
    const int p = tid + NT * k;
    float4 a, b;
    tmem_ld8(tm_lane_base + 8 * k, a, b);          // warp-collective: executed by every lane, valid or not
    if (p < pairs) {
      const int c = (int)((unsigned)p / (unsigned)PQ), j = (int)((unsigned)p % (unsigned)PQ) << 2;
      float4 o;
      if (IS_G) o = g4(a, b, beta[c * nws + left_word + (j >> 5)] >> (j & 31));
      else o = f4(a, b);
      sts4(dst + c * stride + j, o);
    }
  }
}

// ask the L2 for the channel rows of the CTA's next batch (one 1 KB bulk prefetch per thread and round)
PDEV void prefetch_rows_l2(const float *base, size_t bytes, int tid) {
  const char *p = reinterpret_cast<const char *>(base);
  for (size_t off = (size_t)tid * 1024; off < bytes; off += (size_t)NT * 1024) {
    const unsigned sz = (unsigned)((bytes - off) < 1024 ? (bytes - off) : 1024);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(sz & ~15u) : "memory");
  }
}

// ---- one lane per codeword: the 64-leaf subtree below a stage-6 node held in shared memory -------
// `row` is the lane's own shared-memory row: stage 5 at [0,32), stage 6 at [32,96).  Stage 5 is staged through
// shared memory (lane-private, so no synchronisation) instead of 32 live registers.
PDEV uint32_t tree32(const float *x5, uint32_t fm) {
  if (fm == 0) {           // rate-1: hard decisions, unless an LLR is exactly 0 (then the recursion below)
    uint32_t hd = 0;
    float mn = 1.0f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = lds4(x5 + 4 * q);
      hd |= ((f2u(v.x) >> 31) << (4 * q)) | ((f2u(v.y) >> 31) << (4 * q + 1)) | ((f2u(v.z) >> 31) << (4 * q + 2)) |
            ((f2u(v.w) >> 31) << (4 * q + 3));
      mn = fminf(mn, fminf(fminf(fabsf(v.x), fabsf(v.y)), fminf(fabsf(v.z), fabsf(v.w))));
    }
    if (mn != 0.0f) return hd;
  }
  uint32_t bl = 0, bc = 0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    const uint32_t fmc = h ? (fm >> 16) : (fm & 0xFFFFu);
    if (fmc == 0xFFFFu) { bc = 0; continue; }
    float y[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 a = lds4(x5 + 4 * q), b = lds4(x5 + 16 + 4 * q);
      const float4 o = h ? g4(a, b, bl >> (4 * q)) : f4(a, b);
      y[4 * q] = o.x; y[4 * q + 1] = o.y; y[4 * q + 2] = o.z; y[4 * q + 3] = o.w;
    }
    bc = BetaTree<4>::run(y, fmc);
    if (h == 0) bl = bc;
  }
  return (bl ^ bc) | (bc << 16);
}
__device__ __noinline__ uint2 bottom64(float *row, uint32_t fm0, uint32_t fm1) {
  uint32_t bl = 0, bc = 0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {      // rolled: one copy of the 32-leaf subtree code
    const uint32_t fmc = h ? fm1 : fm0;
    if (fmc == FULLMASK) { bc = 0; continue; }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = lds4(row + 32 + 4 * q), b = lds4(row + 64 + 4 * q);
      sts4(row + 4 * q, h ? g4(a, b, bl >> (4 * q)) : f4(a, b));
    }
    bc = tree32(row, fmc);
    if (h == 0) bl = bc;
  }
  return make_uint2(bl ^ bc, bc);
}

template <int M, bool TM>
__global__ void __launch_bounds__(NT, SC3_MINB) sc3_kernel(const float *__restrict__ logit, const uint32_t *__restrict__ fmask_g,
                                                    int cw, int64_t B, int64_t nbatches, int bw_div, int l2_prefetch, int dbg,
                                                    uint32_t *__restrict__ u_packed, float *__restrict__ u_info,
                                                    const int32_t *__restrict__ info_pos, int k) {
  static_assert(TM ? (M >= 10) : (M >= 7), "sc3: stage 6 must exist in shared memory");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = 1 << M, NW = N >> 5, NWS = NW + 1, N64 = N >> 6, TOP = TM ? M - 3 : M - 1;
  constexpr int TM_COLS = TM ? (1 << (M - 4)) : 32;         // 32 codewords x 2^(M-2) floats / 128 lanes
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Sc3Layout lay = sc3_layout(M, TM, cw);
  const int stride = lay.stride;
  uint32_t *fmask = reinterpret_cast<uint32_t *>(smem_raw);
  unsigned char *nz = smem_raw + lay.nz_off;    // nz[(N64 >> lv) + (i >> lv)]: node of 2^lv 64-blocks at block i is rate-0
  uint32_t *tm_slot = reinterpret_cast<uint32_t *>(smem_raw + lay.llr_off - 16);
  float *L = reinterpret_cast<float *>(smem_raw + lay.llr_off);
  uint32_t *beta = reinterpret_cast<uint32_t *>(smem_raw + lay.beta_off);
  const int bw = (int)((blockIdx.x / (unsigned)bw_div) % (NT / 32));   // warp that runs the 64-leaf subtrees
  const int pf_at = (N64 * l2_prefetch) >> 3;                          // prefetch point in eighths of the codeword

  uint32_t tm_lane_base = 0;
  if (TM) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"((uint32_t)__cvta_generic_to_shared(tm_slot)), "n"(TM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  for (int i = tid; i < NW; i += NT) fmask[i] = __ldg(fmask_g + i);
  __syncthreads();
  if (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tm_lane_base = *tm_slot + ((uint32_t)(warp * 32) << 16);
  }
  for (int i = tid; i < N64; i += NT) nz[N64 + i] = (fmask[2 * i] & fmask[2 * i + 1]) == FULLMASK;
  __syncthreads();
  if (tid == 0)
    for (int idx = N64 - 1; idx >= 1; --idx) nz[idx] = nz[2 * idx] & nz[2 * idx + 1];
  __syncthreads();

  long long tlast = clock64();
  const long long tstart = tlast;
  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int64_t cw0 = batch * cw;
    const int nvalid = (int)((B - cw0) < (int64_t)cw ? (B - cw0) : (int64_t)cw);
    int i = 0;                                   // current 64-leaf block
    bool prefetched = !l2_prefetch;
    while (i < N64) {
      // node entered at block i: the root, or the right child whose left sibling just finished
      const int S = (i == 0) ? M : 6 + (__ffs(i) - 1);
      int s = S;
      if (!prefetched && i >= pf_at) {           // late in the batch: pull the next batch's rows into the L2
        prefetched = true;
        const int64_t nb = batch + gridDim.x;
        if (nb < nbatches) {
          const int64_t r0 = nb * cw, r1 = (r0 + cw < B) ? r0 + cw : B;
          prefetch_rows_l2(logit + r0 * (int64_t)N, (size_t)(r1 - r0) * N * 4, tid);
        }
      }
      bool zeroed = nz[(N64 >> (S - 6)) + (i >> (S - 6))] != 0;
      if (!zeroed && S < M && !(TM && S == M - 1)) {
        // g step into (S, i) from its parent at stage S+1; the left sibling's beta starts at word 2*(i - 2^(S-6))
        const int left_word = 2 * (i - (1 << (S - 6)));
        if (TM && S == M - 2) {
          step_virt_tmem<M>(i < N64 / 2 ? 1 : 3, logit, cw0, nvalid, beta, NWS, cw, tid, tm_lane_base);
        } else if (TM && S == M - 3) {
          step_tmem<M, true>(L, beta, stride, NWS, cw, tid, tm_lane_base, left_word);
        } else if (!TM && S == M - 1) {
          step_glob<M, true>(logit, cw0, nvalid, L, beta, stride, NWS, cw, tid);
        } else {
          step_smem_any<TOP - 1, true>(S, L, beta, stride, NWS, cw, tid, left_word);
        }
        __syncthreads();
        SC3_T((TM && S == M - 2) ? 0 : 1);
      }
      while (!zeroed && s > 6) {
        if (nz[(N64 >> (s - 7)) + (i >> (s - 7))]) { zeroed = true; --s; break; }   // left child is rate-0
        if (TM && s == M) { --s; continue; }                                           // virtual stage: nothing stored
        if (TM && s == M - 1) {
          step_virt_tmem<M>(i < N64 / 2 ? 0 : 2, logit, cw0, nvalid, beta, NWS, cw, tid, tm_lane_base);
        } else if (TM && s == M - 2) {
          step_tmem<M, false>(L, beta, stride, NWS, cw, tid, tm_lane_base, 0);
        } else if (!TM && s == M) {
          step_glob<M, false>(logit, cw0, nvalid, L, beta, stride, NWS, cw, tid);
        } else {
          step_smem_any<TOP - 1, false>(s - 1, L, beta, stride, NWS, cw, tid, 0);
        }
        __syncthreads();
        SC3_T((TM && s == M - 1) ? 0 : 2);
        --s;
      }
      const int lv0 = s - 6;                     // the finished node covers 2^lv0 64-blocks starting at i
      if (zeroed) {
        const int nwd = 2 << lv0;
        for (int q = tid; q < cw * nwd; q += NT) {
          const int c = q >> (lv0 + 1), w = q & (nwd - 1);
          beta[c * NWS + 2 * i + w] = 0u;
        }
      } else if (warp == bw && lane < cw) {
        const uint2 b = bottom64(L + lane * stride, fmask[2 * i], fmask[2 * i + 1]);
        uint32_t *bp = beta + lane * NWS + 2 * i;
        bp[0] = b.x; bp[1] = b.y;
      }
      __syncthreads();
      SC3_T(3);
      {  // merge partial sums upward while the finished node is a right child: [bl ^ br, br] (polar_sc.py:83-89)
        int lv = lv0, a = i;
        while (lv < M - 6 && ((a >> lv) & 1)) {
          const int nwd = 2 << lv, left = a - (1 << lv);
          for (int q = tid; q < cw * nwd; q += NT) {
            const int c = q >> (lv + 1), w = q & (nwd - 1);
            beta[c * NWS + 2 * left + w] ^= beta[c * NWS + 2 * a + w];
          }
          __syncthreads();
          a = left; ++lv;
        }
      }
      SC3_T(4);
      i += 1 << lv0;
    }
    // beta now holds the re-encoded codeword x_hat of every codeword; the decisions are u = T(x_hat)
    // (my_sn/fec/polar/enc.py:85-96 is an involution): 5 stages inside each word, M-5 across words.
    for (int q = tid; q < cw * NW; q += NT) {
      const int c = q / NW, w = q % NW;
      beta[c * NWS + w] = ptransform<5>(beta[c * NWS + w]);
    }
    __syncthreads();
#pragma unroll 1
    for (int st = 0; st < M - 5; ++st) {
      for (int q = tid; q < cw * (NW / 2); q += NT) {
        const int c = q / (NW / 2), r = q % (NW / 2);
        const int w = ((r >> st) << (st + 1)) | (r & ((1 << st) - 1));     // word index with bit st clear
        beta[c * NWS + w] ^= beta[c * NWS + w + (1 << st)];
      }
      __syncthreads();
    }
    if (u_packed) {
      for (int q = tid; q < cw * NW; q += NT) {
        const int c = q / NW, w = q % NW;
        if (c < nvalid) u_packed[(cw0 + c) * NW + w] = beta[c * NWS + w];
      }
    }
    if (u_info) {
      for (int q = tid; q < nvalid * k; q += NT) {
        const int c = q / k, t = q - c * k;
        const int p = __ldg(info_pos + t);
        u_info[(cw0 + c) * (int64_t)k + t] = (float)((beta[c * NWS + (p >> 5)] >> (p & 31)) & 1u);
      }
    }
    __syncthreads();
    SC3_T(5);
    if (dbg && tid == 0 && blockIdx.x == 0) g_sc3_dbg[7] += 1;
  }
  if (dbg && tid == 0 && blockIdx.x == 0) g_sc3_dbg[6] += (unsigned long long)(clock64() - tstart);
  if (TM) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tm_slot), "n"(TM_COLS) : "memory");
  }
}

template <int M, bool TM>
int launch_sc3_t(const float *logit, const uint32_t *fmask, int64_t B, uint32_t *u_packed, float *u_info,
                 const int32_t *info_pos, int k, int cw, int ctas_per_sm, cudaStream_t st) {
  const int max_smem = device_max_smem_optin();
  if (cw > 32) cw = 32;
  if (cw < 1) cw = 1;
  while (cw > 1 && sc3_layout(M, TM, cw).total > (size_t)max_smem) --cw;
  if (ctas_per_sm > 0)   // shrink the codeword group until `ctas_per_sm` CTAs fit in the 228 KB of an SM
    while (cw > 1 && (sc3_layout(M, TM, cw).total + 1024) * (size_t)ctas_per_sm > (size_t)228 * 1024) --cw;
  const Sc3Layout lay = sc3_layout(M, TM, cw);
  if (lay.total > (size_t)max_smem) return set_error(POLAR_ENOMEM, "sc: n=%d needs %zu B shared memory per CTA", 1 << M, lay.total);
  auto kern = sc3_kernel<M, TM>;
  POLAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
  int occ = 0;
  POLAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, lay.total));
  if (env_int("POLAR_SC3_DBG", 0)) {
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, kern);
    fprintf(stderr, "sc3: m=%d cw=%d smem=%zu occ=%d regs=%d local=%zu static_smem=%zu\n", M, cw, lay.total, occ, fa.numRegs,
            fa.localSizeBytes, fa.sharedSizeBytes);
  }
  if (env_int("POLAR_SC3_FORCE_OCC", 0) > 0) occ = env_int("POLAR_SC3_FORCE_OCC", 0);
  if (occ < 1) occ = 1;
  if (TM) {                                   // the occupancy calculator does not know about tensor memory:
    const int tm_cols = 1 << (M - 4);         // never schedule more CTAs than 512 columns can serve
    if (occ > 512 / tm_cols) occ = 512 / tm_cols;
  }
  if (ctas_per_sm > 0 && occ > ctas_per_sm) occ = ctas_per_sm;
  const int sms = device_sm_count();
  const int64_t nbatches = (B + cw - 1) / cw;
  int64_t grid = nbatches;
  if (grid > (int64_t)sms * occ) grid = (int64_t)sms * occ;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, NT, lay.total, st>>>(logit, fmask, cw, B, nbatches, env_int("POLAR_SC3_BWDIV", sms),
                                              env_int("POLAR_SC3_PREFETCH", 6), env_int("POLAR_SC3_DBG", 0), u_packed,
                                              u_info, info_pos, k);
  count_launch();
  POLAR_CHECK_LAUNCH("sc3_kernel");
  return POLAR_OK;
}

}  // namespace

// n in [128, 8192].  cw = codewords per CTA (<= 32); ctas = CTAs per SM (0 = as many as fit).
int launch_sc3(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed, float *u_info,
               const int32_t *info_pos, int k, int cw, int ctas, cudaStream_t st) {
  switch (ilog2(n)) {
    case 7: return launch_sc3_t<7, false>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 8: return launch_sc3_t<8, false>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 9: return launch_sc3_t<9, false>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 10: return launch_sc3_t<10, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 11: return launch_sc3_t<11, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 12: return launch_sc3_t<12, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 13: return launch_sc3_t<13, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    default: return set_error(POLAR_EINVAL, "sc3: n=%d not supported by this mapping", n);
  }
}

}  // namespace polar

// debug only (not part of include/polar_b200.h): read and clear the phase timeline of CTA 0
extern "C" int polar_sc3_debug_read(unsigned long long *h_out8) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(h_out8, polar::g_sc3_dbg, sizeof(z)) != cudaSuccess) return POLAR_ECUDA;
  if (cudaMemcpyToSymbol(polar::g_sc3_dbg, z, sizeof(z)) != cudaSuccess) return POLAR_ECUDA;
  return POLAR_OK;
}
