// polar_sc3.cu -- SC decoder, third mapping (default for n >= 128): compile-time tree geometry,
// virtual top stage, 64-leaf register subtrees.
//
// Same algorithm and exact semantics as polar_sc.cu (x_run_sn_polar/polar/polar_sc.py:54-133, SURVEY.md
// Appendix A); what changes is where the LLR tree lives and who walks it:
//   * A CTA owns `cw` <= 32 codewords and walks the SC schedule once for all of them (the schedule
//     depends only on the frozen pattern).
//   * Stage m (channel) stays in global memory.  For n >= 1024 stage m-1 is VIRTUAL: it is never
//     stored; the four stage m-2 nodes are computed straight from the channel row
//     (LL = f(f(c0,c2),f(c1,c3)), LR = g(f,f,b), RL = f(g,g), RR = g(g,g,b)).  That halves the shared
//     memory per codeword (n=1024: 4.2 KB -> 2.1 KB), which doubles the codewords in flight per SM;
//     the price is n extra f/g per codeword and two more passes over the channel row (L2 hits).
//   * Stages 6 .. top live in shared memory (one live node per stage, row stride = odd number of
//     float4 so that both access patterns below are bank-conflict free) and are updated by all warps,
//     four elements per thread with 128-bit LDS/STS, the stage being a template parameter.
//   * Every 64-leaf subtree is decoded by ONE lane per codeword: the lane reads its stage-6 node twice
//     (f pass, then g pass) and runs two 32-leaf register subtrees (SubTree<5>).  No CTA barrier and no
//     shared-memory round trip below stage 6.
//   * The warp that runs the 64-leaf subtrees rotates with the CTA's slot on the SM so that co-resident
//     CTAs keep their serial phases on different SM sub-partitions.
#include "polar_common.cuh"
#include "polar_internal.h"

namespace polar {

namespace {

constexpr unsigned FULLMASK = 0xFFFFFFFFu;

struct Sc3Layout {
  int n, m, nw, nws, n64, top, stride;
  size_t nz_off, llr_off, beta_off, uo_off, total;
};
__host__ __device__ inline Sc3Layout sc3_layout(int m, bool virt, int cw) {
  Sc3Layout l;
  l.n = 1 << m; l.m = m; l.nw = l.n >> 5; l.nws = l.nw + 1; l.n64 = l.n >> 6;
  l.top = virt ? m - 2 : m - 1;                       // highest stage kept in shared memory (>= 6)
  l.stride = (2 << l.top) - 64 + 4;                   // floats per codeword row; stride/4 is odd
  l.nz_off = (size_t)((l.nw * 4 + 15) / 16) * 16;
  l.llr_off = l.nz_off + (size_t)((2 * l.n64 + 15) / 16) * 16;
  l.beta_off = l.llr_off + (size_t)cw * l.stride * 4;
  l.uo_off = l.beta_off + (size_t)cw * l.nws * 4;
  l.total = ((l.uo_off + (size_t)cw * l.nws * 4 + 15) / 16) * 16;
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

// ---- cooperative steps (all threads of the CTA) --------------------------------------------------
// stage S+1 -> S inside shared memory.  out[j] = f(a[j], a[j+H]) or g(a[j], a[j+H], beta_left[j]).
template <int NT, int S, bool IS_G>
PDEV void step_smem(float *L, const uint32_t *beta, int stride, int nws, int cw, int tid, int left_word) {
  constexpr int H = 1 << S, HQ = H >> 2;
  float *dst = L + (H - 64);
  const float *src = L + (2 * H - 64);
  const int items = cw * HQ;
#pragma unroll 1
  for (int it = tid; it < items; it += NT) {
    const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
    const float4 a = lds4(src + c * stride + j), b = lds4(src + c * stride + j + H);
    float4 o;
    if (IS_G) o = g4(a, b, beta[c * nws + left_word + (j >> 5)] >> (j & 31));
    else o = f4(a, b);
    sts4(dst + c * stride + j, o);
  }
}
template <int NT, int SMAX, bool IS_G>
PDEV void step_smem_any(int s, float *L, const uint32_t *beta, int stride, int nws, int cw, int tid, int left_word) {
  if constexpr (SMAX >= 6) {
    if (s == SMAX) step_smem<NT, SMAX, IS_G>(L, beta, stride, nws, cw, tid, left_word);
    else step_smem_any<NT, SMAX - 1, IS_G>(s, L, beta, stride, nws, cw, tid, left_word);
  }
}

// channel (global, stage M) -> stage M-1 in shared memory (non-virtual layouts).
template <int NT, int M, bool IS_G>
PDEV void step_glob(const float *__restrict__ logit, int64_t cw0, int nvalid, float *L, const uint32_t *beta,
                    int stride, int nws, int cw, int tid) {
  constexpr int N = 1 << M, H = N >> 1, HQ = H >> 2;
  float *dst = L + (H - 64);
  const int items = cw * HQ;
#pragma unroll 1
  for (int it = tid; it < items; it += NT) {
    const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
    const int cl = c < nvalid ? c : nvalid - 1;
    const float *row = logit + (cw0 + cl) * (int64_t)N;
    const float4 a = ldg4(row + j), b = ldg4(row + j + H);
    float4 o;
    if (IS_G) o = g4neg(a, b, beta[c * nws + (j >> 5)] >> (j & 31));
    else o = f4(a, b);                          // f(-a,-b) == f(a,b)
    sts4(dst + c * stride + j, o);
  }
}

// channel (global, stage M) -> stage M-2 in shared memory through the virtual stage M-1.
// kind = quarter of the codeword the target node covers: 0 LL, 1 LR, 2 RL, 3 RR (CTA-uniform; one copy of
// the code serves the four calls per codeword -- instruction-cache footprint matters more than the selects).
template <int NT, int M>
PDEV void step_virt(const int kind, const float *__restrict__ logit, int64_t cw0, int nvalid, float *L,
                    const uint32_t *beta, int stride, int nws, int cw, int tid) {
  constexpr int N = 1 << M, H = N >> 2, HQ = H >> 2, HW = H >> 5;
  float *dst = L + (H - 64);
  const int items = cw * HQ;
  const bool right = kind >= 2, is_g = kind & 1;
  const int gw = (kind == 3) ? 2 * HW : 0;
#pragma unroll 1
  for (int it = tid; it < items; it += NT) {
    const int c = (int)((unsigned)it / (unsigned)HQ), j = (int)((unsigned)it % (unsigned)HQ) << 2;
    const int cl = c < nvalid ? c : nvalid - 1;
    const float *row = logit + (cw0 + cl) * (int64_t)N;
    const float4 c0 = ldg4(row + j), c1 = ldg4(row + j + H), c2 = ldg4(row + j + 2 * H), c3 = ldg4(row + j + 3 * H);
    const uint32_t *bw = beta + c * nws + (j >> 5);
    const int sh = j & 31;
    float4 y0, y1, o;
    if (!right) {              // left half of the codeword: stage M-1 node = f(channel)
      y0 = f4(c0, c2); y1 = f4(c1, c3);
    } else {                   // right half: stage M-1 node = g(channel, beta of the left half)
      y0 = g4neg(c0, c2, bw[0] >> sh); y1 = g4neg(c1, c3, bw[HW] >> sh);
    }
    if (!is_g) o = f4(y0, y1);
    else o = g4(y0, y1, bw[gw] >> sh);
    sts4(dst + c * stride + j, o);
  }
}

// ---- one lane per codeword: the 64-leaf subtree below a stage-6 node held in shared memory -------
PDEV void bottom64(const float *node, uint32_t fm0, uint32_t fm1, uint32_t &b0, uint32_t &b1, uint32_t &u0,
                   uint32_t &u1) {
  uint32_t bl = 0, ul = 0, bc = 0, uc = 0;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {      // rolled: one copy of the 32-leaf subtree code (see RollTree)
    const uint32_t fmc = h ? fm1 : fm0;
    if (fmc == FULLMASK) { bc = 0; uc = 0; continue; }
    float x[32];
    const uint32_t gm = h ? 0xFFFFFFFFu : 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = lds4(node + 4 * q), b = lds4(node + 32 + 4 * q);
      if (h == 0) {
        x[4 * q] = f_minsum(a.x, b.x); x[4 * q + 1] = f_minsum(a.y, b.y);
        x[4 * q + 2] = f_minsum(a.z, b.z); x[4 * q + 3] = f_minsum(a.w, b.w);
      } else {
        x[4 * q] = g_minsum(a.x, b.x, (bl << (31 - 4 * q)) & 0x80000000u);
        x[4 * q + 1] = g_minsum(a.y, b.y, (bl << (30 - 4 * q)) & 0x80000000u);
        x[4 * q + 2] = g_minsum(a.z, b.z, (bl << (29 - 4 * q)) & 0x80000000u);
        x[4 * q + 3] = g_minsum(a.w, b.w, (bl << (28 - 4 * q)) & 0x80000000u);
      }
    }
    (void)gm;
    bc = RollTree<5>::run(x, fmc, uc);
    if (h == 0) { bl = bc; ul = uc; }
  }
  b0 = bl ^ bc; b1 = bc; u0 = ul; u1 = uc;
}

template <int M, int NT, bool VIRT>
__global__ void __launch_bounds__(NT, NT == 128 ? 4 : 2) sc3_kernel(const float *__restrict__ logit, const uint32_t *__restrict__ fmask_g,
                                                 int cw, int64_t B, int64_t nbatches, int bw_div,
                                                 uint32_t *__restrict__ u_packed, float *__restrict__ u_info,
                                                 const int32_t *__restrict__ info_pos, int k) {
  static_assert(VIRT ? (M >= 8) : (M >= 7), "sc3: stage 6 must exist in shared memory");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int N = 1 << M, NW = N >> 5, NWS = NW + 1, N64 = N >> 6, TOP = VIRT ? M - 2 : M - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Sc3Layout lay = sc3_layout(M, VIRT, cw);
  const int stride = lay.stride;
  uint32_t *fmask = reinterpret_cast<uint32_t *>(smem_raw);
  unsigned char *nz = smem_raw + lay.nz_off;    // nz[(N64 >> lv) + (i >> lv)]: node of 2^lv 64-blocks at block i is rate-0
  float *L = reinterpret_cast<float *>(smem_raw + lay.llr_off);
  uint32_t *beta = reinterpret_cast<uint32_t *>(smem_raw + lay.beta_off);
  uint32_t *uo = reinterpret_cast<uint32_t *>(smem_raw + lay.uo_off);
  const int bw = (int)((blockIdx.x / (unsigned)bw_div) % (NT / 32));   // warp that runs the 64-leaf subtrees

  for (int i = tid; i < NW; i += NT) fmask[i] = __ldg(fmask_g + i);
  __syncthreads();
  for (int i = tid; i < N64; i += NT) nz[N64 + i] = (fmask[2 * i] & fmask[2 * i + 1]) == FULLMASK;
  __syncthreads();
  if (tid == 0)
    for (int idx = N64 - 1; idx >= 1; --idx) nz[idx] = nz[2 * idx] & nz[2 * idx + 1];
  __syncthreads();

  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int64_t cw0 = batch * cw;
    const int nvalid = (int)((B - cw0) < (int64_t)cw ? (B - cw0) : (int64_t)cw);
    int i = 0;                                   // current 64-leaf block
    while (i < N64) {
      // node entered at block i: the root, or the right child whose left sibling just finished
      const int S = (i == 0) ? M : 6 + (__ffs(i) - 1);
      int s = S;
      bool zeroed = nz[(N64 >> (S - 6)) + (i >> (S - 6))] != 0;
      if (!zeroed && S < M && !(VIRT && S == M - 1)) {
        // g step into (S, i) from its parent at stage S+1; left sibling's beta starts at word 2*(i - 2^(S-6))
        if (VIRT && S == M - 2) {
          step_virt<NT, M>(i < N64 / 2 ? 1 : 3, logit, cw0, nvalid, L, beta, stride, NWS, cw, tid);
        } else if (!VIRT && S == M - 1) {
          step_glob<NT, M, true>(logit, cw0, nvalid, L, beta, stride, NWS, cw, tid);
        } else {
          step_smem_any<NT, TOP - 1, true>(S, L, beta, stride, NWS, cw, tid, 2 * (i - (1 << (S - 6))));
        }
        __syncthreads();
      }
      while (!zeroed && s > 6) {
        if (nz[(N64 >> (s - 7)) + (i >> (s - 7))]) { zeroed = true; --s; break; }   // left child is rate-0
        if (VIRT && s == M) { --s; continue; }                                         // virtual stage: nothing stored
        if (VIRT && s == M - 1) {
          step_virt<NT, M>(i < N64 / 2 ? 0 : 2, logit, cw0, nvalid, L, beta, stride, NWS, cw, tid);
        } else if (!VIRT && s == M) {
          step_glob<NT, M, false>(logit, cw0, nvalid, L, beta, stride, NWS, cw, tid);
        } else {
          step_smem_any<NT, TOP - 1, false>(s - 1, L, beta, stride, NWS, cw, tid, 0);
        }
        __syncthreads();
        --s;
      }
      const int lv0 = s - 6;                     // the finished node covers 2^lv0 64-blocks starting at i
      if (zeroed) {
        const int nwd = 2 << lv0;
        for (int q = tid; q < cw * nwd; q += NT) {
          const int c = q >> (lv0 + 1), w = q & (nwd - 1);
          beta[c * NWS + 2 * i + w] = 0u; uo[c * NWS + 2 * i + w] = 0u;
        }
      } else if (warp == bw && lane < cw) {
        uint32_t b0, b1, u0, u1;
        bottom64(L + lane * stride, fmask[2 * i], fmask[2 * i + 1], b0, b1, u0, u1);
        uint32_t *bp = beta + lane * NWS + 2 * i, *up = uo + lane * NWS + 2 * i;
        bp[0] = b0; bp[1] = b1; up[0] = u0; up[1] = u1;
      }
      __syncthreads();
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
      i += 1 << lv0;
    }
    if (u_packed) {
      for (int q = tid; q < cw * NW; q += NT) {
        const int c = q / NW, w = q % NW;
        if (c < nvalid) u_packed[(cw0 + c) * NW + w] = uo[c * NWS + w];
      }
    }
    if (u_info) {
      for (int q = tid; q < nvalid * k; q += NT) {
        const int c = q / k, t = q - c * k;
        const int p = __ldg(info_pos + t);
        u_info[(cw0 + c) * (int64_t)k + t] = (float)((uo[c * NWS + (p >> 5)] >> (p & 31)) & 1u);
      }
    }
    __syncthreads();
  }
}

template <int M, int NT, bool VIRT>
int launch_sc3_t(const float *logit, const uint32_t *fmask, int64_t B, uint32_t *u_packed, float *u_info,
                 const int32_t *info_pos, int k, int cw, int ctas_per_sm, cudaStream_t st) {
  const int max_smem = device_max_smem_optin();
  if (cw > 32) cw = 32;
  if (cw < 1) cw = 1;
  while (cw > 1 && sc3_layout(M, VIRT, cw).total > (size_t)max_smem) --cw;
  if (ctas_per_sm > 0)   // shrink the codeword group until `ctas_per_sm` CTAs fit in the 228 KB of an SM
    while (cw > 1 && (sc3_layout(M, VIRT, cw).total + 1024) * (size_t)ctas_per_sm > (size_t)228 * 1024) --cw;
  const Sc3Layout lay = sc3_layout(M, VIRT, cw);
  if (lay.total > (size_t)max_smem) return set_error(POLAR_ENOMEM, "sc: n=%d needs %zu B shared memory per CTA", 1 << M, lay.total);
  auto kern = sc3_kernel<M, NT, VIRT>;
  POLAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
  int occ = 0;
  POLAR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, lay.total));
  if (occ < 1) occ = 1;
  if (ctas_per_sm > 0 && occ > ctas_per_sm) occ = ctas_per_sm;
  const int sms = device_sm_count();
  const int64_t nbatches = (B + cw - 1) / cw;
  int64_t grid = nbatches;
  if (grid > (int64_t)sms * occ) grid = (int64_t)sms * occ;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, NT, lay.total, st>>>(logit, fmask, cw, B, nbatches, env_int("POLAR_SC3_BWDIV", sms), u_packed,
                                              u_info, info_pos, k);
  count_launch();
  POLAR_CHECK_LAUNCH("sc3_kernel");
  return POLAR_OK;
}

template <int NT>
int launch_sc3_nt(const float *logit, const uint32_t *fmask, int m, int64_t B, uint32_t *u_packed, float *u_info,
                  const int32_t *info_pos, int k, int cw, int ctas, cudaStream_t st) {
  switch (m) {
    case 7: return launch_sc3_t<7, NT, false>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 8: return launch_sc3_t<8, NT, false>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 9: return launch_sc3_t<9, NT, false>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 10: return launch_sc3_t<10, NT, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 11: return launch_sc3_t<11, NT, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 12: return launch_sc3_t<12, NT, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    case 13: return launch_sc3_t<13, NT, true>(logit, fmask, B, u_packed, u_info, info_pos, k, cw, ctas, st);
    default: return set_error(POLAR_EINVAL, "sc3: n=%d not supported by this mapping", 1 << m);
  }
}

}  // namespace

// n in [128, 8192].  threads in {128, 256}; cw = codewords per CTA (<= 32); ctas = CTAs per SM (0 = as many as fit).
int launch_sc3(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed, float *u_info,
               const int32_t *info_pos, int k, int cw, int threads, int ctas, cudaStream_t st) {
  const int m = ilog2(n);
  if (threads <= 128) return launch_sc3_nt<128>(logit, fmask, m, B, u_packed, u_info, info_pos, k, cw, ctas, st);
  return launch_sc3_nt<256>(logit, fmask, m, B, u_packed, u_info, info_pos, k, cw, ctas, st);
}

}  // namespace polar
