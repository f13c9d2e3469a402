// polar_sc.cu -- SC decoder for sm_100a (fp32 min-sum, bit-exact with the reference).
//
// Replaces x_run_sn_polar/polar/polar_sc.py:54-133 (recursion, f/g, leaf rule, partial sums, info
// gather).  Semantics: SURVEY.md Appendix A.
//
// Mapping (DESIGN.md "SC kernel"):
//   * A warp decodes CW codewords in lock-step (the SC schedule depends only on the frozen pattern,
//     which all codewords share).
//   * Stages >= 5 of the LLR tree (node width >= 32) live in shared memory, one buffer of 2^s floats
//     per stage and codeword (only one node per stage is live at a time: n-32 floats per codeword).
//     f/g over these wide nodes are done cooperatively, 4 elements per lane with 128-bit smem
//     accesses; the channel stage (s = m) is read straight from global memory (coalesced float4,
//     second read served by L2) and negated on the fly (logit -> LLR, polar_sc.py:122).
//   * Every 32-leaf subtree (stages 4..0: 160 f/g updates, 32 leaf decisions, 31 partial-sum merges)
//     is decoded by ONE thread per codeword entirely in registers (SubTree<5>, compile-time indices).
//   * Partial sums and decisions are bit-packed words; merging above stage 5 is word-wise XOR.
//   * Rate-0 nodes (all frozen) are skipped at every level; rate-1 subtrees take the hard-decision
//     shortcut when no LLR is exactly 0 (both are exact, see polar_common.cuh).
#include "polar_common.cuh"
#include "polar_internal.h"

namespace polar {

constexpr unsigned FULLMASK = 0xFFFFFFFFu;

// ------------------------------------------------------------------ n <= 32: thread per codeword
template <int T>
__global__ void __launch_bounds__(128) sc_small_kernel(const float *__restrict__ logit,
                                                       const uint32_t *__restrict__ fmask, int64_t B,
                                                       uint32_t *__restrict__ u_packed,
                                                       float *__restrict__ u_info,
                                                       const int32_t *__restrict__ info_pos, int k) {
  constexpr int N = 1 << T;
  const uint32_t fm = __ldg(fmask);
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float x[N];
    const float *row = logit + b * N;
    if (N >= 4) {
#pragma unroll
      for (int q = 0; q < N / 4; ++q) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(row) + q);
        x[4 * q + 0] = -v.x; x[4 * q + 1] = -v.y; x[4 * q + 2] = -v.z; x[4 * q + 3] = -v.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) x[j] = -__ldg(row + j);
    }
    uint32_t u;
    SubTree<T>::run(x, fm, u);
    if (u_packed) u_packed[b] = u;
    if (u_info)
      for (int t = 0; t < k; ++t) u_info[b * k + t] = (float)((u >> __ldg(info_pos + t)) & 1u);
  }
}

// ------------------------------------------------------------------ n >= 64: warp per CW codewords
struct ScLayout {
  int n, m, nw, stride;          // stride: floats per codeword in smem (n - 28, multiple of 4, /4 odd)
  size_t mask_bytes, per_warp_bytes;
};
__host__ __device__ inline ScLayout sc_layout(int n, int cw) {
  ScLayout l;
  l.n = n; l.m = ilog2(n); l.nw = n >> 5; l.stride = n - 28;
  l.mask_bytes = (size_t)((l.nw * 4 + 15) / 16) * 16;
  l.per_warp_bytes = (size_t)cw * ((size_t)l.stride + 2 * (size_t)(l.nw + 1)) * 4;
  l.per_warp_bytes = (l.per_warp_bytes + 15) / 16 * 16;
  return l;
}

template <int CW>
__global__ void __launch_bounds__(256) sc_tree_kernel(const float *__restrict__ logit,
                                                      const uint32_t *__restrict__ fmask_g, int n,
                                                      int64_t B, int64_t nbatches,
                                                      uint32_t *__restrict__ u_packed,
                                                      float *__restrict__ u_info,
                                                      const int32_t *__restrict__ info_pos, int k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const ScLayout lay = sc_layout(n, CW);
  const int m = lay.m, nw = lay.nw, stride = lay.stride;
  const int nws = nw + 1;   // padded row stride (words) of the beta / decision arrays
  constexpr int LOGCW = (CW == 1) ? 0 : (CW == 2) ? 1 : (CW == 4) ? 2 : (CW == 8) ? 3 : (CW == 16) ? 4 : 5;

  uint32_t *fmask = reinterpret_cast<uint32_t *>(smem_raw);
  float *L = reinterpret_cast<float *>(smem_raw + lay.mask_bytes + (size_t)warp * lay.per_warp_bytes);
  uint32_t *beta = reinterpret_cast<uint32_t *>(L + (size_t)CW * stride);
  uint32_t *uo = beta + CW * nws;

  for (int i = threadIdx.x; i < nw; i += blockDim.x) fmask[i] = __ldg(fmask_g + i);
  __syncthreads();

  const int nblk = nw;  // 32-leaf blocks per codeword
  for (int64_t batch = (int64_t)blockIdx.x * nwarps + warp; batch < nbatches; batch += (int64_t)gridDim.x * nwarps) {
    const int64_t cw0 = batch * CW;
    const int nvalid = (int)((B - cw0) < (int64_t)CW ? (B - cw0) : (int64_t)CW);

    int i = 0;
    while (i < nblk) {
      // node entered at block i: the root (i == 0) or the right child whose left sibling just finished
      const int S = (i == 0) ? m : 5 + (__ffs(i) - 1);
      int s = S;
      bool zeroed = false;
      // ---- rate-0 test for the entered node (saves the g step as well)
      {
        const int nwd = 1 << (s - 5);
        bool ok = true;
        for (int w = lane; w < nwd; w += 32) ok &= (fmask[i + w] == FULLMASK);
        zeroed = __all_sync(FULLMASK, ok);
      }
      if (!zeroed && S < m) {
        // ---- g step: stage S+1 -> S.  a = L[S+1][j], b = L[S+1][j+h], u = beta of the left sibling
        const int h = 1 << S, hq = h >> 2, lq = S - 2;     // hq quads per codeword
        const int left_blk = i - (h >> 5);
        float *dst = L + (h - 32);
        const int Q = CW * hq;
        if (S + 1 == m) {
          for (int q = lane; q < Q; q += 32) {
            const int c = q >> lq, j = (q & (hq - 1)) << 2;
            const int cl = c < nvalid ? c : nvalid - 1;
            const float4 *row = reinterpret_cast<const float4 *>(logit + (cw0 + cl) * (int64_t)n);
            float4 a = __ldg(row + (j >> 2)), b = __ldg(row + ((j + h) >> 2));
            const uint32_t bits = beta[c * nws + left_blk + (j >> 5)] >> (j & 31);
            float4 o;   // LLR = -logit (polar_sc.py:122)
            o.x = g_minsum(-a.x, -b.x, (bits << 31) & 0x80000000u);
            o.y = g_minsum(-a.y, -b.y, (bits << 30) & 0x80000000u);
            o.z = g_minsum(-a.z, -b.z, (bits << 29) & 0x80000000u);
            o.w = g_minsum(-a.w, -b.w, (bits << 28) & 0x80000000u);
            *reinterpret_cast<float4 *>(dst + c * stride + j) = o;
          }
        } else {
          const float *src = L + (2 * h - 32);
          for (int q = lane; q < Q; q += 32) {
            const int c = q >> lq, j = (q & (hq - 1)) << 2;
            const float4 a = *reinterpret_cast<const float4 *>(src + c * stride + j);
            const float4 b = *reinterpret_cast<const float4 *>(src + c * stride + j + h);
            const uint32_t bits = beta[c * nws + left_blk + (j >> 5)] >> (j & 31);
            float4 o;
            o.x = g_minsum(a.x, b.x, (bits << 31) & 0x80000000u);
            o.y = g_minsum(a.y, b.y, (bits << 30) & 0x80000000u);
            o.z = g_minsum(a.z, b.z, (bits << 29) & 0x80000000u);
            o.w = g_minsum(a.w, b.w, (bits << 28) & 0x80000000u);
            *reinterpret_cast<float4 *>(dst + c * stride + j) = o;
          }
        }
        __syncwarp();
      }
      // ---- descend along left children with f steps until a prunable node or a 32-leaf block
      while (!zeroed && s > 5) {
        {  // rate-0 test for the left child (s-1, i)
          const int nwd = 1 << (s - 6);
          bool ok = true;
          for (int w = lane; w < nwd; w += 32) ok &= (fmask[i + w] == FULLMASK);
          if (__all_sync(FULLMASK, ok)) { zeroed = true; --s; break; }
        }
        const int h = 1 << (s - 1), hq = h >> 2, lq = s - 3;
        float *dst = L + (h - 32);
        const int Q = CW * hq;
        if (s == m) {
          for (int q = lane; q < Q; q += 32) {
            const int c = q >> lq, j = (q & (hq - 1)) << 2;
            const int cl = c < nvalid ? c : nvalid - 1;
            const float4 *row = reinterpret_cast<const float4 *>(logit + (cw0 + cl) * (int64_t)n);
            float4 a = __ldg(row + (j >> 2)), b = __ldg(row + ((j + h) >> 2));
            float4 o;   // f(-a,-b) == f(a,b): the negation cancels in sign.sign and |.|
            o.x = f_minsum_neg(a.x, b.x); o.y = f_minsum_neg(a.y, b.y);
            o.z = f_minsum_neg(a.z, b.z); o.w = f_minsum_neg(a.w, b.w);
            *reinterpret_cast<float4 *>(dst + c * stride + j) = o;
          }
        } else {
          const float *src = L + (2 * h - 32);
          for (int q = lane; q < Q; q += 32) {
            const int c = q >> lq, j = (q & (hq - 1)) << 2;
            const float4 a = *reinterpret_cast<const float4 *>(src + c * stride + j);
            const float4 b = *reinterpret_cast<const float4 *>(src + c * stride + j + h);
            float4 o;
            o.x = f_minsum(a.x, b.x); o.y = f_minsum(a.y, b.y);
            o.z = f_minsum(a.z, b.z); o.w = f_minsum(a.w, b.w);
            *reinterpret_cast<float4 *>(dst + c * stride + j) = o;
          }
        }
        __syncwarp();
        --s;
      }
      // ---- node (s, i) is finished here: either zeroed (rate-0) or a 32-leaf block to decode
      const int lv0 = s - 5;   // the finished node covers 2^lv0 blocks starting at i
      if (zeroed) {
        const int nwd = 1 << lv0;
        for (int q = lane; q < (CW << lv0); q += 32) {
          const int c = q >> lv0, w = q & (nwd - 1);
          beta[c * nws + i + w] = 0u; uo[c * nws + i + w] = 0u;
        }
      } else {
        if (lane < CW) {
          float x[32];
          const float4 *src = reinterpret_cast<const float4 *>(L + lane * stride);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 v = src[q];
            x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
          }
          uint32_t u;
          const uint32_t b = SubTree<5>::run(x, fmask[i], u);
          beta[lane * nws + i] = b; uo[lane * nws + i] = u;
        }
      }
      __syncwarp();
      // ---- merge partial sums upward while the finished node is a right child
      {
        int lv = lv0, a = i;
        while (lv < m - 5 && ((a >> lv) & 1)) {
          const int nwd = 1 << lv, left = a - nwd;
          for (int q = lane; q < (CW << lv); q += 32) {
            const int c = q >> lv, w = q & (nwd - 1);
            beta[c * nws + left + w] ^= beta[c * nws + a + w];   // [bl ^ br, br]  (polar_sc.py:83-89)
          }
          __syncwarp();
          a = left; ++lv;
        }
      }
      i += 1 << lv0;
    }
    // ---- outputs
    if (u_packed) {
      for (int q = lane; q < CW * nw; q += 32) {
        const int c = q / nw, w = q - c * nw;
        if (c < nvalid) u_packed[(cw0 + c) * nw + w] = uo[c * nws + w];
      }
    }
    if (u_info) {
      for (int c = 0; c < nvalid; ++c) {
        float *row = u_info + (cw0 + c) * (int64_t)k;
        for (int t = lane; t < k; t += 32) {
          const int p = __ldg(info_pos + t);
          row[t] = (float)((uo[c * nws + (p >> 5)] >> (p & 31)) & 1u);
        }
      }
    }
    __syncwarp();
  }
  (void)LOGCW;
}


// ------------------------------------------------------------------ n >= 64: CTA per `cw` codewords
// Same algorithm as sc_tree_kernel, different mapping: ALL warps of the CTA cooperate on the wide
// (stage >= 5) f/g/merge steps of `cw` <= 32 codewords, and ONE warp then decodes the 32-leaf subtrees
// with one lane per codeword.  The bottom phase issues the same ~800 instructions whether 8 or 32 lanes
// are active, so putting up to 32 codewords behind one bottom warp cuts the issue slots per codeword
// ~3x versus the warp-per-8-codewords mapping, while shared memory (4.2 KB per codeword) still lets
// two CTAs share an SM and overlap each other's phases.
struct ScCtaLayout {
  int nw, nws, stride;
  size_t mask_bytes, node_bytes, llr_off, beta_off, uo_off, total;
};
__host__ __device__ inline ScCtaLayout sc_cta_layout(int n, int cw) {
  ScCtaLayout l;
  l.nw = n >> 5; l.nws = l.nw + 1; l.stride = n - 28;
  l.mask_bytes = (size_t)((l.nw * 4 + 15) / 16) * 16;
  l.node_bytes = (size_t)((2 * l.nw + 15) / 16) * 16;
  l.llr_off = l.mask_bytes + l.node_bytes;
  l.beta_off = l.llr_off + (size_t)cw * l.stride * 4;
  l.uo_off = l.beta_off + (size_t)cw * l.nws * 4;
  l.total = ((l.uo_off + (size_t)cw * l.nws * 4 + 15) / 16) * 16;
  return l;
}

template <bool IS_G, bool FROM_GLOBAL>
__device__ __forceinline__ void sc_top_step(const float *__restrict__ logit, int64_t cw0, int nvalid, int n, float *L,
                                            const uint32_t *beta, int nws, int stride, int cw, int h, int lgh,
                                            int left_blk, int tid, int nthr) {
  // stage 2h -> h : out[j] = f(a[j], a[j+h]) or g(a[j], a[j+h], beta_left[j]), 4 elements per thread
  const int hq = h >> 2, lq = lgh - 2;
  float *dst = L + (h - 32);
  const float *src = L + (2 * h - 32);
  const int Q = cw * hq;
  for (int q = tid; q < Q; q += nthr) {
    const int c = q >> lq, j = (q & (hq - 1)) << 2;
    float4 a, b;
    if (FROM_GLOBAL) {
      const int cl = c < nvalid ? c : nvalid - 1;
      const float4 *row = reinterpret_cast<const float4 *>(logit + (cw0 + cl) * (int64_t)n);
      a = __ldg(row + (j >> 2)); b = __ldg(row + ((j + h) >> 2));
      if (IS_G) {   // LLR = -logit (polar_sc.py:122); for f the negation cancels
        a.x = -a.x; a.y = -a.y; a.z = -a.z; a.w = -a.w; b.x = -b.x; b.y = -b.y; b.z = -b.z; b.w = -b.w;
      }
    } else {
      a = *reinterpret_cast<const float4 *>(src + c * stride + j);
      b = *reinterpret_cast<const float4 *>(src + c * stride + j + h);
    }
    float4 o;
    if (IS_G) {
      const uint32_t bits = beta[c * nws + left_blk + (j >> 5)] >> (j & 31);
      o.x = g_minsum(a.x, b.x, (bits << 31) & 0x80000000u);
      o.y = g_minsum(a.y, b.y, (bits << 30) & 0x80000000u);
      o.z = g_minsum(a.z, b.z, (bits << 29) & 0x80000000u);
      o.w = g_minsum(a.w, b.w, (bits << 28) & 0x80000000u);
    } else if (FROM_GLOBAL) {
      o.x = f_minsum_neg(a.x, b.x); o.y = f_minsum_neg(a.y, b.y); o.z = f_minsum_neg(a.z, b.z); o.w = f_minsum_neg(a.w, b.w);
    } else {
      o.x = f_minsum(a.x, b.x); o.y = f_minsum(a.y, b.y); o.z = f_minsum(a.z, b.z); o.w = f_minsum(a.w, b.w);
    }
    *reinterpret_cast<float4 *>(dst + c * stride + j) = o;
  }
}

__global__ void __launch_bounds__(256) sc_cta_kernel(const float *__restrict__ logit,
                                                     const uint32_t *__restrict__ fmask_g, int n, int cw,
                                                     int64_t B, int64_t nbatches,
                                                     uint32_t *__restrict__ u_packed, float *__restrict__ u_info,
                                                     const int32_t *__restrict__ info_pos, int k) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const ScCtaLayout lay = sc_cta_layout(n, cw);
  const int m = ilog2(n), nw = lay.nw, nws = lay.nws, stride = lay.stride;
  uint32_t *fmask = reinterpret_cast<uint32_t *>(smem_raw);
  unsigned char *nz = smem_raw + lay.mask_bytes;            // nz[(nw >> lv) + (i >> lv)] = node of 2^lv blocks at block i is rate-0
  float *L = reinterpret_cast<float *>(smem_raw + lay.llr_off);
  uint32_t *beta = reinterpret_cast<uint32_t *>(smem_raw + lay.beta_off);
  uint32_t *uo = reinterpret_cast<uint32_t *>(smem_raw + lay.uo_off);

  for (int i = tid; i < nw; i += nthr) {
    const uint32_t w = __ldg(fmask_g + i);
    fmask[i] = w; nz[nw + i] = (w == FULLMASK);
  }
  __syncthreads();
  if (tid == 0)
    for (int idx = nw - 1; idx >= 1; --idx) nz[idx] = nz[2 * idx] & nz[2 * idx + 1];
  __syncthreads();

  for (int64_t batch = blockIdx.x; batch < nbatches; batch += gridDim.x) {
    const int64_t cw0 = batch * cw;
    const int nvalid = (int)((B - cw0) < (int64_t)cw ? (B - cw0) : (int64_t)cw);
    int i = 0;
    while (i < nw) {
      const int S = (i == 0) ? m : 5 + (__ffs(i) - 1);
      int s = S;
      bool zeroed = nz[(nw >> (S - 5)) + (i >> (S - 5))] != 0;
      if (!zeroed && S < m) {
        const int h = 1 << S;
        if (S + 1 == m) sc_top_step<true, true>(logit, cw0, nvalid, n, L, beta, nws, stride, cw, h, S, i - (h >> 5), tid, nthr);
        else sc_top_step<true, false>(logit, cw0, nvalid, n, L, beta, nws, stride, cw, h, S, i - (h >> 5), tid, nthr);
        __syncthreads();
      }
      while (!zeroed && s > 5) {
        if (nz[(nw >> (s - 6)) + (i >> (s - 6))]) { zeroed = true; --s; break; }
        const int h = 1 << (s - 1);
        if (s == m) sc_top_step<false, true>(logit, cw0, nvalid, n, L, beta, nws, stride, cw, h, s - 1, 0, tid, nthr);
        else sc_top_step<false, false>(logit, cw0, nvalid, n, L, beta, nws, stride, cw, h, s - 1, 0, tid, nthr);
        __syncthreads();
        --s;
      }
      const int lv0 = s - 5;
      if (zeroed) {
        const int nwd = 1 << lv0;
        for (int q = tid; q < (cw << lv0); q += nthr) {
          const int c = q >> lv0, w = q & (nwd - 1);
          beta[c * nws + i + w] = 0u; uo[c * nws + i + w] = 0u;
        }
      } else if (tid < cw) {
        float x[32];
        const float4 *src = reinterpret_cast<const float4 *>(L + tid * stride);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = src[q];
          x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
        }
        uint32_t u;
        const uint32_t bb = SubTree<5>::run(x, fmask[i], u);
        beta[tid * nws + i] = bb; uo[tid * nws + i] = u;
      }
      __syncthreads();
      {
        int lv = lv0, a = i;
        while (lv < m - 5 && ((a >> lv) & 1)) {
          const int nwd = 1 << lv, left = a - nwd;
          for (int q = tid; q < (cw << lv); q += nthr) {
            const int c = q >> lv, w = q & (nwd - 1);
            beta[c * nws + left + w] ^= beta[c * nws + a + w];
          }
          __syncthreads();
          a = left; ++lv;
        }
      }
      i += 1 << lv0;
    }
    if (u_packed) {
      for (int q = tid; q < cw * nw; q += nthr) {
        const int c = q / nw, w = q - c * nw;
        if (c < nvalid) u_packed[(cw0 + c) * nw + w] = uo[c * nws + w];
      }
    }
    if (u_info) {
      for (int q = tid; q < nvalid * k; q += nthr) {
        const int c = q / k, t = q - c * k;
        const int p = __ldg(info_pos + t);
        u_info[(cw0 + c) * (int64_t)k + t] = (float)((uo[c * nws + (p >> 5)] >> (p & 31)) & 1u);
      }
    }
    __syncthreads();
  }
}

static int launch_cta(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed, float *u_info,
                      const int32_t *info_pos, int k, int cw, int threads, int ctas_per_sm, cudaStream_t st) {
  const int max_smem = device_max_smem_optin();
  if (cw > 32) cw = 32;
  if (cw < 1) cw = 1;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  // fit `ctas_per_sm` CTAs in the 228 KB of an SM (1 KB reserved per CTA)
  while (cw > 1 && sc_cta_layout(n, cw).total + 1024 > (size_t)(228 * 1024) / ctas_per_sm) --cw;
  ScCtaLayout lay = sc_cta_layout(n, cw);
  if (lay.total > (size_t)max_smem) return set_error(POLAR_ENOMEM, "sc: n=%d needs %zu B shared memory per CTA", n, lay.total);
  POLAR_CUDA(cudaFuncSetAttribute(sc_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
  const int64_t nbatches = (B + cw - 1) / cw;
  int64_t grid = nbatches;
  const int64_t cap = (int64_t)device_sm_count() * ctas_per_sm;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  sc_cta_kernel<<<(unsigned)grid, threads, lay.total, st>>>(logit, fmask, n, cw, B, nbatches, u_packed, u_info, info_pos, k);
  count_launch();
  POLAR_CHECK_LAUNCH("sc_cta_kernel");
  return POLAR_OK;
}

template <int CW>
static int launch_tree(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed,
                       float *u_info, const int32_t *info_pos, int k, int warps, cudaStream_t st) {
  const ScLayout lay = sc_layout(n, CW);
  const int max_smem = device_max_smem_optin();
  while (warps > 1 && lay.mask_bytes + (size_t)warps * lay.per_warp_bytes > (size_t)max_smem) warps >>= 1;
  const size_t smem = lay.mask_bytes + (size_t)warps * lay.per_warp_bytes;
  if (smem > (size_t)max_smem) return set_error(POLAR_ENOMEM, "sc: n=%d CW=%d needs %zu B shared memory", n, CW, smem);
  auto kern = sc_tree_kernel<CW>;
  POLAR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nbatches = (B + CW - 1) / CW;
  int ctas_per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  if (ctas_per_sm > 16) ctas_per_sm = 16;
  int64_t grid = (nbatches + warps - 1) / warps;
  const int64_t cap = (int64_t)device_sm_count() * ctas_per_sm;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, warps * 32, smem, st>>>(logit, fmask, n, B, nbatches, u_packed, u_info, info_pos, k);
  count_launch();
  POLAR_CHECK_LAUNCH("sc_tree_kernel");
  return POLAR_OK;
}

template <int T>
static int launch_small(const float *logit, const uint32_t *fmask, int64_t B, uint32_t *u_packed, float *u_info,
                        const int32_t *info_pos, int k, cudaStream_t st) {
  int64_t grid = (B + 127) / 128;
  const int64_t cap = (int64_t)device_sm_count() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  sc_small_kernel<T><<<(unsigned)grid, 128, 0, st>>>(logit, fmask, B, u_packed, u_info, info_pos, k);
  count_launch();
  POLAR_CHECK_LAUNCH("sc_small_kernel");
  return POLAR_OK;
}

int sc_default_cw(int n) {
  // per-warp shared memory ~ CW * 4n bytes; keep it near 16-32 KB so several warps fit per SM
  int cw = (32 * 1024) / (4 * n);
  int p = 1;
  while (p * 2 <= cw && p < 32) p *= 2;
  return p;
}

}  // namespace polar

extern "C" int polar_sc_decode_f32(const float *d_logit, const uint32_t *d_frozen_mask, int n, int64_t B,
                                   uint32_t *d_u_packed, float *d_u_info_f32, const int32_t *d_info_pos,
                                   int k, void *stream) {
  using namespace polar;
  if (!is_pow2(n) || n < 2 || n > POLAR_MAX_N) return set_error(POLAR_EINVAL, "sc: n=%d must be a power of two in [2,%d]", n, POLAR_MAX_N);
  if (B < 0) return set_error(POLAR_EINVAL, "sc: B=%lld < 0", (long long)B);
  if (B == 0) return POLAR_OK;
  if (!d_logit || !d_frozen_mask) return set_error(POLAR_EINVAL, "sc: null logit / frozen_mask");
  if (!d_u_packed && !d_u_info_f32) return set_error(POLAR_EINVAL, "sc: no output buffer");
  if (d_u_info_f32 && (!d_info_pos || k < 0 || k > n)) return set_error(POLAR_EINVAL, "sc: u_info requested without valid info_pos/k");
  if (n >= 4 && ((uintptr_t)d_logit & 15)) return set_error(POLAR_EALIGN, "sc: logit must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 32) {
    switch (n) {
      case 2: return launch_small<1>(d_logit, d_frozen_mask, B, d_u_packed, d_u_info_f32, d_info_pos, k, st);
      case 4: return launch_small<2>(d_logit, d_frozen_mask, B, d_u_packed, d_u_info_f32, d_info_pos, k, st);
      case 8: return launch_small<3>(d_logit, d_frozen_mask, B, d_u_packed, d_u_info_f32, d_info_pos, k, st);
      case 16: return launch_small<4>(d_logit, d_frozen_mask, B, d_u_packed, d_u_info_f32, d_info_pos, k, st);
      default: return launch_small<5>(d_logit, d_frozen_mask, B, d_u_packed, d_u_info_f32, d_info_pos, k, st);
    }
  }
  const int mode = env_int("POLAR_SC_MODE", 3);
  if (mode == 3 && n >= 128 && n <= 2048)   // default: warp-autonomous decoder, tensor memory scratch (polar_sc4.cu)
    return polar::launch_sc4(d_logit, d_frozen_mask, n, B, d_u_packed, d_u_info_f32, d_info_pos, k, env_int("POLAR_SC_WARPS_SM", 0), st);
  if (mode >= 2 && n >= 128)   // default: compile-time tree, virtual top stage, 64-leaf register subtrees (polar_sc3.cu)
    return polar::launch_sc3(d_logit, d_frozen_mask, n, B, d_u_packed, d_u_info_f32, d_info_pos, k, env_int("POLAR_SC_CTA_CW", 32),
                      env_int("POLAR_SC_CTAS", 0), st);
  if (mode >= 1) {
    // CTA mapping (default): up to 32 codewords per CTA, `ctas` CTAs per SM
    int ctas = env_int("POLAR_SC_CTAS", 3);
    if (ctas < 1) ctas = 3;
    int threads = env_int("POLAR_SC_THREADS", 256);
    if (threads < 32) threads = 32;
    if (threads > 256) threads = 256;
    threads &= ~31;
    return launch_cta(d_logit, d_frozen_mask, n, B, d_u_packed, d_u_info_f32, d_info_pos, k,
                      env_int("POLAR_SC_CTA_CW", 32), threads, ctas, st);
  }
  int cw = env_int("POLAR_SC_CW", sc_default_cw(n));
  int warps = env_int("POLAR_SC_WARPS", 2);
  if (warps < 1) warps = 1;
  if (warps > 8) warps = 8;
#define POLAR_SC_CASE(C) \
  case C: return launch_tree<C>(d_logit, d_frozen_mask, n, B, d_u_packed, d_u_info_f32, d_info_pos, k, warps, st)
  switch (cw) {
    POLAR_SC_CASE(1); POLAR_SC_CASE(2); POLAR_SC_CASE(4); POLAR_SC_CASE(8); POLAR_SC_CASE(16); POLAR_SC_CASE(32);
    default: return set_error(POLAR_EINVAL, "sc: POLAR_SC_CW=%d must be 1,2,4,8,16 or 32", cw);
  }
#undef POLAR_SC_CASE
}
