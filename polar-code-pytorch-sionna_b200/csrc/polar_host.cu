// polar_host.cu -- host-buffer entry points: the end-to-end path a caller with HOST tensors uses.
// The batch is cut into chunks; chunk c is copied H2D, decoded and copied back D2H on stream c%2,
// so PCIe transfers of one chunk overlap the kernel of the other.  Device staging buffers are cached
// per device (grow-only) so repeated calls do not pay cudaMalloc.
#include <mutex>

#include "polar_internal.h"

namespace polar {

struct HostCtx {
  cudaStream_t st[2] = {nullptr, nullptr};
  void *logit[2] = {nullptr, nullptr}; size_t logit_bytes = 0;
  void *out[2] = {nullptr, nullptr}; size_t out_bytes = 0;
  void *pm[2] = {nullptr, nullptr}; size_t pm_bytes = 0;
  void *ws[2] = {nullptr, nullptr}; size_t ws_bytes = 0;
  void *mask = nullptr; size_t mask_bytes = 0;
  void *crc = nullptr; size_t crc_bytes = 0;
};
static HostCtx g_ctx[64];
static std::mutex g_mu;

static int grow(void **p, size_t *have, size_t need) {
  if (*have >= need) return POLAR_OK;
  if (*p) cudaFree(*p);
  *p = nullptr; *have = 0;
  if (cudaMalloc(p, need) != cudaSuccess) return set_error(POLAR_ENOMEM, "host path: cudaMalloc(%zu) failed", need);
  *have = need;
  return POLAR_OK;
}
static int grow2(void *p[2], size_t *have, size_t need) {
  if (*have >= need) return POLAR_OK;
  for (int i = 0; i < 2; ++i) {
    if (p[i]) cudaFree(p[i]);
    p[i] = nullptr;
    if (cudaMalloc(&p[i], need) != cudaSuccess) { *have = 0; return set_error(POLAR_ENOMEM, "host path: cudaMalloc(%zu) failed", need); }
  }
  *have = need;
  return POLAR_OK;
}

static int64_t chunk_codewords(int n, int64_t B) {
  int64_t c = env_int("POLAR_HOST_CHUNK_MB", 128) * (int64_t)(1 << 20) / ((int64_t)n * 4);
  if (c < 1) c = 1;
  if (c > B) c = B;
  return c;
}

}  // namespace polar

using namespace polar;

extern "C" int polar_sc_decode_host(const float *h_logit, const uint32_t *h_frozen_mask, int n, int64_t B,
                                    uint32_t *h_u_packed, int device) {
  if (!h_logit || !h_frozen_mask || !h_u_packed) return set_error(POLAR_EINVAL, "sc host: null pointer");
  if (!is_pow2(n) || n < 2 || n > POLAR_MAX_N || B < 0) return set_error(POLAR_EINVAL, "sc host: bad n/B");
  if (device < 0 || device >= 64) return set_error(POLAR_EINVAL, "sc host: bad device");
  if (B == 0) return POLAR_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  POLAR_CUDA(cudaSetDevice(device));
  HostCtx &C = g_ctx[device];
  for (int i = 0; i < 2; ++i)
    if (!C.st[i]) POLAR_CUDA(cudaStreamCreateWithFlags(&C.st[i], cudaStreamNonBlocking));
  const int nw = POLAR_WORDS(n);
  const int64_t chunk = chunk_codewords(n, B);
  int rc;
  if ((rc = grow2(C.logit, &C.logit_bytes, (size_t)chunk * n * 4))) return rc;
  if ((rc = grow2(C.out, &C.out_bytes, (size_t)chunk * nw * 4))) return rc;
  if ((rc = grow(&C.mask, &C.mask_bytes, (size_t)nw * 4))) return rc;
  POLAR_CUDA(cudaMemcpyAsync(C.mask, h_frozen_mask, (size_t)nw * 4, cudaMemcpyHostToDevice, C.st[0]));
  POLAR_CUDA(cudaStreamSynchronize(C.st[0]));
  int idx = 0;
  for (int64_t b0 = 0; b0 < B; b0 += chunk, idx ^= 1) {
    const int64_t nb = (B - b0) < chunk ? (B - b0) : chunk;
    cudaStream_t st = C.st[idx];
    POLAR_CUDA(cudaMemcpyAsync(C.logit[idx], h_logit + b0 * n, (size_t)nb * n * 4, cudaMemcpyHostToDevice, st));
    rc = polar_sc_decode_f32((const float *)C.logit[idx], (const uint32_t *)C.mask, n, nb, (uint32_t *)C.out[idx], nullptr, nullptr, 0, st);
    if (rc) return rc;
    POLAR_CUDA(cudaMemcpyAsync(h_u_packed + b0 * nw, C.out[idx], (size_t)nb * nw * 4, cudaMemcpyDeviceToHost, st));
  }
  POLAR_CUDA(cudaStreamSynchronize(C.st[0]));
  POLAR_CUDA(cudaStreamSynchronize(C.st[1]));
  return POLAR_OK;
}

extern "C" int polar_scl_decode_host(const float *h_logit, const uint32_t *h_frozen_mask, int n, int L, int64_t B,
                                     uint32_t *h_best_packed, double *h_pm_sorted, const uint32_t *h_crc_rows,
                                     int crc_len, int device) {
  if (!h_logit || !h_frozen_mask || !h_best_packed) return set_error(POLAR_EINVAL, "scl host: null pointer");
  if (!is_pow2(n) || n < 2 || n > POLAR_SCL_MAX_N || !is_pow2(L) || L > POLAR_SCL_MAX_L || B < 0) return set_error(POLAR_EINVAL, "scl host: bad n/L/B");
  if (device < 0 || device >= 64) return set_error(POLAR_EINVAL, "scl host: bad device");
  if (B == 0) return POLAR_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  POLAR_CUDA(cudaSetDevice(device));
  HostCtx &C = g_ctx[device];
  for (int i = 0; i < 2; ++i)
    if (!C.st[i]) POLAR_CUDA(cudaStreamCreateWithFlags(&C.st[i], cudaStreamNonBlocking));
  const int nw = POLAR_WORDS(n);
  int64_t chunk = chunk_codewords(n, B);
  if (scl3_supported(n, L) && chunk < B) {
    // the list kernel is persistent: every resident warp decodes the same number of 32/L-codeword groups, so a chunk
    // should be a whole number of such rounds (a 128 MB chunk of n = 1024 is 3.46 rounds: 13 % of the last one idle)
    Scl3Plan p3;
    if (launch_scl3(nullptr, nullptr, n, L, B, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, &p3) == POLAR_OK) {
      const int64_t round = p3.grid * (32 / L);
      if (round > 0 && chunk > round) chunk = chunk / round * round;
    }
  }
  const size_t ws_need = polar_scl_workspace_bytes(n, L, chunk);
  int rc;
  if ((rc = grow2(C.logit, &C.logit_bytes, (size_t)chunk * n * 4))) return rc;
  if ((rc = grow2(C.out, &C.out_bytes, (size_t)chunk * nw * 4))) return rc;
  if (h_pm_sorted && (rc = grow2(C.pm, &C.pm_bytes, (size_t)chunk * L * 8))) return rc;
  if (ws_need && (rc = grow2(C.ws, &C.ws_bytes, ws_need))) return rc;
  if ((rc = grow(&C.mask, &C.mask_bytes, (size_t)nw * 4))) return rc;
  POLAR_CUDA(cudaMemcpyAsync(C.mask, h_frozen_mask, (size_t)nw * 4, cudaMemcpyHostToDevice, C.st[0]));
  if (crc_len > 0 && h_crc_rows) {
    if ((rc = grow(&C.crc, &C.crc_bytes, (size_t)n * 4))) return rc;
    POLAR_CUDA(cudaMemcpyAsync(C.crc, h_crc_rows, (size_t)n * 4, cudaMemcpyHostToDevice, C.st[0]));
  }
  POLAR_CUDA(cudaStreamSynchronize(C.st[0]));
  int idx = 0;
  for (int64_t b0 = 0; b0 < B; b0 += chunk, idx ^= 1) {
    const int64_t nb = (B - b0) < chunk ? (B - b0) : chunk;
    cudaStream_t st = C.st[idx];
    POLAR_CUDA(cudaMemcpyAsync(C.logit[idx], h_logit + b0 * n, (size_t)nb * n * 4, cudaMemcpyHostToDevice, st));
    rc = polar_scl_decode((const float *)C.logit[idx], (const uint32_t *)C.mask, n, L, nb, (uint32_t *)C.out[idx], nullptr,
                          nullptr, 0, h_pm_sorted ? (double *)C.pm[idx] : nullptr, nullptr,
                          (crc_len > 0 && h_crc_rows) ? (const uint32_t *)C.crc : nullptr, (crc_len > 0 && h_crc_rows) ? crc_len : 0,
                          ws_need ? C.ws[idx] : nullptr, ws_need ? C.ws_bytes : 0, st);
    if (rc) return rc;
    POLAR_CUDA(cudaMemcpyAsync(h_best_packed + b0 * nw, C.out[idx], (size_t)nb * nw * 4, cudaMemcpyDeviceToHost, st));
    if (h_pm_sorted) POLAR_CUDA(cudaMemcpyAsync(h_pm_sorted + b0 * L, C.pm[idx], (size_t)nb * L * 8, cudaMemcpyDeviceToHost, st));
  }
  POLAR_CUDA(cudaStreamSynchronize(C.st[0]));
  POLAR_CUDA(cudaStreamSynchronize(C.st[1]));
  return POLAR_OK;
}
