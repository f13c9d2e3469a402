// polar_count.cu -- bit / block error counters (sm_100a).
// Replaces count_errors / count_block_errors (my_sn/sim.py:7-18): sum(b != b_hat) and
// sum(any(b != b_hat, -1)).  Popcount over packed words (or a compare over the fp32 API tensors),
// per-warp accumulation in registers, one atomicAdd pair per warp at the end.
#include "polar_internal.h"

namespace polar {

__device__ __forceinline__ unsigned warp_sum(unsigned v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  return v;
}

__global__ void __launch_bounds__(256) count_packed_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                           const uint32_t *__restrict__ mask, int nw, int64_t B,
                                                           unsigned long long *__restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long bit_acc = 0, blk_acc = 0;
  for (int64_t row = warp0; row < B; row += nwarps) {
    unsigned cnt = 0;
    for (int w = lane; w < nw; w += 32) {
      uint32_t d = __ldg(a + row * nw + w) ^ __ldg(b + row * nw + w);
      if (mask) d &= __ldg(mask + w);
      cnt += __popc(d);
    }
    cnt = warp_sum(cnt);
    bit_acc += cnt; blk_acc += (cnt != 0);
  }
  if (lane == 0 && (bit_acc | blk_acc)) {
    atomicAdd(counters, bit_acc);
    atomicAdd(counters + 1, blk_acc);
  }
}

__global__ void __launch_bounds__(256) count_f32_kernel(const float *__restrict__ a, const float *__restrict__ b, int k,
                                                        int64_t B, unsigned long long *__restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long bit_acc = 0, blk_acc = 0;
  for (int64_t row = warp0; row < B; row += nwarps) {
    unsigned cnt = 0;
    for (int t = lane; t < k; t += 32) cnt += (__ldg(a + row * k + t) != __ldg(b + row * k + t));
    cnt = warp_sum(cnt);
    bit_acc += cnt; blk_acc += (cnt != 0);
  }
  if (lane == 0 && (bit_acc | blk_acc)) {
    atomicAdd(counters, bit_acc);
    atomicAdd(counters + 1, blk_acc);
  }
}

}  // namespace polar

using namespace polar;

static unsigned cnt_grid(int64_t rows) {
  int64_t g = (rows + 7) / 8;
  const int64_t cap = (int64_t)device_sm_count() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (unsigned)g;
}

extern "C" int polar_count_errors_packed(const uint32_t *d_a, const uint32_t *d_b, const uint32_t *d_mask, int n,
                                         int64_t B, unsigned long long *d_counters, void *stream) {
  if (n < 1 || B < 0) return set_error(POLAR_EINVAL, "count: bad n/B");
  if (B == 0) return POLAR_OK;
  if (!d_a || !d_b || !d_counters) return set_error(POLAR_EINVAL, "count: null pointer");
  count_packed_kernel<<<cnt_grid(B), 256, 0, (cudaStream_t)stream>>>(d_a, d_b, d_mask, POLAR_WORDS(n), B, d_counters);
  count_launch();
  POLAR_CHECK_LAUNCH("count_packed");
  return POLAR_OK;
}

extern "C" int polar_count_errors_f32(const float *d_b, const float *d_b_hat, int k, int64_t B,
                                      unsigned long long *d_counters, void *stream) {
  if (k < 1 || B < 0) return set_error(POLAR_EINVAL, "count: bad k/B");
  if (B == 0) return POLAR_OK;
  if (!d_b || !d_b_hat || !d_counters) return set_error(POLAR_EINVAL, "count: null pointer");
  count_f32_kernel<<<cnt_grid(B), 256, 0, (cudaStream_t)stream>>>(d_b, d_b_hat, k, B, d_counters);
  count_launch();
  POLAR_CHECK_LAUNCH("count_f32");
  return POLAR_OK;
}
