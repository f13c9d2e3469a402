// polar_internal.h -- host-side plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/polar_b200.h"

namespace polar {

int set_error(int code, const char *fmt, ...);      // stores thread-local text, returns code
void count_launch(int n = 1);                        // gpu_launches accounting
int device_sm_count();
int device_max_smem_optin();

// Tuning / test options (POLAR_* names).  An option is an explicit override set through polar_set_option(), else the
// environment variable of the same name, else the call site's default.  The environment is NOT read on the launch path:
// every call site caches its lookup and repeats it only after polar_set_option() / polar_clear_options() bumped the
// generation counter, so a decode costs one relaxed atomic load per option.
struct OptSlot { std::atomic<unsigned> gen{0}; std::atomic<int> has{0}; std::atomic<int> val{0}; };
extern std::atomic<unsigned> g_opt_gen;
bool opt_lookup(const char *name, int *value);
inline int opt_get(OptSlot &s, const char *name, int dflt) {
  const unsigned g = g_opt_gen.load(std::memory_order_acquire);
  if (s.gen.load(std::memory_order_relaxed) != g) {
    int v = 0;
    const bool has = opt_lookup(name, &v);
    s.val.store(v, std::memory_order_relaxed); s.has.store(has ? 1 : 0, std::memory_order_relaxed);
    s.gen.store(g, std::memory_order_release);
  }
  return s.has.load(std::memory_order_relaxed) ? s.val.load(std::memory_order_relaxed) : dflt;
}
#define env_int(NAME, DFLT) ([&]() -> int { static ::polar::OptSlot s__; return ::polar::opt_get(s__, NAME, (DFLT)); }())

// polar_sc5.cu: SC decoder with TMA-staged top stages and a scratch hierarchy (n in [1024, 8192])
int launch_sc5(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed, float *u_info,
               const int32_t *info_pos, int k, int warps, cudaStream_t st);

// per-device stage scratch of the SC kernel for n >= 1024 (polar_sc5.cu), allocated by polar_init(): one slot per
// physical SM; 4 MB hold the live nodes of stages 9 .. m-1 of every codeword in flight on the SM up to n = 8192
constexpr size_t kScScratchPerSm = (size_t)4 << 20;
int sc_scratch_init(int device);
float *sc_scratch();                                 // nullptr until polar_init(device) ran

// polar_sc4.cu: warp-autonomous SC decoder, everything on chip (n in [128, 512])
int launch_sc4(const float *logit, const uint32_t *fmask, int n, int64_t B, uint32_t *u_packed, float *u_info,
               const int32_t *info_pos, int k, int warps, cudaStream_t st);

// polar_scl3.cu: SCL decoder with compile-time tree, virtual top stages and on-chip LLR tree (n in [64, 4096], L in [2, 32])
struct Scl3Plan { int64_t grid; size_t ws_bytes_per_warp; };
bool scl3_supported(int n, int L);
int launch_scl3(const float *logit, const uint32_t *fmask, int n, int L, int64_t B, uint32_t *best, float *u_info,
                const int32_t *info_pos, int k, double *pm_out, uint32_t *list, const uint32_t *crc_rows, int crc_len,
                void *ws, cudaStream_t st, Scl3Plan *plan_only);

inline bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

#define POLAR_CHECK_LAUNCH(what)                                                           \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) return polar::set_error(POLAR_ECUDA, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

#define POLAR_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) return polar::set_error(POLAR_ECUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace polar
