// polar_api.cu -- error text, launch accounting, device queries for the C ABI (include/polar_b200.h).
#include <atomic>
#include <string.h>

#include "polar_internal.h"

namespace polar {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  return atoi(v);
}

static int g_sm_count[64];
static int g_smem_optin[64];
static void query_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  if (g_sm_count[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count[dev] = v > 0 ? v : 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    g_smem_optin[dev] = v > 0 ? v : 227 * 1024;
  }
}
int device_sm_count() {
  int dev = 0; cudaGetDevice(&dev); query_device();
  return (dev >= 0 && dev < 64 && g_sm_count[dev]) ? g_sm_count[dev] : 148;
}
int device_max_smem_optin() {
  int dev = 0; cudaGetDevice(&dev); query_device();
  return (dev >= 0 && dev < 64 && g_smem_optin[dev]) ? g_smem_optin[dev] : 227 * 1024;
}

}  // namespace polar

extern "C" {
const char *polar_last_error(void) { return polar::g_err; }
const char *polar_version(void) { return "polar_b200 0.1 (sm_100a; SC/SCL/encoder/AWGN front end/error counters)"; }
unsigned long long polar_launch_count(void) { return polar::g_launches.load(); }
}
