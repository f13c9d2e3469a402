// polar_api.cu -- error text, launch accounting, device queries for the C ABI (include/polar_b200.h).
#include <atomic>
#include <mutex>
#include <string>
#include <string.h>
#include <vector>

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

std::atomic<unsigned> g_opt_gen{1};
static std::mutex g_opt_mu;
static std::vector<std::pair<std::string, int>> g_opt_over;

bool opt_lookup(const char *name, int *value) {
  {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    for (auto &kv : g_opt_over)
      if (kv.first == name) { *value = kv.second; return true; }
  }
  const char *v = getenv(name);
  if (!v || !*v) return false;
  *value = atoi(v);
  return true;
}

static std::atomic<int> g_sm_count[64];
static std::atomic<int> g_smem_optin[64];
static void query_device(int dev) {           // attribute queries only: no allocation, no synchronisation
  if (dev < 0 || dev >= 64 || g_sm_count[dev].load(std::memory_order_relaxed)) return;
  int v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  g_smem_optin[dev].store(v > 0 ? v : 227 * 1024, std::memory_order_relaxed);
  v = 0;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  g_sm_count[dev].store(v > 0 ? v : 148, std::memory_order_release);
}
int device_sm_count() {
  int dev = 0; cudaGetDevice(&dev); query_device(dev);
  return (dev >= 0 && dev < 64 && g_sm_count[dev]) ? g_sm_count[dev].load() : 148;
}
int device_max_smem_optin() {
  int dev = 0; cudaGetDevice(&dev); query_device(dev);
  return (dev >= 0 && dev < 64 && g_smem_optin[dev]) ? g_smem_optin[dev].load() : 227 * 1024;
}

}  // namespace polar

extern "C" {
const char *polar_last_error(void) { return polar::g_err; }
const char *polar_version(void) { return "polar_b200 0.2 (sm_100a; SC/SCL/encoder/AWGN front end/error counters)"; }

int polar_init(int device) {
  using namespace polar;
  if (device < 0 || device >= 64) return set_error(POLAR_EINVAL, "init: bad device %d", device);
  int prev = -1;
  POLAR_CUDA(cudaGetDevice(&prev));
  POLAR_CUDA(cudaSetDevice(device));
  query_device(device);
  const int rc = sc_scratch_init(device);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  return rc;
}

int polar_set_option(const char *name, int value) {
  using namespace polar;
  if (!name || strncmp(name, "POLAR_", 6) != 0) return set_error(POLAR_EINVAL, "set_option: names start with POLAR_");
  {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    bool found = false;
    for (auto &kv : g_opt_over)
      if (kv.first == name) { kv.second = value; found = true; }
    if (!found) g_opt_over.emplace_back(name, value);
  }
  g_opt_gen.fetch_add(1, std::memory_order_acq_rel);
  return POLAR_OK;
}

void polar_clear_options(void) {
  using namespace polar;
  {
    std::lock_guard<std::mutex> lk(g_opt_mu);
    g_opt_over.clear();
  }
  g_opt_gen.fetch_add(1, std::memory_order_acq_rel);
}
unsigned long long polar_launch_count(void) { return polar::g_launches.load(); }
}
