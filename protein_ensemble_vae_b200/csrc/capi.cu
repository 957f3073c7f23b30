// C-ABI runtime: error string, launch counter, device query.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"

namespace pev {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int after_launch(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(2, "%s: %s", kernel_name, cudaGetErrorString(e));
  return 0;
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

}  // namespace pev

extern "C" {
int pev_abi_version(void) { return PEV_ABI_VERSION; }
const char* pev_last_error(void) { return pev::g_err; }
int64_t pev_launch_count(void) { return pev::g_launches.load(); }
}
