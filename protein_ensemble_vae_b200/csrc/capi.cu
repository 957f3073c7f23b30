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

// block = 32 columns x 32 row lanes: row lane r sums partials r, r + 32, ... (independent loads, unrolled), then the 32
// row lanes of a column are combined by a fixed shared-memory tree
__global__ void __launch_bounds__(1024)
partial_reduce2_kernel(const float* __restrict__ partial, int G, int64_t stride, int n, float scale,
                       float* __restrict__ out) {
  __shared__ float sm[32][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    int g = r;
    for (; g + 96 < G; g += 128) {
      s0 += partial[(int64_t)g * stride + i];
      s1 += partial[(int64_t)(g + 32) * stride + i];
      s2 += partial[(int64_t)(g + 64) * stride + i];
      s3 += partial[(int64_t)(g + 96) * stride + i];
    }
    for (; g < G; g += 32) s0 += partial[(int64_t)g * stride + i];
  }
  sm[r][c] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int half = 16; half >= 1; half >>= 1) {
    if (r < half) sm[r][c] += sm[r + half][c];
    __syncthreads();
  }
  if (r == 0 && i < n) out[i] = scale * sm[0][c];
}

// the same reduction for a [rows, cols] block written with leading dimension ldc (a column block of a wider matrix)
__global__ void __launch_bounds__(1024)
partial_reduce2d_kernel(const float* __restrict__ partial, int G, int64_t stride, int rows, int cols, float scale,
                        float* __restrict__ out, int ldc) {
  __shared__ float sm[32][33];
  const int c = threadIdx.x & 31, r = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + c;                 // flat index into the [rows, cols] block
  const int n = rows * cols;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    int g = r;
    for (; g + 96 < G; g += 128) {
      s0 += partial[(int64_t)g * stride + i];
      s1 += partial[(int64_t)(g + 32) * stride + i];
      s2 += partial[(int64_t)(g + 64) * stride + i];
      s3 += partial[(int64_t)(g + 96) * stride + i];
    }
    for (; g < G; g += 32) s0 += partial[(int64_t)g * stride + i];
  }
  sm[r][c] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  for (int half = 16; half >= 1; half >>= 1) {
    if (r < half) sm[r][c] += sm[r + half][c];
    __syncthreads();
  }
  if (r == 0 && i < n) out[(int64_t)(i / cols) * ldc + (i % cols)] = scale * sm[0][c];
}

int launch_partial_reduce_2d(const float* partial, int G, int64_t stride, int rows, int cols, float scale, float* out,
                             int ldc, cudaStream_t st) {
  partial_reduce2d_kernel<<<(rows * cols + 31) / 32, 1024, 0, st>>>(partial, G, stride, rows, cols, scale, out, ldc);
  return after_launch("partial_reduce_kernel");
}

int launch_partial_reduce(const float* partial, int G, int64_t stride, int n, float scale, float* out, cudaStream_t st) {
  partial_reduce2_kernel<<<(n + 31) / 32, 1024, 0, st>>>(partial, G, stride, n, scale, out);
  return after_launch("partial_reduce_kernel");
}

}  // namespace pev

extern "C" {
int pev_abi_version(void) { return PEV_ABI_VERSION; }
const char* pev_last_error(void) { return pev::g_err; }
int64_t pev_launch_count(void) { return pev::g_launches.load(); }
}
