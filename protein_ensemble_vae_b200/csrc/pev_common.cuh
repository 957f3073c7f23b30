// Runtime helpers shared by the .cu translation units (CUDA only; not used by tests/hostcheck).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pev {

// thread-local message returned by pev_last_error(); returns a non-zero code for convenience
int set_error(int code, const char* fmt, ...);
// counts one kernel launch and converts cudaGetLastError() into the ABI's error code
int after_launch(const char* kernel_name);
int sm_count();            // of the CURRENT device
constexpr int kMaxDevices = 64;
int current_device();      // cudaGetDevice(), clamped to [0, kMaxDevices)

// out[i] = scale * sum_g partial[g * stride + i], i < n, summed in a FIXED order (32 interleaved chains per column, then
// a fixed tree): the second stage of every per-CTA-partial reduction (bit-reproducible, no atomics).  capi.cu
int launch_partial_reduce(const float* partial, int G, int64_t stride, int n, float scale, float* out, cudaStream_t st);
int launch_partial_reduce_2d(const float* partial, int G, int64_t stride, int rows, int cols, float scale, float* out,
                             int ldc, cudaStream_t st);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum `v` over the block; result valid in thread 0.  `red` is a __shared__ array of >= 32 T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();            // protect `red` from a previous use
  if (lane == 0) red[wid] = v;
  __syncthreads();
  T r = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
  if (wid == 0) r = warp_sum(r);
  return r;
}

struct AtomicAddD {
  __device__ __forceinline__ void operator()(double* p, double v) const {
    if (v != 0.0) atomicAdd(p, v);
  }
};

}  // namespace pev

#define PEV_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) return pev::set_error(1, "%s: %s", __func__, msg); \
  } while (0)
