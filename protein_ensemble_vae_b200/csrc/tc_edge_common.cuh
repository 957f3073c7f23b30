// Helpers shared by the tcgen05 edge kernels (edge_tc2_kernels.cu: single-CTA kernels; edge_tc3_kernels.cu: fused CTA-pair
// kernels): bf16 / fp16 packing, warp-level sums, per-feature vector layout, TMA tensor maps of [rows, 256] bf16 matrices.
#pragma once
#include <cuda.h>          // CUtensorMap types only: the encoder is resolved at run time (no libcuda link dependency)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pev_common.cuh"
#include "tc_common.cuh"

namespace pev {
namespace tce {
using namespace tcx;

__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  return make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
}
__device__ __forceinline__ void unpack8(const uint4 v, float (&o)[8]) {
  o[0] = bf16_lo(v.x); o[1] = bf16_hi(v.x); o[2] = bf16_lo(v.y); o[3] = bf16_hi(v.y);
  o[4] = bf16_lo(v.z); o[5] = bf16_hi(v.z); o[6] = bf16_lo(v.w); o[7] = bf16_hi(v.w);
}
// The node projection ABh is staged in HBM as fp16 (11-bit significand: its rounding adds to the first edge
// linear's output, and bf16 would double the path's error); A_i + B_j is formed with one packed add and widened.
__device__ __forceinline__ void add_f16x2_to_f32(uint32_t a, uint32_t b, float& lo, float& hi) {
  uint32_t r;
  asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(lo), "=f"(hi)
      : "r"(r));
}
__device__ __forceinline__ void add_f16x8_to_f32(const uint4 a, const uint4 b, float (&o)[8]) {
  add_f16x2_to_f32(a.x, b.x, o[0], o[1]);
  add_f16x2_to_f32(a.y, b.y, o[2], o[3]);
  add_f16x2_to_f32(a.z, b.z, o[4], o[5]);
  add_f16x2_to_f32(a.w, b.w, o[6], o[7]);
}
__device__ __forceinline__ float sum32(const float (&m)[32]) {
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = (m[j] + m[j + 8]) + (m[j + 16] + m[j + 24]);
  return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}
// sum of m[j] over the bits set in `msk` (warp-uniform mask; no branches)
__device__ __forceinline__ float masked_sum32(const float (&m)[32], uint32_t msk) {
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 32; ++j) s[j & 3] += ((msk >> j) & 1u) ? m[j] : 0.f;
  return (s[0] + s[1]) + (s[2] + s[3]);
}
// Layout of a per-feature vector that the producers read as two float4 per (K-chunk, 16-byte column chunk): the first
// (second) halves of the eight column chunks of a K-chunk are contiguous, so a quarter-warp's eight 16-byte reads cover
// 128 contiguous bytes (conflict-free) instead of eight 16-byte pieces 32 bytes apart (2-way bank conflict).
__device__ __forceinline__ int vec_slot(int k) { return ((k >> 6) * 2 + ((k >> 2) & 1)) * 32 + ((k >> 3) & 7) * 4 + (k & 3); }
// 32-byte global load (LDG.256): two 16-byte chunks
__device__ __forceinline__ void ld_256(const void* addr, uint4& a, uint4& b) {
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(addr));
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the library must load without libcuda.so)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// tensor map of a bf16 [rows, 256] row-major matrix with a [32 rows x 32 columns] box (64-byte inner extent, SWIZZLE_64B)
inline int make_rows_map(void* base, int64_t rows, CUtensorMap* out) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
      return set_error(2, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {256, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {256 * 2};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(2, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}


}  // namespace tce
}  // namespace pev
