// Ragged packed batches.  The reference pads every conformer to the batch maximum on the host
// (models/data.py:219-266) after centring it on its CA centroid (:166-172), and ships the padded tensors; here the host
// ships only the real rows (one contiguous pinned buffer per field + cu_seqlens) and ONE kernel centres and pads on
// the device: block per (conformer, row block), centroid in double, coalesced row writes, float4 copies of the
// sequence embeddings.
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_data_body.cuh"

namespace pev {

constexpr int kUnpackThreads = 256;

__global__ void __launch_bounds__(kUnpackThreads) unpack_center_kernel(const UnpackArgs a, int center) {
  __shared__ double red[32];
  __shared__ float cen[3];
  const int b = blockIdx.x;
  const int L = a.cu[b + 1] - a.cu[b];
  double s[3] = {0.0, 0.0, 0.0}, cnt = 0.0;
  if (center)
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
      const int64_t t = (int64_t)a.cu[b] + l;
      if (a.mask[t] != 0.f) {
        s[0] += a.ca[3 * t]; s[1] += a.ca[3 * t + 1]; s[2] += a.ca[3 * t + 2];
        cnt += 1.0;
      }
    }
  for (int k = 0; k < 3; ++k) s[k] = block_sum(s[k], red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0)
    for (int k = 0; k < 3; ++k) cen[k] = cnt > 0.0 ? (float)(s[k] / cnt) : 0.f;
  __syncthreads();
  for (int l = blockIdx.y * blockDim.x + threadIdx.x; l < a.Lmax; l += gridDim.y * blockDim.x) unpack_row(a, b, l, cen);
  if (a.emb) {
    // [Lmax, D] rows of this conformer: copy or zero, 16 bytes per thread when D allows
    const int64_t total = (int64_t)a.Lmax * a.D;
    float* dst = a.o_emb + (int64_t)b * total;
    const float* src = a.emb + (int64_t)a.cu[b] * a.D;
    const int64_t live = (int64_t)L * a.D;
    if ((a.D & 3) == 0) {
      for (int64_t i = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * 4; i < total; i += (int64_t)gridDim.y * blockDim.x * 4)
        *reinterpret_cast<float4*>(dst + i) = i < live ? *reinterpret_cast<const float4*>(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int64_t i = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.y * blockDim.x)
        dst[i] = i < live ? src[i] : 0.f;
    }
  }
}

}  // namespace pev

using namespace pev;

extern "C" int pev_unpack_center(const float* n, const float* ca, const float* c, const float* mask, const float* dih,
                                 const int64_t* labels, const float* emb, const int32_t* cu_seqlens, int32_t B, int32_t Lmax,
                                 int32_t D, int32_t center, float* o_n, float* o_ca, float* o_c, float* o_mask, float* o_dih,
                                 int64_t* o_labels, float* o_emb, void* stream) {
  PEV_REQUIRE(n && ca && c && mask && dih && labels && cu_seqlens && o_n && o_ca && o_c && o_mask && o_dih && o_labels,
              "null argument");
  PEV_REQUIRE(B >= 0 && Lmax >= 0 && D >= 0 && (emb == nullptr) == (o_emb == nullptr), "bad argument");
  if (B == 0 || Lmax == 0) return 0;
  UnpackArgs a = {n, ca, c, mask, dih, labels, emb, cu_seqlens, B, Lmax, D, o_n, o_ca, o_c, o_mask, o_dih, o_labels, o_emb};
  int ysplit = emb ? (int)(((int64_t)Lmax * D / 4 + kUnpackThreads * 8 - 1) / (kUnpackThreads * 8)) : 1;
  if (ysplit < 1) ysplit = 1;
  if (ysplit > 64) ysplit = 64;
  unpack_center_kernel<<<dim3(B, ysplit), kUnpackThreads, 0, as_stream(stream)>>>(a, center);
  return after_launch("unpack_center_kernel");
}
