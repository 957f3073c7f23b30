// K4: batched Kabsch RMSD, one warp per conformer.  Three passes over the conformer's (cached)
// coordinates: centroids, 3x3 covariance, residual after rotation; the 3x3 rotation is solved in
// double precision by lane 0 (pev_kabsch_body.cuh) and broadcast.  HBM-bound: 12 L bytes per
// conformer (+4 L for a per-conformer mask); the shared reference structure stays in L2.
// Reference: generate_ensemble_pdbs.py:343-373, scripts/validation_metrics.py:57-85.
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_kabsch_body.cuh"

namespace pev {

__global__ void __launch_bounds__(128)
kabsch_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int S,
              int L, int b_batch, int mask_batch, int mode, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const float* pa = a + (int64_t)s * L * 3;
  const float* pb = b + (b_batch ? (int64_t)s * L * 3 : 0);
  const float* pm = mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr;
  // pass 1: centroids
  double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
  int n = 0;
  for (int l = lane; l < L; l += 32) {
    if (pm && pm[l] == 0.f) continue;
    ++n;
#pragma unroll
    for (int k = 0; k < 3; ++k) { ca[k] += pa[3 * l + k]; cb[k] += pb[3 * l + k]; }
  }
  n = warp_sum(n);
  if (n == 0) {
    if (lane == 0) out[s] = 0.f;
    return;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { ca[k] = warp_sum(ca[k]) / n; cb[k] = warp_sum(cb[k]) / n; }
  // pass 2: covariance of the centred sets
  double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int l = lane; l < L; l += 32) {
    if (pm && pm[l] == 0.f) continue;
    double p[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = pa[3 * l + k] - ca[k]; q[k] = pb[3 * l + k] - cb[k]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) H[i][j] += p[i] * q[j];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) H[i][j] = warp_sum(H[i][j]);
  double R[3][3];
  if (lane == 0) kabsch_rotation(H, R);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = __shfl_sync(0xffffffffu, R[i][j], 0);
  // pass 3: residual
  double e = 0.0;
  for (int l = lane; l < L; l += 32) {
    if (pm && pm[l] == 0.f) continue;
    double p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = pa[3 * l + k] - ca[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double q = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) q += (mode == 0 ? R[i][k] : R[k][i]) * p[k];
      double d = q - (pb[3 * l + i] - cb[i]);
      e += d * d;
    }
  }
  e = warp_sum(e);
  if (lane == 0) out[s] = (float)sqrt(e / n);
}

}  // namespace pev

extern "C" int pev_kabsch_rmsd(const float* a, const float* b, const float* mask, int32_t S, int32_t L,
                               int32_t b_batch, int32_t mask_batch, int32_t mode, float* out, void* stream) {
  using namespace pev;
  PEV_REQUIRE(a && b && out && S >= 0 && L > 0 && (mode == 0 || mode == 1), "bad argument");
  if (S == 0) return 0;
  kabsch_kernel<<<(S + 3) / 4, 128, 0, as_stream(stream)>>>(a, b, mask, S, L, b_batch, mask_batch, mode, out);
  return after_launch("kabsch_kernel");
}
