// K4: batched Kabsch RMSD, one warp per conformer.  Three passes over the conformer's (cached)
// coordinates: centroids, 3x3 covariance, residual after rotation; the 3x3 rotation is solved in
// double precision by lane 0 (pev_kabsch_body.cuh) and broadcast.  HBM-bound: 12 L bytes per
// conformer (+4 L for a per-conformer mask); the shared reference structure stays in L2.
// Reference: generate_ensemble_pdbs.py:343-373, scripts/validation_metrics.py:57-85.
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_kabsch_body.cuh"

namespace pev {

// RMSD of one pair of structures after Kabsch superposition, computed by one warp; returns the value in every lane
__device__ __forceinline__ float kabsch_pair(const float* __restrict__ pa, const float* __restrict__ pb,
                                             const float* __restrict__ pm, int L, int mode, int lane) {
  // The per-residue arithmetic runs in fp32 (B200 issues fp64 at 1/64 of the fp32 rate: the all-double form of round 1
  // was fp64-bound at 3 % of the HBM roofline); every lane keeps at most ceil(L / 32) terms per partial sum, the warp
  // reductions and the 3 x 3 eigen-solve stay in double.
  // pass 1: centroids
  float sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f};
  int n = 0;
  for (int l = lane; l < L; l += 32) {
    if (pm && pm[l] == 0.f) continue;
    ++n;
#pragma unroll
    for (int k = 0; k < 3; ++k) { sa[k] += pa[3 * l + k]; sb[k] += pb[3 * l + k]; }
  }
  n = warp_sum(n);
  if (n == 0) return 0.f;
  double ca[3], cb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { ca[k] = warp_sum((double)sa[k]) / n; cb[k] = warp_sum((double)sb[k]) / n; }
  const float caf[3] = {(float)ca[0], (float)ca[1], (float)ca[2]}, cbf[3] = {(float)cb[0], (float)cb[1], (float)cb[2]};
  // pass 2: covariance of the centred sets
  float Hf[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
  for (int l = lane; l < L; l += 32) {
    if (pm && pm[l] == 0.f) continue;
    float p[3], q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { p[k] = pa[3 * l + k] - caf[k]; q[k] = pb[3 * l + k] - cbf[k]; }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Hf[i][j] = fmaf(p[i], q[j], Hf[i][j]);
  }
  double H[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) H[i][j] = warp_sum((double)Hf[i][j]);
  double R[3][3];
  if (lane == 0) kabsch_rotation(H, R);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = __shfl_sync(0xffffffffu, R[i][j], 0);
  // pass 3: residual
  float Rf[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) Rf[i][k] = (float)(mode == 0 ? R[i][k] : R[k][i]);
  float ef = 0.f;
  for (int l = lane; l < L; l += 32) {
    if (pm && pm[l] == 0.f) continue;
    float p[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = pa[3 * l + k] - caf[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const float q = fmaf(Rf[i][2], p[2], fmaf(Rf[i][1], p[1], Rf[i][0] * p[0]));
      const float d = q - (pb[3 * l + i] - cbf[i]);
      ef = fmaf(d, d, ef);
    }
  }
  double e = (double)ef;
  e = warp_sum(e);
  return (float)sqrt(e / n);
}

__global__ void __launch_bounds__(128)
kabsch_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int S,
              int L, int b_batch, int mask_batch, int mode, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (s >= S) return;
  const float* pa = a + (int64_t)s * L * 3;
  const float* pb = b + (b_batch ? (int64_t)s * L * 3 : 0);
  const float* pm = mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr;
  const float r = kabsch_pair(pa, pb, pm, L, mode, lane);
  if (lane == 0) out[s] = r;
}

// all pairs i < j of one ensemble (the diversity loop of generate_ensemble_pdbs.py:591-595): one warp per pair;
// out[S,S] gets the value at (i,j) and (j,i), zeros on the diagonal
__global__ void __launch_bounds__(128)
kabsch_pairs_kernel(const float* __restrict__ a, const float* __restrict__ mask, int S, int L, int mode,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t pidx = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t npairs = (int64_t)S * (S - 1) / 2;
  if (pidx >= npairs) {
    return;
  }
  // pair index -> (i, j), i < j, rows of the strict upper triangle in order
  int64_t i = (int64_t)((2.0 * S - 1.0 - sqrt((2.0 * S - 1.0) * (2.0 * S - 1.0) - 8.0 * (double)pidx)) * 0.5);
  while (i > 0 && i * (2 * S - i - 1) / 2 > pidx) --i;
  while ((i + 1) * (2 * S - i - 2) / 2 <= pidx) ++i;
  const int64_t j = pidx - i * (2 * S - i - 1) / 2 + i + 1;
  const float r = kabsch_pair(a + i * L * 3, a + j * L * 3, mask, L, mode, lane);
  if (lane == 0) {
    out[i * S + j] = r;
    out[j * S + i] = r;
  }
}

// geometry validity filter: one thread per conformer (12 L bytes each; 100 000 x L=100 = 120 MB, latency-trivial)
__global__ void __launch_bounds__(128)
validate_geometry_kernel(const float* __restrict__ ca, const float* __restrict__ mask, int S, int L, int mask_batch,
                         int32_t* __restrict__ status, float* __restrict__ stats) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float st[3];
  status[s] = validate_geometry_serial(ca + (int64_t)s * L * 3, mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr,
                                       L, st);
  if (stats) { stats[3 * s] = st[0]; stats[3 * s + 1] = st[1]; stats[3 * s + 2] = st[2]; }
}

}  // namespace pev

extern "C" int pev_validate_geometry(const float* ca, const float* mask, int32_t S, int32_t L, int32_t mask_batch,
                                     int32_t* status, float* stats, void* stream) {
  using namespace pev;
  PEV_REQUIRE(ca && status && S >= 0 && L > 0, "bad argument");
  if (S == 0) return 0;
  validate_geometry_kernel<<<(S + 127) / 128, 128, 0, as_stream(stream)>>>(ca, mask, S, L, mask_batch, status, stats);
  return after_launch("validate_geometry_kernel");
}

extern "C" int pev_kabsch_rmsd(const float* a, const float* b, const float* mask, int32_t S, int32_t L,
                               int32_t b_batch, int32_t mask_batch, int32_t mode, float* out, void* stream) {
  using namespace pev;
  PEV_REQUIRE(a && b && out && S >= 0 && L > 0 && (mode == 0 || mode == 1), "bad argument");
  if (S == 0) return 0;
  kabsch_kernel<<<(S + 3) / 4, 128, 0, as_stream(stream)>>>(a, b, mask, S, L, b_batch, mask_batch, mode, out);
  return after_launch("kabsch_kernel");
}

extern "C" int pev_kabsch_rmsd_pairs(const float* a, const float* mask, int32_t S, int32_t L, int32_t mode, float* out,
                                     void* stream) {
  using namespace pev;
  PEV_REQUIRE(a && out && S >= 0 && L > 0 && (mode == 0 || mode == 1), "bad argument");
  if (S == 0) return 0;
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(out, 0, sizeof(float) * (size_t)S * S, st);
  const int64_t npairs = (int64_t)S * (S - 1) / 2;
  if (npairs == 0) return 0;
  kabsch_pairs_kernel<<<(unsigned)((npairs + 3) / 4), 128, 0, st>>>(a, mask, S, L, mode, out);
  return after_launch("kabsch_pairs_kernel");
}
