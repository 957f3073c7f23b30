// K4: batched Kabsch RMSD.  Three passes over a conformer's (cached) coordinates: centroids, 3x3 covariance, residual after
// rotation, each spread over a warp; a warp takes a GROUP of conformers and solves their 3x3 rotations lane-parallel in
// double precision (pev_kabsch_body.cuh), one lane per conformer.  HBM-bound: 12 L bytes per conformer (+4 L for a
// per-conformer mask); the shared reference structure stays in L2.
// Reference: generate_ensemble_pdbs.py:343-373, scripts/validation_metrics.py:57-85.
#include <cstdlib>
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_kabsch_body.cuh"

namespace pev {

// RMSDs of up to 32 pairs of structures after Kabsch superposition, computed by one warp.  The warp is split into four
// SLOTS of eight lanes; a slot streams one pair at a time (pairs c = slot, slot + 4, ... of the warp's group), so every
// reduction is three shuffle steps shared by four pairs, and the 3 x 3 solves of ALL the warp's pairs run at once, lane
// `k` of a slot solving the slot's k-th pair.  Why: B200 issues fp64 at 1/64 of the fp32 rate.  With one pair per warp
// (round 1 / early round 2) the 17 double-precision butterfly sums per pair (85 warp-wide DADDs) and lane 0's Jacobi sweeps
// with 31 lanes idle made the kernel fp64-bound at 3 - 4 % of the HBM roofline; here a pair costs 13 warp-wide DADDs and
// 1/32 of a warp-wide solve.  The per-residue arithmetic is fp32 (a lane keeps ceil(L / 8) terms per partial sum), the
// cross-lane sums and the eigen-solve are double.
// `pair(c, pa, pb, pm)` yields the pointers of the group's c-th pair (called with c < count only); lane 8 * slot + k
// returns the RMSD of pair slot + 4 k (0 beyond count).
constexpr int KB_W = 8;                 // lanes per slot
constexpr int KB_SLOTS = 32 / KB_W;     // pairs in flight per warp

__device__ __forceinline__ double slot_sum(double v) {
#pragma unroll
  for (int o = KB_W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int slot_sum(int v) {
#pragma unroll
  for (int o = KB_W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <class PairFn>
__device__ __forceinline__ float kabsch_group(PairFn pair, int count, int L, int mode, int lane) {
  const int slot = lane / KB_W, sub = lane % KB_W;
  const int steps = (count + KB_SLOTS - 1) / KB_SLOTS;       // <= 8: pairs per slot
  double Hm[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  float cam[3] = {0.f, 0.f, 0.f}, cbm[3] = {0.f, 0.f, 0.f};
  int nm = 0;
  for (int k = 0; k < steps; ++k) {
    const int c = slot + KB_SLOTS * k;
    const bool active = c < count;
    const float *pa = nullptr, *pb = nullptr, *pm = nullptr;
    if (active) pair(c, pa, pb, pm);
    // pass 1: centroids
    float sa[3] = {0.f, 0.f, 0.f}, sb[3] = {0.f, 0.f, 0.f};
    int n = 0;
    if (active)
      for (int l = sub; l < L; l += KB_W) {
        if (pm && pm[l] == 0.f) continue;
        ++n;
#pragma unroll
        for (int j = 0; j < 3; ++j) { sa[j] += pa[3 * l + j]; sb[j] += pb[3 * l + j]; }
      }
    n = slot_sum(n);
    const double inv_n = n > 0 ? 1.0 / n : 0.0;
    float caf[3], cbf[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      caf[j] = (float)(slot_sum((double)sa[j]) * inv_n);
      cbf[j] = (float)(slot_sum((double)sb[j]) * inv_n);
    }
    // pass 2: covariance of the centred sets
    float Hf[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (active && n > 0)
      for (int l = sub; l < L; l += KB_W) {
        if (pm && pm[l] == 0.f) continue;
        float p[3], q[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { p[j] = pa[3 * l + j] - caf[j]; q[j] = pb[3 * l + j] - cbf[j]; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) Hf[i][j] = fmaf(p[i], q[j], Hf[i][j]);
      }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double h = slot_sum((double)Hf[i][j]);
        if (sub == k) Hm[i][j] = h;
      }
    if (sub == k) {
      nm = n;
#pragma unroll
      for (int j = 0; j < 3; ++j) { cam[j] = caf[j]; cbm[j] = cbf[j]; }
    }
  }
  double Rm[3][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}};
  if (nm > 0) kabsch_rotation(Hm, Rm);                       // all the warp's solves at once, one per lane
  __syncwarp();
  float result = 0.f;
  for (int k = 0; k < steps; ++k) {
    const int c = slot + KB_SLOTS * k;
    const int n = __shfl_sync(0xffffffffu, nm, k, KB_W);     // from lane k of the own slot
    // pass 3: residual
    float Rf[3][3], caf[3], cbf[3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) Rf[i][j] = (float)__shfl_sync(0xffffffffu, mode == 0 ? Rm[i][j] : Rm[j][i], k, KB_W);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      caf[j] = __shfl_sync(0xffffffffu, cam[j], k, KB_W);
      cbf[j] = __shfl_sync(0xffffffffu, cbm[j], k, KB_W);
    }
    float ef = 0.f;
    if (c < count && n > 0) {
      const float *pa, *pb, *pm;
      pair(c, pa, pb, pm);
      for (int l = sub; l < L; l += KB_W) {
        if (pm && pm[l] == 0.f) continue;
        float p[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) p[j] = pa[3 * l + j] - caf[j];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const float q = fmaf(Rf[i][2], p[2], fmaf(Rf[i][1], p[1], Rf[i][0] * p[0]));
          const float d = q - (pb[3 * l + i] - cbf[i]);
          ef = fmaf(d, d, ef);
        }
      }
    }
    const double e = slot_sum((double)ef);
    if (sub == k && n > 0) result = (float)sqrt(e / n);
  }
  return result;
}

// index (within the warp's group) of the pair whose result `lane` holds
__device__ __forceinline__ int kabsch_result_index(int lane) { return lane / KB_W + KB_SLOTS * (lane % KB_W); }

// each warp takes `group` (<= 32) consecutive conformers
__global__ void __launch_bounds__(128)
kabsch_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int S,
              int L, int b_batch, int mask_batch, int mode, int group, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t s0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * group;
  if (s0 >= S) return;
  const int count = (int)(S - s0 < group ? S - s0 : group);
  auto pair = [&](int c, const float*& pa, const float*& pb, const float*& pm) {
    const int64_t s = s0 + c;
    pa = a + s * L * 3;
    pb = b + (b_batch ? s * L * 3 : 0);
    pm = mask ? mask + (mask_batch ? s * L : 0) : nullptr;
  };
  const float r = kabsch_group(pair, count, L, mode, lane);
  const int c = kabsch_result_index(lane);
  if (c < count) out[s0 + c] = r;
}

// all pairs i < j of one ensemble (the diversity loop of generate_ensemble_pdbs.py:591-595): each warp takes `group`
// consecutive pair indices; out[S,S] gets the value at (i,j) and (j,i), zeros on the diagonal
__device__ __forceinline__ void pair_of_index(int64_t pidx, int S, int64_t& i, int64_t& j) {
  // pair index -> (i, j), i < j, rows of the strict upper triangle in order
  i = (int64_t)((2.0 * S - 1.0 - sqrt((2.0 * S - 1.0) * (2.0 * S - 1.0) - 8.0 * (double)pidx)) * 0.5);
  while (i > 0 && i * (2 * S - i - 1) / 2 > pidx) --i;
  while ((i + 1) * (2 * S - i - 2) / 2 <= pidx) ++i;
  j = pidx - i * (2 * S - i - 1) / 2 + i + 1;
}

__global__ void __launch_bounds__(128)
kabsch_pairs_kernel(const float* __restrict__ a, const float* __restrict__ mask, int S, int L, int mode, int group,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t p0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * group;
  const int64_t npairs = (int64_t)S * (S - 1) / 2;
  if (p0 >= npairs) return;
  const int count = (int)(npairs - p0 < group ? npairs - p0 : group);
  auto pair = [&](int c, const float*& pa, const float*& pb, const float*& pm) {
    int64_t i, j;
    pair_of_index(p0 + c, S, i, j);
    pa = a + i * L * 3;
    pb = a + j * L * 3;
    pm = mask;
  };
  const float r = kabsch_group(pair, count, L, mode, lane);
  const int c = kabsch_result_index(lane);
  if (c < count) {
    int64_t i, j;
    pair_of_index(p0 + c, S, i, j);
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

// Pairs per warp: as many as keep every SM supplied with warps (the solves of a warp's pairs run lane-parallel, so larger
// groups waste less of the fp64 pipe; small problems keep one pair per warp).  PEV_KABSCH_GROUP overrides (measurements).
static int kabsch_group_size(int64_t pairs) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("PEV_KABSCH_GROUP");
    forced = e ? atoi(e) : 0;
    if (forced < 0 || forced > 32) forced = 0;
    if (forced) forced = (forced + 3) / 4 * 4;
  }
  if (forced) return forced;
  const int64_t want_warps = (int64_t)pev::sm_count() * 64;   // measured: 100 000 x L=100 is fastest at 8 per warp (84 warps / SM)
  int g = 32;
  while (g > 4 && pairs / g < want_warps) g >>= 1;          // at least one pair per 8-lane slot
  return g;
}

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
  const int group = kabsch_group_size(S);
  const int64_t warps = ((int64_t)S + group - 1) / group;
  kabsch_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, as_stream(stream)>>>(a, b, mask, S, L, b_batch, mask_batch, mode, group,
                                                                          out);
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
  const int group = kabsch_group_size(npairs);
  const int64_t warps = (npairs + group - 1) / group;
  kabsch_pairs_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(a, mask, S, L, mode, group, out);
  return after_launch("kabsch_pairs_kernel");
}
