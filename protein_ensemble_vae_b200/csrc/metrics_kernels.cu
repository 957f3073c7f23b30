// GPU evaluation metrics: scripts/validation_metrics.py's TM-score, lDDT, GDT and RMSF for whole
// ensembles at once -- the reference evaluates one structure pair per call in numpy (SVD + cdist + Python loops).
// The arithmetic lives in pev_metrics_body.cuh (host/device; tests/hostcheck runs the same bodies on the CPU).
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_metrics_body.cuh"

namespace pev {

// one thread per conformer: superposition onto its target, per-residue distances, TM / GDT scores
__global__ void __launch_bounds__(64)
superpose_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask, int S, int L,
                 int b_batch, int mask_batch, float* __restrict__ aligned, float* __restrict__ dist, float* __restrict__ tm,
                 float* __restrict__ gdt_ts, float* __restrict__ gdt_ha) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float* d = dist + (int64_t)s * L;
  superpose_serial(a + (int64_t)s * L * 3, b + (b_batch ? (int64_t)s * L * 3 : 0), L,
                   aligned ? aligned + (int64_t)s * L * 3 : nullptr, d);
  float t, g1, g2;
  superposition_scores(d, mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr, L, &t, &g1, &g2);
  if (tm) tm[s] = t;
  if (gdt_ts) gdt_ts[s] = g1;
  if (gdt_ha) gdt_ha[s] = g2;
}

// lDDT: block per conformer, both structures staged in shared memory, thread per residue
__global__ void __launch_bounds__(128)
lddt_kernel(const float* __restrict__ pred, const float* __restrict__ tru, const float* __restrict__ mask, int L, int t_batch,
            int mask_batch, float cutoff, float* __restrict__ per_res, float* __restrict__ global) {
  extern __shared__ float sm[];
  float* sp = sm;
  float* st = sm + 3 * L;
  float* smk = st + 3 * L;
  __shared__ float red[32];
  const int s = blockIdx.x;
  const float* gp = pred + (int64_t)s * L * 3;
  const float* gt = tru + (t_batch ? (int64_t)s * L * 3 : 0);
  const float* gm = mask ? mask + (mask_batch ? (int64_t)s * L : 0) : nullptr;
  for (int k = threadIdx.x; k < 3 * L; k += blockDim.x) { sp[k] = gp[k]; st[k] = gt[k]; }
  for (int k = threadIdx.x; k < L; k += blockDim.x) smk[k] = gm ? gm[k] : 1.f;
  __syncthreads();
  float sum = 0.f, cnt = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float v = lddt_residue(sp, st, smk, L, i, cutoff);
    per_res[(int64_t)s * L + i] = v;
    if (smk[i] != 0.f) { sum += v; cnt += 1.f; }
  }
  sum = block_sum(sum, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) global[s] = cnt > 0.f ? sum / cnt : 0.f;          // :146
}

__global__ void rmsf_kernel(const float* __restrict__ aligned, int N, int L, float* __restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L) out[l] = rmsf_residue(aligned, N, L, l);
}

}  // namespace pev

using namespace pev;

extern "C" {

int pev_superpose_scores(const float* a, const float* b, const float* mask, int32_t S, int32_t L, int32_t b_batch,
                         int32_t mask_batch, float* aligned, float* dist, float* tm, float* gdt_ts, float* gdt_ha,
                         void* stream) {
  PEV_REQUIRE(a && b && dist && S >= 0 && L > 0, "bad argument");
  if (S == 0) return 0;
  superpose_kernel<<<(S + 63) / 64, 64, 0, as_stream(stream)>>>(a, b, mask, S, L, b_batch, mask_batch, aligned, dist, tm,
                                                               gdt_ts, gdt_ha);
  return after_launch("superpose_kernel");
}

int pev_lddt(const float* pred, const float* tru, const float* mask, int32_t S, int32_t L, int32_t t_batch,
             int32_t mask_batch, float cutoff, float* per_residue, float* global, void* stream) {
  PEV_REQUIRE(pred && tru && per_residue && global && S >= 0 && L > 0, "bad argument");
  if (S == 0) return 0;
  const size_t smem = sizeof(float) * 7 * (size_t)L;
  PEV_REQUIRE(smem <= 200 * 1024, "structure too long for the shared-memory tile");
  if (smem > 40 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lddt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(2, "lddt_kernel: %s", cudaGetErrorString(e));
  }
  lddt_kernel<<<S, 128, smem, as_stream(stream)>>>(pred, tru, mask, L, t_batch, mask_batch, cutoff, per_residue, global);
  return after_launch("lddt_kernel");
}

int pev_rmsf(const float* aligned, int32_t N, int32_t L, float* out, void* stream) {
  PEV_REQUIRE(aligned && out && N > 0 && L > 0, "bad argument");
  rmsf_kernel<<<(L + 127) / 128, 128, 0, as_stream(stream)>>>(aligned, N, L, out);
  return after_launch("rmsf_kernel");
}

}  // extern "C"
