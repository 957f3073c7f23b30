// K3: fused loss kernels (forward accumulators, finalisation, backward).  The arithmetic lives in
// pev_loss_body.cuh / pev_loss_final.cuh (host/device, checked on the CPU by tests/hostcheck);
// this file holds the launch geometry, shared-memory staging and reductions.
//
// Reference: models/losses.py:12-613.  HBM-bound streams (K3a) and on-chip pair tiles (K3b);
// DESIGN.md gives the algorithmic bytes per residue.
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_loss_body.cuh"
#include "pev_loss_final.cuh"

namespace pev {

constexpr int kResThreads = 128;
constexpr int kPairThreads = 128;
constexpr int kClashThreads = 128;
constexpr int kMaxDynSmem = 200 * 1024;

// ------------------------------------------------------------------------------------------ K3a fwd
__global__ void __launch_bounds__(kResThreads)
loss_residue_fwd_kernel(pev_loss_args A, double* __restrict__ ag, double* __restrict__ as) {
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int i = blockIdx.x * kResThreads + threadIdx.x;
  float acc[RA_COUNT];
#pragma unroll
  for (int k = 0; k < RA_COUNT; ++k) acc[k] = 0.f;
  if (i < A.L) residue_fwd(A, b, i, acc);
  float tot[RA_COUNT];
#pragma unroll
  for (int k = 0; k < RA_COUNT; ++k) tot[k] = block_sum(acc[k], red);
  if (threadIdx.x == 0) scatter_residue_acc_t(tot, ag, as + 8 * (int64_t)b, AtomicAddD());
}

// KL rows: one warp per row of D values, grid-stride over rows; sum_rows w_row * sum_d kl(mu, lv)
__global__ void __launch_bounds__(256)
loss_kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                   const float* __restrict__ row_weight, int64_t rows, int D, double* __restrict__ out) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  double acc = 0.0;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const float* m = mu + r * D;
    const float* l = lv + r * D;
    float s = 0.f;
    if ((D & 3) == 0) {
      const float4* m4 = reinterpret_cast<const float4*>(m);
      const float4* l4 = reinterpret_cast<const float4*>(l);
      for (int k = lane; k < (D >> 2); k += 32) {
        float4 a = __ldg(m4 + k), c = __ldg(l4 + k);
        s += kl_elem(a.x, c.x) + kl_elem(a.y, c.y) + kl_elem(a.z, c.z) + kl_elem(a.w, c.w);
      }
    } else {
      for (int k = lane; k < D; k += 32) s += kl_elem(m[k], l[k]);
    }
    s = warp_sum(s);
    if (lane == 0) acc += (double)(s * (row_weight ? row_weight[r] : 1.0f));
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && acc != 0.0) atomicAdd(out, acc);
}

// ------------------------------------------------------------------------------------------ K3b
// strided pair-distance tile: selected points of one conformer staged in shared memory
__global__ void __launch_bounds__(kPairThreads)
loss_pair_kernel(pev_loss_args A, int M, double* __restrict__ ag, const float* __restrict__ coef,
                 const float* __restrict__ inv_den, float* __restrict__ gCA) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  float* P = sm;
  float* T = sm + 3 * M;
  float* pm = sm + 6 * M;
  const int b = blockIdx.y;
  for (int s = threadIdx.x; s < M; s += blockDim.x) {
    const int64_t bi = (int64_t)b * A.L + (int64_t)s * A.pair_stride;
    for (int k = 0; k < 3; ++k) {
      P[3 * s + k] = A.pred_CA[bi * 3 + k];
      T[3 * s + k] = A.target_CA[bi * 3 + k];
    }
    pm[s] = A.mask[bi];
  }
  __syncthreads();
  const int i = blockIdx.x * kPairThreads + threadIdx.x;
  float n = 0.f, d = 0.f;
  if (gCA == nullptr) {
    if (i < M) pair_row(P, T, pm, M, i, &n, &d, nullptr, 0.f);
    n = block_sum(n, red);
    d = block_sum(d, red);
    if (threadIdx.x == 0) {
      AtomicAddD add;
      add(ag + 2 * PEV_T_PAIR, (double)n);
      add(ag + 2 * PEV_T_PAIR + 1, (double)d);
    }
  } else if (i < M) {
    v3 g;
    pair_row(P, T, pm, M, i, &n, &d, &g, coef[PEV_T_PAIR] * inv_den[PEV_T_PAIR]);
    float* dst = gCA + ((int64_t)b * A.L + (int64_t)i * A.pair_stride) * 3;
    dst[0] += g.x; dst[1] += g.y; dst[2] += g.z;
  }
}

// clash tile (models/losses.py:439-517).  One block per (conformer, 256-atom row tile): all 3L backbone atoms of the
// conformer are staged once as float4 {x, y, z, mask} (one broadcast LDS.128 per partner and warp), every thread keeps
// kClashIPT of the tile's atoms in registers and walks ALL partners, so a partner load is shared by kClashIPT pair
// evaluations and the hot loop is 3 subtractions + 3 FMAs + one compare per pair: the distance test d^2 < clash_dist^2
// is the only thing evaluated for the >99 % of pairs that are far apart; the rare path (sqrt, residue-separation rule
// |res_i - res_j| >= 2, penalty, force) runs for close pairs only -- bonded neighbours enter it and are dropped there.
// The pair count (denominator) needs no loop: m_a (sum of all masks - masks of the nine atoms of residues r-1..r+1).
// Walking the full row (both triangles) keeps the gradient atomic-free and deterministic: thread a owns atom a.
// Forward: numerator / denominator per conformer, each unordered pair counted twice (pev_loss_final.cuh).
// Backward: the same walk, accumulating the force on the thread's own atoms.
constexpr int kClashIPT = 2;                               // atoms per thread
constexpr int kClashTile = kClashThreads * kClashIPT;      // 256 atoms per block (L = 256: three full blocks)
template <bool BACKWARD>
__global__ void __launch_bounds__(kClashThreads)
loss_clash_kernel(pev_loss_args A, double* __restrict__ as, const float* __restrict__ coef,
                  const float* __restrict__ inv_den, float* __restrict__ gN, float* __restrict__ gCA,
                  float* __restrict__ gC) {
  extern __shared__ float4 sat[];                          // [3L] {x, y, z, mask}
  __shared__ float red[32];
  const int n_atoms = 3 * A.L;
  const int b = blockIdx.y;
  float msum_part = 0.f;
  for (int idx = threadIdx.x; idx < n_atoms; idx += blockDim.x) {
    const int i = idx / 3, k = idx - 3 * i;
    const int64_t bi = (int64_t)b * A.L + i;
    const float* src = (k == 0 ? A.pred_N : (k == 1 ? A.pred_CA : A.pred_C)) + bi * 3;
    const float m = A.mask[bi];
    sat[idx] = make_float4(src[0], src[1], src[2], m);
    msum_part += m;
  }
  const float msum = block_sum(msum_part, red);            // valid in thread 0
  __shared__ float msum_all;
  if (threadIdx.x == 0) msum_all = msum;
  __syncthreads();
  const float cd = A.clash_dist, cd2 = cd * cd, sm = A.soft_margin;
  float4 pa[kClashIPT];
  int ai[kClashIPT];
  float num[kClashIPT];
  float gx[kClashIPT], gy[kClashIPT], gz[kClashIPT];
#pragma unroll
  for (int u = 0; u < kClashIPT; ++u) {
    ai[u] = blockIdx.x * kClashTile + u * kClashThreads + threadIdx.x;
    pa[u] = ai[u] < n_atoms ? sat[ai[u]] : make_float4(1e30f, 1e30f, 1e30f, 0.f);   // far from everything
    num[u] = gx[u] = gy[u] = gz[u] = 0.f;
  }
  auto close_pair = [&](int u, int c, const float4 pc, float dx, float dy, float dz, float d2) {
    const int sep = c / 3 - ai[u] / 3;
    if (sep < 2 && sep > -2) return;                                                // :478-482
    const float w = pa[u].w * pc.w;
    const float dist = sqrtf(d2);
    const float v = fmaxf(cd - dist, 0.f);                                          // :494-495
    num[u] += (v < sm ? 0.5f * v * v : v * v) * w;                                  // :500-504
    if (BACKWARD && v > 0.f && dist > 0.f) {
      const float f = (v < sm ? v : 2.0f * v) * w / dist;
      gx[u] -= dx * f; gy[u] -= dy * f; gz[u] -= dz * f;
    }
  };
#pragma unroll 4
  for (int c = 0; c < n_atoms; ++c) {
    const float4 pc = sat[c];
#pragma unroll
    for (int u = 0; u < kClashIPT; ++u) {
      const float dx = pa[u].x - pc.x, dy = pa[u].y - pc.y, dz = pa[u].z - pc.z;
      const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      if (d2 < cd2) close_pair(u, c, pc, dx, dy, dz, d2);
    }
  }
  if (!BACKWARD) {
    float n = 0.f, d = 0.f;
#pragma unroll
    for (int u = 0; u < kClashIPT; ++u) {
      if (ai[u] >= n_atoms) continue;
      n += num[u];
      const int r = ai[u] / 3;
      float near = 0.f;                                    // masks of the atoms of residues r-1, r, r+1 (excluded pairs)
      for (int rr = max(r - 1, 0); rr <= min(r + 1, A.L - 1); ++rr) near += 3.f * sat[3 * rr].w;
      d += pa[u].w * (msum_all - near);
    }
    n = block_sum(n, red);
    d = block_sum(d, red);
    if (threadIdx.x == 0) {
      AtomicAddD add;
      add(as + 8 * (int64_t)b + 4, (double)n);
      add(as + 8 * (int64_t)b + 5, (double)d);
    }
  } else {
    const float cf = coef[PEV_T_CLASH] * inv_den[PEV_NUM_TERMS + A.B + b];
#pragma unroll
    for (int u = 0; u < kClashIPT; ++u) {
      const int a = ai[u];
      if (a >= n_atoms) continue;
      float* base = (a % 3 == 0) ? gN : ((a % 3 == 1) ? gCA : gC);
      if (base) {
        float* dst = base + ((int64_t)b * A.L + a / 3) * 3;
        dst[0] += gx[u] * cf; dst[1] += gy[u] * cf; dst[2] += gz[u] * cf;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ finalize
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const double* __restrict__ ag, const double* __restrict__ as, int B,
                     float* __restrict__ terms, float* __restrict__ inv_den) {
  __shared__ double red[32];
  double part[FIN_PARTS];
#pragma unroll
  for (int k = 0; k < FIN_PARTS; ++k) part[k] = 0.0;
  finalize_partial(as, B, threadIdx.x, blockDim.x, part, inv_den);
  double tot[FIN_PARTS];
#pragma unroll
  for (int k = 0; k < FIN_PARTS; ++k) tot[k] = block_sum(part[k], red);
  if (threadIdx.x == 0) finalize_combine(ag, tot, B, terms, inv_den);
}

// ------------------------------------------------------------------------------------------ K3a bwd
__global__ void __launch_bounds__(kResThreads)
loss_residue_bwd_kernel(pev_loss_args A, const float* __restrict__ coef, const float* __restrict__ inv_den,
                        float* __restrict__ gN, float* __restrict__ gCA, float* __restrict__ gC,
                        float* __restrict__ glog) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * kResThreads + threadIdx.x;
  if (i >= A.L) return;
  float cf[PEV_NUM_TERMS];
#pragma unroll
  for (int t = 0; t < PEV_NUM_TERMS; ++t) cf[t] = coef[t] * inv_den[t];
  float cf_rec[3];
  const float ids = inv_den[PEV_NUM_TERMS + b];
  cf_rec[0] = coef[PEV_T_REC_CA] * ids;
  cf_rec[1] = coef[PEV_T_REC_N] * ids;
  cf_rec[2] = coef[PEV_T_REC_C] * ids;
  const int64_t bi = (int64_t)b * A.L + i;
  if (gN || gCA || gC) {
    v3 n, ca, c;
    residue_bwd(A, cf, cf_rec, b, i, n, ca, c);
    if (gN) st3(gN + 3 * bi, n);
    if (gCA) st3(gCA + 3 * bi, ca);
    if (gC) st3(gC + 3 * bi, c);
  }
  if (glog && A.logits) ce_row_bwd(A, cf[PEV_T_SEQ], bi, glog + bi * A.C);
}

__global__ void __launch_bounds__(256)
loss_kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                   const float* __restrict__ row_weight, int64_t rows, int D, const float* __restrict__ coef,
                   const float* __restrict__ inv_den, int term, float* __restrict__ gmu,
                   float* __restrict__ glv) {
  const float cf = coef[term] * inv_den[term];
  const int64_t n = rows * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const float w = row_weight ? row_weight[idx / D] : 1.0f;
    kl_elem_bwd(mu[idx], lv[idx], cf * w, gmu + idx, glv + idx);
  }
}

// ------------------------------------------------------------------------------------------ dihedral API
__global__ void dihedrals_fwd_kernel(pev_loss_args A, const float* N, const float* CA, const float* C,
                                     float* __restrict__ out) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.L) return;
  ResDih r;
  residue_dihedrals(A, N, CA, C, b, i, r);
  float* o = out + ((int64_t)b * A.L + i) * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = r.slot[k];
}

__global__ void dihedrals_bwd_kernel(pev_loss_args A, const float* __restrict__ gout, float* gN, float* gCA,
                                     float* gC) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.L) return;
  v3 n, ca, c;
  dihedrals_bwd_residue(A, gout, b, i, n, ca, c);
  const int64_t bi = (int64_t)b * A.L + i;
  st3(gN + 3 * bi, n); st3(gCA + 3 * bi, ca); st3(gC + 3 * bi, c);
}

__global__ void __launch_bounds__(256)
dihedral_terms_fwd_kernel(const float* __restrict__ dih, const float* __restrict__ tgt,
                          const float* __restrict__ mask, int64_t n, double* __restrict__ sums) {
  __shared__ float red[32];
  float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t bi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; bi < n; bi += (int64_t)gridDim.x * blockDim.x) {
    dihterm_fwd(dih + 6 * bi, tgt ? tgt + 6 * bi : nullptr, mask[bi], &v[0], &v[1], &v[2], &v[3]);
    v[4] += mask[bi];
  }
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    float t = block_sum(v[k], red);
    if (threadIdx.x == 0 && t != 0.f) atomicAdd(sums + k, (double)t);
  }
}

__global__ void __launch_bounds__(256)
dihedral_terms_bwd_kernel(const float* __restrict__ dih, const float* __restrict__ tgt,
                          const float* __restrict__ mask, const float* __restrict__ coef3, int64_t n,
                          float* __restrict__ gdih) {
  float cf3[3] = {coef3[0], coef3[1], coef3[2]};
  for (int64_t bi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; bi < n; bi += (int64_t)gridDim.x * blockDim.x)
    dihterm_bwd(dih + 6 * bi, tgt ? tgt + 6 * bi : nullptr, mask[bi], cf3, gdih + 6 * bi);
}

static int grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes, const char* what) {
  if (bytes > (size_t)kMaxDynSmem) return set_error(1, "%s: conformer too long for the shared-memory tile", what);
  if (bytes > 40 * 1024) {      // static shared memory counts against the 48 KB default as well
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return set_error(2, "%s: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

}  // namespace pev

using namespace pev;

extern "C" {

int pev_loss_fwd(const pev_loss_args* ap, double* acc_global, double* acc_sample, void* stream) {
  PEV_REQUIRE(ap && acc_global && acc_sample, "null argument");
  const pev_loss_args A = *ap;
  PEV_REQUIRE(A.mask && A.B > 0 && A.L > 0, "mask / shape missing");
  PEV_REQUIRE(!A.enable_geometry || (A.pred_N && A.pred_CA && A.pred_C), "geometry terms need N, CA, C");
  PEV_REQUIRE(!A.enable_clash || (A.pred_N && A.pred_CA && A.pred_C), "clash needs N, CA, C");
  PEV_REQUIRE(A.pair_stride <= 0 || (A.pred_CA && A.target_CA), "pair term needs pred/target CA");
  PEV_REQUIRE(!A.logits || A.labels, "logits without labels");
  cudaStream_t st = as_stream(stream);
  int rc;
  dim3 grid((A.L + kResThreads - 1) / kResThreads, A.B);
  loss_residue_fwd_kernel<<<grid, kResThreads, 0, st>>>(A, acc_global, acc_sample);
  if ((rc = after_launch("loss_residue_fwd_kernel"))) return rc;
  if (A.mu_l && A.lv_l) {
    const int64_t rows = (int64_t)A.B * A.L;
    loss_kl_fwd_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(A.mu_l, A.lv_l, A.mask, rows, A.D,
                                                               acc_global + 2 * PEV_T_KL_L);
    if ((rc = after_launch("loss_kl_fwd_kernel"))) return rc;
  }
  if (A.mu_g && A.lv_g) {
    loss_kl_fwd_kernel<<<grid_for((int64_t)A.B * 32, 256), 256, 0, st>>>(A.mu_g, A.lv_g, nullptr, A.B, A.G,
                                                                       acc_global + 2 * PEV_T_KL_G);
    if ((rc = after_launch("loss_kl_fwd_kernel"))) return rc;
  }
  if (A.pair_stride > 0) {
    const int M = (A.L + A.pair_stride - 1) / A.pair_stride;
    const size_t smem = sizeof(float) * 7 * (size_t)M;
    if ((rc = opt_in_smem(loss_pair_kernel, smem, "pair_distance"))) return rc;
    dim3 g((M + kPairThreads - 1) / kPairThreads, A.B);
    loss_pair_kernel<<<g, kPairThreads, smem, st>>>(A, M, acc_global, nullptr, nullptr, nullptr);
    if ((rc = after_launch("loss_pair_kernel"))) return rc;
  }
  if (A.enable_clash) {
    const size_t smem = sizeof(float) * 12 * (size_t)A.L;
    if ((rc = opt_in_smem(loss_clash_kernel<false>, smem, "clash"))) return rc;
    dim3 g((3 * A.L + kClashTile - 1) / kClashTile, A.B);
    loss_clash_kernel<false><<<g, kClashThreads, smem, st>>>(A, acc_sample, nullptr, nullptr, nullptr, nullptr, nullptr);
    if ((rc = after_launch("loss_clash_kernel"))) return rc;
  }
  return 0;
}

int pev_loss_finalize(const double* acc_global, const double* acc_sample, int32_t B, float* terms,
                      float* inv_den, void* stream) {
  PEV_REQUIRE(acc_global && acc_sample && terms && inv_den && B > 0, "null argument");
  loss_finalize_kernel<<<1, 256, 0, as_stream(stream)>>>(acc_global, acc_sample, B, terms, inv_den);
  return after_launch("loss_finalize_kernel");
}

int pev_loss_bwd(const pev_loss_args* ap, const float* coef, const float* inv_den, float* gN, float* gCA,
                 float* gC, float* glog, float* gmul, float* glvl, float* gmug, float* glvg, void* stream) {
  PEV_REQUIRE(ap && coef && inv_den, "null argument");
  const pev_loss_args A = *ap;
  cudaStream_t st = as_stream(stream);
  int rc;
  if (gN || gCA || gC || glog) {
    dim3 grid((A.L + kResThreads - 1) / kResThreads, A.B);
    loss_residue_bwd_kernel<<<grid, kResThreads, 0, st>>>(A, coef, inv_den, gN, gCA, gC, glog);
    if ((rc = after_launch("loss_residue_bwd_kernel"))) return rc;
  }
  if (A.pair_stride > 0 && gCA) {
    const int M = (A.L + A.pair_stride - 1) / A.pair_stride;
    const size_t smem = sizeof(float) * 7 * (size_t)M;
    if ((rc = opt_in_smem(loss_pair_kernel, smem, "pair_distance"))) return rc;
    dim3 g((M + kPairThreads - 1) / kPairThreads, A.B);
    loss_pair_kernel<<<g, kPairThreads, smem, st>>>(A, M, nullptr, coef, inv_den, gCA);
    if ((rc = after_launch("loss_pair_kernel"))) return rc;
  }
  if (A.enable_clash && (gN || gCA || gC)) {
    const size_t smem = sizeof(float) * 12 * (size_t)A.L;
    if ((rc = opt_in_smem(loss_clash_kernel<true>, smem, "clash"))) return rc;
    dim3 g((3 * A.L + kClashTile - 1) / kClashTile, A.B);
    loss_clash_kernel<true><<<g, kClashThreads, smem, st>>>(A, nullptr, coef, inv_den, gN, gCA, gC);
    if ((rc = after_launch("loss_clash_kernel"))) return rc;
  }
  if (A.mu_l && gmul && glvl) {
    const int64_t rows = (int64_t)A.B * A.L;
    loss_kl_bwd_kernel<<<grid_for(rows * A.D, 256), 256, 0, st>>>(A.mu_l, A.lv_l, A.mask, rows, A.D, coef,
                                                                inv_den, PEV_T_KL_L, gmul, glvl);
    if ((rc = after_launch("loss_kl_bwd_kernel"))) return rc;
  }
  if (A.mu_g && gmug && glvg) {
    loss_kl_bwd_kernel<<<grid_for((int64_t)A.B * A.G, 256), 256, 0, st>>>(A.mu_g, A.lv_g, nullptr, A.B, A.G,
                                                                        coef, inv_den, PEV_T_KL_G, gmug, glvg);
    if ((rc = after_launch("loss_kl_bwd_kernel"))) return rc;
  }
  return 0;
}

static pev_loss_args dihedral_args(const float* N, const float* CA, const float* C, const float* mask,
                                   int32_t B, int32_t L) {
  pev_loss_args A = {};
  A.pred_N = N; A.pred_CA = CA; A.pred_C = C; A.mask = mask; A.B = B; A.L = L;
  return A;
}

int pev_dihedrals_fwd(const float* N, const float* CA, const float* C, const float* mask, int32_t B,
                      int32_t L, float* out, void* stream) {
  PEV_REQUIRE(N && CA && C && mask && out && B > 0 && L > 0, "null argument");
  dim3 grid((L + 127) / 128, B);
  dihedrals_fwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(dihedral_args(N, CA, C, mask, B, L), N, CA, C, out);
  return after_launch("dihedrals_fwd_kernel");
}

int pev_dihedrals_bwd(const float* N, const float* CA, const float* C, const float* mask,
                      const float* gout, int32_t B, int32_t L, float* gN, float* gCA, float* gC, void* stream) {
  PEV_REQUIRE(N && CA && C && mask && gout && gN && gCA && gC && B > 0 && L > 0, "null argument");
  dim3 grid((L + 127) / 128, B);
  dihedrals_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(dihedral_args(N, CA, C, mask, B, L), gout, gN, gCA, gC);
  return after_launch("dihedrals_bwd_kernel");
}

int pev_dihedral_terms_fwd(const float* dih, const float* target, const float* mask, int32_t B, int32_t L,
                           double* sums, void* stream) {
  PEV_REQUIRE(dih && mask && sums && B > 0 && L > 0, "null argument");
  const int64_t n = (int64_t)B * L;
  dihedral_terms_fwd_kernel<<<grid_for(n, 256, 2), 256, 0, as_stream(stream)>>>(dih, target, mask, n, sums);
  return after_launch("dihedral_terms_fwd_kernel");
}

int pev_dihedral_terms_bwd(const float* dih, const float* target, const float* mask, const float* coef3,
                           int32_t B, int32_t L, float* gdih, void* stream) {
  PEV_REQUIRE(dih && mask && coef3 && gdih && B > 0 && L > 0, "null argument");
  const int64_t n = (int64_t)B * L;
  dihedral_terms_bwd_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(dih, target, mask, coef3, n, gdih);
  return after_launch("dihedral_terms_bwd_kernel");
}

}  // extern "C"
