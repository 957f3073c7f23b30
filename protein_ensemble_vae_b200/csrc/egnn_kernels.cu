// Graph builder, fp32 edge prologue (K1 fp32 form) and exact-order scatter / coordinate update (K2).
// Per-thread arithmetic: pev_egnn_body.cuh (host/device, checked by tests/hostcheck).
// Reference: models/en_gnn_decoder.py:53-87, :174-198.
#include <cuda_bf16.h>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_egnn_body.cuh"

namespace pev {

__global__ void band_graph_kernel(const int32_t* __restrict__ cu, const int64_t* __restrict__ edge_base, int B,
                                  int W, int64_t N, int32_t* row_ptr, int32_t* row, int32_t* col,
                                  int32_t* csc_perm, float* dinv) {
  const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= N) return;
  band_node_build(cu, edge_base, B, W, node, row_ptr, row, col, csc_perm, dinv, node == N - 1);
}

// one thread per (edge, feature); consecutive threads = consecutive features -> coalesced
__global__ void __launch_bounds__(256)
edge_prologue_fwd_kernel(const float* __restrict__ AB, const float* __restrict__ x,
                         const float* __restrict__ wd, const float* __restrict__ b1,
                         const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t E, int H,
                         float* __restrict__ u) {
  const int64_t total = E * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / H;
    const int k = (int)(idx - e * H);
    const int r = row[e], c = col[e];
    u[idx] = edge_prologue_elem(AB, wd, b1, H, r, c, edge_d2(x, r, c), k);
  }
}

// gd2[e] = wd . gu[e,:]   (one warp per edge)
__global__ void __launch_bounds__(256)
edge_gd2_kernel(const float* __restrict__ gu, const float* __restrict__ wd, int64_t E, int H,
                float* __restrict__ gd2) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t e = warp; e < E; e += nwarps) {
    float s = 0.f;
    for (int k = lane; k < H; k += 32) s += wd[k] * gu[e * H + k];
    s = warp_sum(s);
    if (lane == 0) gd2[e] = s;
  }
}

__global__ void __launch_bounds__(256)
edge_prologue_bwd_feat_kernel(const float* __restrict__ gu, const float* __restrict__ x,
                              const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                              const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ csc_perm,
                              int64_t N, int H, float* __restrict__ gAB, float* __restrict__ gwd_part) {
  const int64_t total = N * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / H;
    const int k = (int)(idx - i * H);
    edge_prologue_bwd_feat(gu, x, row_ptr, col, col_ptr, csc_perm, H, i, k, &gAB[i * 2 * H + k],
                           &gAB[i * 2 * H + H + k], &gwd_part[idx]);
  }
}

__global__ void __launch_bounds__(128)
edge_prologue_bwd_coord_kernel(const float* __restrict__ gd2, const float* __restrict__ x,
                               const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ row,
                               const int32_t* __restrict__ col, const int32_t* __restrict__ col_ptr,
                               const int32_t* __restrict__ csc_perm, int64_t N, float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  st3(gx + 3 * i, edge_prologue_bwd_coord(gd2, x, row_ptr, row, col, col_ptr, csc_perm, i));
}

// K2 forward: thread per (node, feature) walks the node's contiguous edge segment in order
__global__ void __launch_bounds__(256)
scatter_feature_kernel(const float* __restrict__ m, const int32_t* __restrict__ row_ptr, int64_t N, int H,
                       float* __restrict__ agg) {
  const int64_t total = N * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / H;
    agg[idx] = scatter_feature(m, row_ptr, H, i, (int)(idx - i * H));
  }
}

__global__ void __launch_bounds__(128)
coord_update_kernel(const float* __restrict__ w, const float* __restrict__ x, const float* __restrict__ dinv,
                    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, int64_t N,
                    float* __restrict__ x_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  st3(x_out + 3 * i, coord_update(w, x, dinv, row_ptr, col, i));
}

__global__ void __launch_bounds__(256)
scatter_bwd_gm_kernel(const float* __restrict__ gagg, const int32_t* __restrict__ row, int64_t E, int H,
                      float* __restrict__ gm) {
  const int64_t total = E * H;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = idx / H;
    gm[idx] = gagg[(int64_t)row[e] * H + (idx - e * H)];
  }
}

__global__ void __launch_bounds__(256)
scatter_bwd_gw_kernel(const float* __restrict__ gxo, const float* __restrict__ x, const float* __restrict__ dinv,
                      const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t E,
                      float* __restrict__ gw) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
    gw[e] = coord_update_bwd_w(gxo, x, dinv, row[e], col[e]);
}

__global__ void __launch_bounds__(128)
scatter_bwd_gx_kernel(const float* __restrict__ gxo, const float* __restrict__ w, const float* __restrict__ dinv,
                      const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ row,
                      const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ csc_perm, int64_t N,
                      float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  st3(gx + 3 * i, coord_update_bwd_x(gxo, w, dinv, row_ptr, row, col_ptr, csc_perm, i));
}

__global__ void __launch_bounds__(128)
edge_prologue_bwd_coord_accum_kernel(const float* __restrict__ gd2, const float* __restrict__ x,
                                     const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ row,
                                     const int32_t* __restrict__ col, const int32_t* __restrict__ col_ptr,
                                     const int32_t* __restrict__ csc_perm, int64_t N, float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const v3 g = edge_prologue_bwd_coord(gd2, x, row_ptr, row, col, col_ptr, csc_perm, i);
  gx[3 * i] += g.x; gx[3 * i + 1] += g.y; gx[3 * i + 2] += g.z;
}

static int grid_cap(int64_t n, int threads, int per_sm = 16) {
  int64_t g = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace pev

using namespace pev;

extern "C" {

int pev_band_graph_build(const int32_t* cu_seqlens, const int64_t* edge_base, int32_t B, int32_t W, int64_t N,
                         int32_t* row_ptr, int32_t* row, int32_t* col, int32_t* csc_perm, float* dinv,
                         void* stream) {
  PEV_REQUIRE(cu_seqlens && edge_base && B > 0 && W >= 0 && N >= 0, "bad argument");
  if (N == 0) {
    if (row_ptr) cudaMemsetAsync(row_ptr, 0, sizeof(int32_t), as_stream(stream));
    return 0;
  }
  band_graph_kernel<<<(unsigned)((N + 127) / 128), 128, 0, as_stream(stream)>>>(cu_seqlens, edge_base, B, W, N,
                                                                               row_ptr, row, col, csc_perm, dinv);
  return after_launch("band_graph_kernel");
}

int pev_edge_prologue_fwd(const float* AB, const float* x, const float* wd, const float* b1, const int32_t* row,
                          const int32_t* col, int64_t E, int32_t H, float* u, void* stream) {
  PEV_REQUIRE(AB && x && wd && b1 && u && H > 0 && E >= 0, "bad argument");
  if (E == 0) return 0;
  PEV_REQUIRE(row && col, "edge list missing");
  edge_prologue_fwd_kernel<<<grid_cap(E * H, 256), 256, 0, as_stream(stream)>>>(AB, x, wd, b1, row, col, E, H, u);
  return after_launch("edge_prologue_fwd_kernel");
}

int pev_edge_prologue_bwd(const float* gu, const float* x, const float* wd, const int32_t* row_ptr,
                          const int32_t* row, const int32_t* col, const int32_t* col_ptr, const int32_t* csc_perm,
                          int64_t N, int64_t E, int32_t H, float* gAB, float* gx, float* gwd_part,
                          float* scratch_gd2, void* stream) {
  PEV_REQUIRE(x && wd && row_ptr && col_ptr && gAB && gx && gwd_part && H > 0, "bad argument");
  if (N == 0) return 0;
  PEV_REQUIRE(E == 0 || (gu && row && col && csc_perm && scratch_gd2), "edge arrays missing");
  cudaStream_t st = as_stream(stream);
  int rc;
  if (E > 0) {
    edge_gd2_kernel<<<grid_cap(E * 32, 256), 256, 0, st>>>(gu, wd, E, H, scratch_gd2);
    if ((rc = after_launch("edge_gd2_kernel"))) return rc;
  }
  edge_prologue_bwd_feat_kernel<<<grid_cap(N * H, 256), 256, 0, st>>>(gu, x, row_ptr, col, col_ptr, csc_perm, N,
                                                                     H, gAB, gwd_part);
  if ((rc = after_launch("edge_prologue_bwd_feat_kernel"))) return rc;
  edge_prologue_bwd_coord_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(scratch_gd2, x, row_ptr, row, col,
                                                                             col_ptr, csc_perm, N, gx);
  return after_launch("edge_prologue_bwd_coord_kernel");
}

int pev_edge_coord_bwd_accum(const float* gd2, const float* x, const int32_t* row_ptr, const int32_t* row,
                             const int32_t* col, const int32_t* col_ptr, const int32_t* csc_perm, int64_t N, int64_t E,
                             float* gx_accum, void* stream) {
  PEV_REQUIRE(x && row_ptr && col_ptr && gx_accum && N >= 0, "bad argument");
  if (N == 0 || E == 0) return 0;
  PEV_REQUIRE(gd2 && row && col && csc_perm, "edge arrays missing");
  edge_prologue_bwd_coord_accum_kernel<<<(unsigned)((N + 127) / 128), 128, 0, as_stream(stream)>>>(
      gd2, x, row_ptr, row, col, col_ptr, csc_perm, N, gx_accum);
  return after_launch("edge_prologue_bwd_coord_accum_kernel");
}

int pev_scatter_coord_fwd(const float* m, const float* w, const float* x, const float* dinv,
                          const int32_t* row_ptr, const int32_t* col, int64_t N, int32_t H, float* agg,
                          float* x_out, void* stream) {
  PEV_REQUIRE(row_ptr && N >= 0 && H > 0, "bad argument");
  if (N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  int rc;
  if (agg) {
    scatter_feature_kernel<<<grid_cap(N * H, 256), 256, 0, st>>>(m, row_ptr, N, H, agg);
    if ((rc = after_launch("scatter_feature_kernel"))) return rc;
  }
  if (x_out) {
    PEV_REQUIRE(x, "x missing");
    coord_update_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(w, x, dinv, row_ptr, col, N, x_out);
    if ((rc = after_launch("coord_update_kernel"))) return rc;
  }
  return 0;
}

int pev_scatter_coord_bwd(const float* gagg, const float* gxo, const float* w, const float* x, const float* dinv,
                          const int32_t* row_ptr, const int32_t* row, const int32_t* col, const int32_t* col_ptr,
                          const int32_t* csc_perm, int64_t N, int64_t E, int32_t H, float* gm, float* gw,
                          float* gx, void* stream) {
  PEV_REQUIRE(row_ptr && col_ptr && N >= 0 && H > 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  int rc;
  if (gm && E > 0) {
    scatter_bwd_gm_kernel<<<grid_cap(E * H, 256), 256, 0, st>>>(gagg, row, E, H, gm);
    if ((rc = after_launch("scatter_bwd_gm_kernel"))) return rc;
  }
  if (gw && E > 0) {
    scatter_bwd_gw_kernel<<<grid_cap(E, 256), 256, 0, st>>>(gxo, x, dinv, row, col, E, gw);
    if ((rc = after_launch("scatter_bwd_gw_kernel"))) return rc;
  }
  if (gx && N > 0) {
    scatter_bwd_gx_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(gxo, w, dinv, row_ptr, row, col_ptr,
                                                                      csc_perm, N, gx);
    if ((rc = after_launch("scatter_bwd_gx_kernel"))) return rc;
  }
  return 0;
}

}  // extern "C"
