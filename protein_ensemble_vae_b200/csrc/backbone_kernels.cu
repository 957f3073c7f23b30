// Backbone placement of EGNNDecoder.forward (models/en_gnn_decoder.py:260-310): N / C from the direction heads and the
// 3-step peptide pull, one thread per residue of the packed batch, forward and backward in one launch each (the eager
// form was ~20 element-wise launches per direction).  Arithmetic in pev_backbone_body.cuh (shared with tests/hostcheck).
#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "pev_backbone_body.cuh"

namespace pev {

__global__ void __launch_bounds__(256)
backbone_fwd_kernel(const float* __restrict__ n_dir, int ldn, const float* __restrict__ c_dir, int ldc,
                    const float* __restrict__ x_ca, const uint8_t* __restrict__ starts, int64_t N, float* __restrict__ x_n,
                    float* __restrict__ x_c) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < N) backbone_fwd_residue(n_dir, ldn, c_dir, ldc, x_ca, starts, N, k, x_n, x_c);
}

__global__ void __launch_bounds__(256)
backbone_bwd_kernel(const float* __restrict__ n_dir, int ldn, const float* __restrict__ c_dir, int ldc,
                    const float* __restrict__ x_ca, const uint8_t* __restrict__ starts, int64_t N, const float* __restrict__ g_xn,
                    const float* __restrict__ g_xc, float* __restrict__ g_ndir, float* __restrict__ g_cdir,
                    float* __restrict__ g_xca) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < N) backbone_bwd_residue(n_dir, ldn, c_dir, ldc, x_ca, starts, N, k, g_xn, g_xc, g_ndir, g_cdir, g_xca);
}

}  // namespace pev

using namespace pev;

extern "C" int pev_backbone_fwd(const float* n_dir, int32_t ldn, const float* c_dir, int32_t ldc, const float* x_ca,
                                const uint8_t* starts, int64_t N, float* x_n, float* x_c, void* stream) {
  PEV_REQUIRE(N >= 0 && ldn >= 3 && ldc >= 3, "bad argument");
  if (N == 0) return 0;
  PEV_REQUIRE(n_dir && c_dir && x_ca && starts && x_n && x_c, "null argument");
  backbone_fwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, as_stream(stream)>>>(n_dir, ldn, c_dir, ldc, x_ca, starts, N, x_n, x_c);
  return after_launch("backbone_fwd_kernel");
}

extern "C" int pev_backbone_bwd(const float* n_dir, int32_t ldn, const float* c_dir, int32_t ldc, const float* x_ca,
                                const uint8_t* starts, int64_t N, const float* g_xn, const float* g_xc, float* g_ndir,
                                float* g_cdir, float* g_xca, void* stream) {
  PEV_REQUIRE(N >= 0 && ldn >= 3 && ldc >= 3, "bad argument");
  if (N == 0) return 0;
  PEV_REQUIRE(n_dir && c_dir && x_ca && starts && g_xn && g_xc && g_ndir && g_cdir && g_xca, "null argument");
  backbone_bwd_kernel<<<(unsigned)((N + 255) / 256), 256, 0, as_stream(stream)>>>(n_dir, ldn, c_dir, ldc, x_ca, starts, N, g_xn,
                                                                                  g_xc, g_ndir, g_cdir, g_xca);
  return after_launch("backbone_bwd_kernel");
}
