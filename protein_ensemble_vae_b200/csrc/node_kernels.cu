// Node-level kernels of the EGNN layer: y = LayerNorm(x + res) (models/en_gnn_decoder.py:72-73, `h = norm_h(h + h_update)`)
// and its backward, one warp per row, fp32.  HBM-bound streams: forward reads x, res and writes r = x + res, y
// (+ mean, rstd); backward reads gy, r and writes gr in ONE pass, with the column sums d gamma = sum gy xhat and
// d beta = sum gy accumulated in registers across the rows a warp visits (grid-stride), then block -> global atomics.
#include "../../include/pev_b200.h"
#include "pev_common.cuh"

namespace pev {

template <int D>   // D = 128, 256 or 512: D / 128 float4 per lane
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ gamma,
                         const float* __restrict__ beta, float eps, int64_t N, float* __restrict__ r_out,
                         float* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd) {
  constexpr int V = D / 128;
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g4[V], b4[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    g4[k] = reinterpret_cast<const float4*>(gamma)[lane + 32 * k];
    b4[k] = reinterpret_cast<const float4*>(beta)[lane + 32 * k];
  }
  for (int64_t row = warp0; row < N; row += nwarps) {
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      v[k] = reinterpret_cast<const float4*>(x + row * D)[lane + 32 * k];
      if (res) {
        const float4 q = reinterpret_cast<const float4*>(res + row * D)[lane + 32 * k];
        v[k].x += q.x; v[k].y += q.y; v[k].z += q.z; v[k].w += q.w;
      }
      s += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float mu = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float a = v[k].x - mu, b = v[k].y - mu, c = v[k].z - mu, d = v[k].w - mu;
      ss += (a * a + b * b) + (c * c + d * d);
    }
    const float rs = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      if (r_out) reinterpret_cast<float4*>(r_out + row * D)[lane + 32 * k] = v[k];
      float4 o;
      o.x = (v[k].x - mu) * rs * g4[k].x + b4[k].x;
      o.y = (v[k].y - mu) * rs * g4[k].y + b4[k].y;
      o.z = (v[k].z - mu) * rs * g4[k].z + b4[k].z;
      o.w = (v[k].w - mu) * rs * g4[k].w + b4[k].w;
      reinterpret_cast<float4*>(y + row * D)[lane + 32 * k] = o;
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

template <int D>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ r, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, int64_t N,
                     float* __restrict__ gr, float* __restrict__ partial /*[gridDim.x][2][D]*/) {
  constexpr int V = D / 128;
  __shared__ float sG[8][D], sB[8][D];                  // one slot per warp: summed in a fixed order, no atomics
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 g4[V], dg[V], db[V];
#pragma unroll
  for (int k = 0; k < V; ++k) {
    g4[k] = reinterpret_cast<const float4*>(gamma)[lane + 32 * k];
    dg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int64_t row = warp0; row < N; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float4 xh[V], gg[V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const float4 g = reinterpret_cast<const float4*>(gy + row * D)[lane + 32 * k];
      const float4 v = reinterpret_cast<const float4*>(r + row * D)[lane + 32 * k];
      xh[k] = make_float4((v.x - mu) * rs, (v.y - mu) * rs, (v.z - mu) * rs, (v.w - mu) * rs);
      db[k].x += g.x; db[k].y += g.y; db[k].z += g.z; db[k].w += g.w;
      dg[k].x = fmaf(g.x, xh[k].x, dg[k].x); dg[k].y = fmaf(g.y, xh[k].y, dg[k].y);
      dg[k].z = fmaf(g.z, xh[k].z, dg[k].z); dg[k].w = fmaf(g.w, xh[k].w, dg[k].w);
      gg[k] = make_float4(g.x * g4[k].x, g.y * g4[k].y, g.z * g4[k].z, g.w * g4[k].w);
      s1 += (gg[k].x + gg[k].y) + (gg[k].z + gg[k].w);
      s2 += (gg[k].x * xh[k].x + gg[k].y * xh[k].y) + (gg[k].z * xh[k].z + gg[k].w * xh[k].w);
    }
    const float m1 = warp_sum(s1) * (1.0f / D), m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      float4 o;
      o.x = rs * (gg[k].x - m1 - xh[k].x * m2);
      o.y = rs * (gg[k].y - m1 - xh[k].y * m2);
      o.z = rs * (gg[k].z - m1 - xh[k].z * m2);
      o.w = rs * (gg[k].w - m1 - xh[k].w * m2);
      reinterpret_cast<float4*>(gr + row * D)[lane + 32 * k] = o;
    }
  }
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const int c = 4 * (lane + 32 * k), w = threadIdx.x >> 5;
    *reinterpret_cast<float4*>(&sG[w][c]) = dg[k];
    *reinterpret_cast<float4*>(&sB[w][c]) = db[k];
  }
  __syncthreads();
  float* dst = partial + (int64_t)blockIdx.x * 2 * D;
  for (int k = threadIdx.x; k < D; k += blockDim.x) {
    dst[k] = ((sG[0][k] + sG[1][k]) + (sG[2][k] + sG[3][k])) + ((sG[4][k] + sG[5][k]) + (sG[6][k] + sG[7][k]));
    dst[D + k] = ((sB[0][k] + sB[1][k]) + (sB[2][k] + sB[3][k])) + ((sB[4][k] + sB[5][k]) + (sB[6][k] + sB[7][k]));
  }
}

// out[c] = sum_rows g[row, c]  (bias gradients): block of 256 threads owns 64 columns x 4 row lanes
__global__ void __launch_bounds__(256)
column_sum_kernel(const float* __restrict__ g, int64_t N, int D, float* __restrict__ partial /*[gridDim.y][D]*/) {
  __shared__ float4 sm[256];
  const int c4 = threadIdx.x & 15, rl = threadIdx.x >> 4;          // 16 float4 columns x 16 row lanes
  const int col4 = blockIdx.x * 16 + c4;                            // float4 column index
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col4 * 4 < D) {
    for (int64_t row = (int64_t)blockIdx.y * 16 + rl; row < N; row += (int64_t)gridDim.y * 16) {
      const float4 v = reinterpret_cast<const float4*>(g + row * D)[col4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sm[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0 && col4 * 4 < D) {
    for (int k = 1; k < 16; ++k) {
      const float4 v = sm[k * 16 + c4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * D)[col4] = acc;
  }
}

static int rows_grid(int64_t N) {
  int64_t g = (N + 7) / 8;                      // 8 warps per block
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// Backward of the dropout / ReLU epilogues of pev_linear: out = keep(seed, element) && (y == null || y != 0) ? g * scale : 0
// (the dropout mask is the counter hash the forward epilogue used; ReLU zeros are read off the stored output y)
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__global__ void __launch_bounds__(256)
mask_grad_kernel(const float* __restrict__ g, const float* __restrict__ y, int64_t n4, float scale, uint32_t thresh, uint32_t seed,
                 float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(g)[i];
    float4 yy = y ? reinterpret_cast<const float4*>(y)[i] : make_float4(1.f, 1.f, 1.f, 1.f);
    float r[4] = {v.x, v.y, v.z, v.w};
    const float ym[4] = {yy.x, yy.y, yy.z, yy.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool keep = thresh == 0u || mix32(seed + (uint32_t)(4 * i + k) * 0x9e3779b9U) >= thresh;   // index mod 2^32
      r[k] = (keep && ym[k] != 0.f) ? r[k] * scale : 0.f;
    }
    reinterpret_cast<float4*>(out)[i] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// out[c, r] = scale * in[r, c]: 32 x 32 tiles through shared memory, coalesced both ways (the weight of a data-gradient
// GEMM; a strided element-wise copy of a 1536 x 512 weight took 40 us, this takes 4)
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ in, int rows, int cols, float scale, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8)
    if (r0 + k < rows && c0 + tx < cols) tile[k][tx] = in[(int64_t)(r0 + k) * cols + c0 + tx];
  __syncthreads();
  for (int k = ty; k < 32; k += 8)
    if (c0 + k < cols && r0 + tx < rows) out[(int64_t)(c0 + k) * rows + r0 + tx] = scale * tile[tx][k];
}

}  // namespace pev

using namespace pev;

extern "C" {

int pev_transpose(const float* in, int32_t rows, int32_t cols, float scale, float* out, void* stream) {
  PEV_REQUIRE(in && out && rows > 0 && cols > 0, "bad argument");
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), 256, 0, as_stream(stream)>>>(in, rows, cols, scale, out);
  return after_launch("transpose_kernel");
}

int pev_mask_grad(const float* g, const float* y, int64_t n, float p_drop, uint32_t seed, float* out, void* stream) {
  PEV_REQUIRE(g && out && n >= 0 && n % 4 == 0 && p_drop >= 0.f && p_drop < 1.f, "bad argument (n must be a multiple of 4)");
  if (n == 0) return 0;
  const uint32_t thresh = p_drop > 0.f ? (uint32_t)(p_drop * 4294967296.0) : 0u;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  mask_grad_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(g, y, n / 4, 1.0f / (1.0f - p_drop), thresh, seed, out);
  return after_launch("mask_grad_kernel");
}

int pev_add_layernorm_fwd(const float* x, const float* res, const float* gamma, const float* beta, float eps,
                          int64_t N, int32_t D, float* r_out, float* y, float* mean, float* rstd, void* stream) {
  PEV_REQUIRE(N >= 0 && (D == 128 || D == 256 || D == 512), "D must be 128, 256 or 512");
  if (N == 0) return 0;
  PEV_REQUIRE(x && gamma && beta && y && mean && rstd, "null argument");
  cudaStream_t st = as_stream(stream);
  if (D == 128)
    add_layernorm_fwd_kernel<128><<<rows_grid(N), 256, 0, st>>>(x, res, gamma, beta, eps, N, r_out, y, mean, rstd);
  else if (D == 256)
    add_layernorm_fwd_kernel<256><<<rows_grid(N), 256, 0, st>>>(x, res, gamma, beta, eps, N, r_out, y, mean, rstd);
  else
    add_layernorm_fwd_kernel<512><<<rows_grid(N), 256, 0, st>>>(x, res, gamma, beta, eps, N, r_out, y, mean, rstd);
  return after_launch("add_layernorm_fwd_kernel");
}

int64_t pev_node_workspace_bytes(void) { return (int64_t)sm_count() * 8 * 2 * 512 * (int64_t)sizeof(float); }

int pev_layernorm_bwd(const float* gy, const float* r, const float* gamma, const float* mean, const float* rstd,
                      int64_t N, int32_t D, float* workspace, float* gr, float* dgamma, float* dbeta, void* stream) {
  PEV_REQUIRE(N >= 0 && (D == 128 || D == 256 || D == 512) && dgamma && dbeta, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    cudaMemsetAsync(dgamma, 0, sizeof(float) * D, st);
    cudaMemsetAsync(dbeta, 0, sizeof(float) * D, st);
    return 0;
  }
  PEV_REQUIRE(gy && r && gamma && mean && rstd && gr && workspace, "null argument");
  const int grid = rows_grid(N);
  if (D == 128)
    layernorm_bwd_kernel<128><<<grid, 256, 0, st>>>(gy, r, gamma, mean, rstd, N, gr, workspace);
  else if (D == 256)
    layernorm_bwd_kernel<256><<<grid, 256, 0, st>>>(gy, r, gamma, mean, rstd, N, gr, workspace);
  else
    layernorm_bwd_kernel<512><<<grid, 256, 0, st>>>(gy, r, gamma, mean, rstd, N, gr, workspace);
  if (int rc = after_launch("layernorm_bwd_kernel")) return rc;
  // dgamma | dbeta are adjacent in a partial row: one launch over 2 D columns when the outputs are adjacent too
  if (dbeta == dgamma + D) return launch_partial_reduce(workspace, grid, 2 * D, 2 * D, 1.0f, dgamma, st);
  if (int rc = launch_partial_reduce(workspace, grid, 2 * D, D, 1.0f, dgamma, st)) return rc;
  return launch_partial_reduce(workspace + D, grid, 2 * D, D, 1.0f, dbeta, st);
}

int pev_column_sum(const float* g, int64_t N, int32_t D, float* workspace, float* out, void* stream) {
  PEV_REQUIRE(N >= 0 && D > 0 && D % 4 == 0 && D <= 4096 && out, "bad argument (D must be a multiple of 4, <= 4096)");
  cudaStream_t st = as_stream(stream);
  if (N == 0) {
    cudaMemsetAsync(out, 0, sizeof(float) * D, st);
    return 0;
  }
  PEV_REQUIRE(g && workspace, "null argument");
  const int gx = (D / 4 + 15) / 16;
  int64_t gy = (N + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8 / gx + 1;
  if (gy > cap) gy = cap;
  column_sum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(g, N, D, workspace);
  if (int rc = after_launch("column_sum_kernel")) return rc;
  return launch_partial_reduce(workspace, (int)gy, D, D, 1.0f, out, st);
}

}  // extern "C"
