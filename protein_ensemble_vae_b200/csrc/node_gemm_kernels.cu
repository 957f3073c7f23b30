// Node-level linears of the bf16 path on tcgen05 (kind::tf32): the per-node GEMMs of one EGNN layer
// (models/en_gnn_decoder.py:65-73 -- the h_i / h_j halves of phi_e[0], phi_h[0], phi_h[2]) with what used to be separate
// element-wise kernels fused into the epilogue:
//
//   ABH     ABh = 0.5 (h [Wa|Wb]^T + [b1|0])  -> fp16 [N,512]   (the half-domain node projection the edge kernels gather;
//                                                                replaces cuBLAS addmm + a float16 copy kernel)
//   SILU    p = [h | agg] W3^T + b3 ; q = silu(p)               (two A operands: the concatenation is never built)
//   RES_LN  r = h + q W4^T + b4 ; h' = LayerNorm(r)             (row statistics in the epilogue: lane = row)
//   PLAIN   C = A B^T (+ bias) (+ residual)                     (data gradients: g W with a pre-transposed weight)
//   DSILU   C = (A B^T) * silu'(p)                              (backward through phi_h[1])
//
// C[M, Nout] = A[M, K] . W[Nout, K]^T, fp32 operands read by the tensor core as TF32 (10-bit mantissa, fp32 accumulate)
// -- the same arithmetic the round-1 path ran through cuBLAS with allow_tf32.  Tile 128 x 256, K-chunks of 32 floats
// (128-byte SWIZZLE_128B rows) loaded by TMA tensor copies into a 4-stage ring, two TMEM accumulator stages, eight
// epilogue warps (TMEM lane quarter x column half), persistent over the M tiles.  HBM-bound (AI ~ 100 flop/B).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstring>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "tc_common.cuh"

namespace pev {
namespace ng {
using namespace tcx;

constexpr int BM = 128, BN = 256, BK = 32;
constexpr int STAGES = 4;                // ncu: with 3 stages the tensor pipe idled half the time waiting for TMA (1 chunk = 4 MMAs)
constexpr int A_BYTES = BM * BK * 4;     // 16 KB
constexpr int B_BYTES = BN * BK * 4;     // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int VEC_OFF = STAGES * STAGE_BYTES;          // bias[512] | gamma[256] | beta[256]
constexpr int LN_OFF = VEC_OFF + 1024 * 4;             // [2 column halves][128 rows][2] row statistics
constexpr int STG_OFF = LN_OFF + 2 * BM * 2 * 4;        // 8 warps x 2 KB: [16 rows][32 floats] transposition buffers (two passes)
constexpr int BAR_OFF = STG_OFF + 8 * 2048;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
constexpr int EPI_WARPS = 8, TMA_WARP = 8, MMA_WARP = 9;
constexpr int THREADS = 32 * 10;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

enum Epi { EPI_ABH = 0, EPI_SILU = 1, EPI_RES_LN = 2, EPI_PLAIN = 3, EPI_DSILU = 4 };

struct Params {
  int64_t M;
  int Nout;               // 256 or 512
  int k1_chunks;          // K-chunks (32 floats) taken from A1, then k2_chunks from A2
  int k2_chunks;
  const float* bias;      // [Nout] or null
  float scale;            // ABH: 0.5
  // epilogue operands / outputs (row-major, leading dimension = Nout unless stated)
  __half* out16;          // ABH
  float* out;             // SILU: q; RES_LN: y; PLAIN / DSILU: C
  float* out2;            // SILU: p (null in inference); RES_LN: r (null in inference)
  const float* res;       // RES_LN: h [M,256]; PLAIN: residual [M,Nout] or null; DSILU: p [M,256]
  const float* gamma;     // RES_LN
  const float* beta;
  float eps;
  float* mean;            // RES_LN (null in inference)
  float* rstd;
  int64_t ldc;            // PLAIN (and the 3xTF32 kernel): leading dimensions of out / res; 0 = Nout
  int64_t ldres;
  int relu;               // PLAIN: out = max(., 0)
  uint32_t drop_thresh;   // PLAIN: dropout on (C + bias [relu]) BEFORE the residual is added: keep iff hash(seed, row * Nout + col) >= thresh
  uint32_t drop_seed;
  float drop_scale;       // 1 / (1 - p)
};

__device__ __forceinline__ uint32_t mix32(uint32_t x) {      // the counter hash of pev_mask_grad / attn_kernels.cu
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// keep iff hash(seed, element index mod 2^32) >= thresh; `base` = (row * Nout + first column of the block) * golden ratio,
// hoisted out of the per-element loop (32-bit arithmetic: this sits in an issue-bound epilogue)
__device__ __forceinline__ uint32_t drop_base(const Params& p, int64_t row, int col0) {
  return p.drop_seed + ((uint32_t)row * (uint32_t)p.Nout + (uint32_t)col0) * 0x9e3779b9U;
}
__device__ __forceinline__ float drop_elem(const Params& p, float v, uint32_t base, int j) {
  return mix32(base + (uint32_t)j * 0x9e3779b9U) >= p.drop_thresh ? v * p.drop_scale : 0.f;
}

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void half_barrier(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }

// Coalesced store of a warp's [32 rows x 32 floats] register block (lane = row, as tcgen05.ld delivers it): through a 4 KB
// staging buffer (16-byte chunks XOR-swizzled by row: conflict-free both ways) so that one store instruction writes four
// complete 128-byte row segments instead of 32 scattered 16-byte pieces (the per-lane-row form ran at 2 TB/s).
__device__ __forceinline__ void store_block32(float* stg, const float (&v)[32], float* gbase, int64_t ld, int rows_valid,
                                              int lane) {
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float4*>(stg + lane * 32 + ((k ^ (lane & 7)) << 2)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  __syncwarp();
  const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + rr;
    const float4 x = *reinterpret_cast<const float4*>(stg + row * 32 + ((ch ^ (row & 7)) << 2));
    if (row < rows_valid) *reinterpret_cast<float4*>(gbase + row * ld + 4 * ch) = x;
  }
}

// The same through a 2 KB buffer in two passes of 16 rows (frees the shared memory of a fourth operand stage).
__device__ __forceinline__ void store_block32_h(float* stg, const float (&v)[32], float* gbase, int64_t ld, int rows_valid,
                                                int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    __syncwarp();
    if ((lane >> 4) == h) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(stg + (lane & 15) * 32 + ((k ^ (lane & 7)) << 2)) =
            make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
    __syncwarp();
    const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int r16 = it * 4 + rr, row = h * 16 + r16;
      const float4 x = *reinterpret_cast<const float4*>(stg + r16 * 32 + ((ch ^ (r16 & 7)) << 2));
      if (row < rows_valid) *reinterpret_cast<float4*>(gbase + row * ld + 4 * ch) = x;
    }
  }
}

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
node_gemm_kernel(const Params p, const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapA2,
                 const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sBias = reinterpret_cast<float*>(smem + VEC_OFF);
  float* sGamma = sBias + 512;
  float* sBeta = sGamma + 256;
  float* sLn = reinterpret_cast<float*>(smem + LN_OFF);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full = bars;                 // [STAGES]
  uint64_t* empty = bars + STAGES;       // [STAGES]
  uint64_t* tfull = bars + 2 * STAGES;   // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.Nout / BN;
  const int m_tiles = (int)((p.M + BM - 1) / BM);
  const int total = m_tiles * n_tiles;
  const int kchunks = p.k1_chunks + p.k2_chunks;

  for (int k = threadIdx.x; k < 512; k += THREADS) sBias[k] = (p.bias && k < p.Nout) ? p.bias[k] : 0.f;
  const bool wide = p.Nout > 512;                      // PLAIN only: bias read from global memory
  const int64_t ldc = p.ldc ? p.ldc : p.Nout, ldres = p.ldres ? p.ldres : p.Nout;
  if (EPI == EPI_RES_LN)
    for (int k = threadIdx.x; k < 256; k += THREADS) {
      sGamma[k] = p.gamma[k];
      sBeta[k] = p.beta[k];
    }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* dst = smem + stage * STAGE_BYTES;
          if (kc < p.k1_chunks) tma_load_2d(dst, &mapA1, kc * BK, m0, &full[stage]);
          else tma_load_2d(dst, &mapA2, (kc - p.k1_chunks) * BK, m0, &full[stage]);
          tma_load_2d(dst + A_BYTES, &mapB, kc * BK, n0, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_tf32(BM, BN);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks)       // one tf32 MMA covers K = 8 (32 bytes)
            umma_tf32(tmem_base + acc * BN, desc_kmajor(a_base + ks * 32), desc_kmajor(b_base + ks * 32), IDESC,
                      (kc | ks) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: lane = row, registers = columns
    const int q = warp & 3, hh = warp >> 2;            // TMEM lane quarter, 128-column half of the tile
    float* stg = reinterpret_cast<float*>(smem + STG_OFF) + warp * 512;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * BN;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool valid = row < p.M;
      const int64_t wrow0 = (int64_t)m0 + q * 32;                                  // first row of this warp's block
      const int rows_valid = (int)((p.M - wrow0) < 0 ? 0 : ((p.M - wrow0) > 32 ? 32 : (p.M - wrow0)));
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + hh * 128);
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      if (EPI == EPI_RES_LN) {
        // pass 1: row sums of r = acc + b4 + h over this warp's 128 columns; the two column halves meet in shared memory
        float s1 = 0.f, s2 = 0.f;
        const float* hrow = p.res + (valid ? row : 0) * 256 + hh * 128;
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          float hv[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 v = *reinterpret_cast<const float4*>(hrow + 32 * b + 4 * j4);
            hv[4 * j4] = v.x; hv[4 * j4 + 1] = v.y; hv[4 * j4 + 2] = v.z; hv[4 * j4 + 3] = v.w;
          }
          tmem_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float r = __uint_as_float(raw[j]) + sBias[hh * 128 + 32 * b + j] + hv[j];
            s1 += r;
            s2 = fmaf(r, r, s2);
          }
        }
        float* mine = sLn + (hh * BM + q * 32 + lane) * 2;
        const float* other = sLn + ((hh ^ 1) * BM + q * 32 + lane) * 2;
        mine[0] = s1; mine[1] = s2;
        half_barrier(q);
        const float t1 = hh == 0 ? s1 + other[0] : other[0] + s1;       // the same order in both halves
        const float t2 = hh == 0 ? s2 + other[1] : other[1] + s2;
        const float mu = t1 * (1.0f / 256.0f);
        const float var = fmaxf(t2 * (1.0f / 256.0f) - mu * mu, 0.f);
        const float rs = rsqrtf(var + p.eps);
        half_barrier(q);                                               // both halves have read before the next tile writes
        if (valid && hh == 0 && p.mean) { p.mean[row] = mu; p.rstd[row] = rs; }
        // pass 2: normalise and write
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          float hv[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 v = *reinterpret_cast<const float4*>(hrow + 32 * b + 4 * j4);
            hv[4 * j4] = v.x; hv[4 * j4 + 1] = v.y; hv[4 * j4 + 2] = v.z; hv[4 * j4 + 3] = v.w;
          }
          tmem_wait();
          if (b == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          {
            const int c0 = hh * 128 + 32 * b;
            float yv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              hv[j] = __uint_as_float(raw[j]) + sBias[c0 + j] + hv[j];                   // r
              yv[j] = (hv[j] - mu) * rs * sGamma[c0 + j] + sBeta[c0 + j];
            }
            store_block32_h(stg, yv, p.out + wrow0 * 256 + c0, 256, rows_valid, lane);
            if (p.out2) store_block32_h(stg, hv, p.out2 + wrow0 * 256 + c0, 256, rows_valid, lane);
          }
        }
      } else {
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          const int c0 = n0 + hh * 128 + 32 * b;                         // global output column
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          float aux[32];
          if (EPI == EPI_DSILU || (EPI == EPI_PLAIN && p.res)) {
            const float* arow = p.res + (valid ? row : 0) * (EPI == EPI_PLAIN ? ldres : (int64_t)p.Nout) + c0;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 v = *reinterpret_cast<const float4*>(arow + 4 * j4);
              aux[4 * j4] = v.x; aux[4 * j4 + 1] = v.y; aux[4 * j4 + 2] = v.z; aux[4 * j4 + 3] = v.w;
            }
          }
          tmem_wait();
          if (b == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          float val[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            val[j] = __uint_as_float(raw[j]) + ((EPI == EPI_PLAIN && wide) ? (p.bias ? __ldg(p.bias + c0 + j) : 0.f) : sBias[c0 + j]);
          if (EPI == EPI_ABH) {
            if (valid) {
              __half* o = p.out16 + row * p.Nout + c0;
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint32_t h4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const __half2 hp = __floats2half2_rn(p.scale * val[8 * j8 + 2 * u], p.scale * val[8 * j8 + 2 * u + 1]);
                  h4[u] = *reinterpret_cast<const uint32_t*>(&hp);
                }
                *reinterpret_cast<uint4*>(o + 8 * j8) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
              }
            }
          } else {
            if (EPI == EPI_SILU && p.out2) store_block32_h(stg, val, p.out2 + wrow0 * p.Nout + c0, p.Nout, rows_valid, lane);
            const uint32_t dbase = EPI == EPI_PLAIN ? drop_base(p, row, c0) : 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float v = val[j];
              if (EPI == EPI_SILU) val[j] = v / (1.0f + __expf(-v));
              else if (EPI == EPI_DSILU) {
                const float pz = aux[j], sg = 1.0f / (1.0f + __expf(-pz));
                val[j] = v * sg * (1.0f + pz * (1.0f - sg));
              } else {
                float o = p.relu ? fmaxf(v, 0.f) : v;
                if (p.drop_thresh) o = drop_elem(p, o, dbase, j);
                val[j] = p.res ? o + aux[j] : o;
              }
            }
            if (EPI == EPI_PLAIN) store_block32_h(stg, val, p.out + wrow0 * ldc + c0, ldc, rows_valid, lane);
            else store_block32_h(stg, val, p.out + wrow0 * p.Nout + c0, p.Nout, rows_valid, lane);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// fp32 [rows, cols] row-major (leading dimension ld floats), box [32 columns x box_rows], SWIZZLE_128B
static int make_f32_map(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, CUtensorMap* out,
                        CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
      return set_error(2, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(2, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int EPI>
static int launch(const Params& p, const float* A1, int64_t K1, const float* A2, int64_t K2, const float* W, cudaStream_t st,
                  int64_t lda = 0, int64_t ldw = 0) {
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(node_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error(2, "node_gemm_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  alignas(64) CUtensorMap mA1, mA2, mB;
  if (int rc = make_f32_map(A1, p.M, K1, lda ? lda : K1, BM, &mA1)) return rc;
  if (A2) {
    if (int rc = make_f32_map(A2, p.M, K2, K2, BM, &mA2)) return rc;
  } else {
    memcpy(&mA2, &mA1, sizeof(mA1));
  }
  if (int rc = make_f32_map(W, p.Nout, K1 + K2, ldw ? ldw : K1 + K2, BN, &mB)) return rc;
  const int total = (int)((p.M + BM - 1) / BM) * (p.Nout / BN);
  const int grid = total < sm_count() ? total : sm_count();
  node_gemm_kernel<EPI><<<grid, THREADS, SMEM_BYTES, st>>>(p, mA1, mA2, mB);
  return after_launch("node_gemm_kernel");
}

// =================================================================================================== weight gradients
// C[Mo, 256] = scale * G^T X, G fp32 [N, Mo] (Mo = 256 or 512), X fp32 [N, 256], contraction over the N node rows:
// split-K over row slices (one per CTA and 256-row output block), both operands MN-major (the contraction index is the
// slow one in memory), K-chunks of 32 rows loaded as [32 rows x 128 B] TMA boxes (eight per operand), the 256 x 256
// fp32 block accumulated in TMEM over the CTA's whole slice (2 x 256 columns) and written once as a per-CTA partial;
// launch_partial_reduce sums the partials in a fixed order.  Replaces the cuBLAS g^T x GEMMs of the node-level linears
// (K = 65 536 with 256 x 256 outputs: sixteen CTAs' worth of output tiles for a library GEMM).
// 32-bit MN-major operands need their own swizzle: 32-byte chunks XOR-ed with the row index inside atoms of FOUR
// 128-byte rows (UMMA layout type SWIZZLE_128B_BASE32B = TMA CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B); plain SWIZZLE_128B
// silently produces zeros for kind::tf32.
__device__ __forceinline__ uint64_t desc_mn_32b(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
constexpr int WK = 32;                                  // rows per K-chunk
constexpr int W_BOX = WK * 128;                         // 4 KB: [32 rows][32 floats]
constexpr int W_OP = 8 * W_BOX;                         // 32 KB: 256 columns of one operand
constexpr int W_STAGE = 2 * W_OP;
constexpr int W_STAGES = 3;
constexpr int W_BAR_OFF = W_STAGES * W_STAGE;
constexpr int W_SMEM = W_BAR_OFF + 128 + 1024;
constexpr int W_THREADS = 32 * 6;                       // 0..3 flush, 4 TMA, 5 MMA

struct WParams {
  int64_t N;
  int nblk_x;             // 256-column blocks of X handled by this launch (0 = 1): block id = bg * nblk_x + bx
  int nblk;               // Mo / 256
  int rows_per_slice;     // multiple of WK
  float* partial;         // [slices][nblk][256*256]
};

__global__ void __launch_bounds__(W_THREADS, 1)
node_wgrad_kernel(const WParams p, const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W_BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + W_STAGES;
  uint64_t* done = bars + 2 * W_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbx = p.nblk_x > 0 ? p.nblk_x : 1, nball = p.nblk * nbx;
  const int blk = blockIdx.x % nball, slice = blockIdx.x / nball;
  const int bg = blk / nbx, bxo = (blk % nbx) * 256;
  const int64_t r0 = (int64_t)slice * p.rows_per_slice;
  int64_t r1 = r0 + p.rows_per_slice;
  if (r1 > p.N) r1 = p.N;
  const int chunks = r1 > r0 ? (int)((r1 - r0 + WK - 1) / WK) : 0;     // rows past N are zero-filled by TMA
  if (threadIdx.x == 0) {
    for (int s = 0; s < W_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], W_STAGE);
        uint8_t* dst = smem + stage * W_STAGE;
        const int row = (int)(r0 + (int64_t)c * WK);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          tma_load_2d(dst + b * W_BOX, &mapG, bg * 256 + 32 * b, row, &full[stage]);
          tma_load_2d(dst + W_OP + b * W_BOX, &mapX, bxo + 32 * b, row, &full[stage]);
        }
        if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      // D[m, n] += sum_k G[k, m] X[k, n]: A = G (MN-major: 32 m per 128-byte row, 8 k rows per atom), B = X likewise
      constexpr uint32_t IDESC = idesc_tf32(128, 256) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * W_STAGE);
        const uint32_t b_base = a_base + W_OP;
#pragma unroll
        for (int ks = 0; ks < WK / 8; ++ks)
#pragma unroll
          for (int mh = 0; mh < 2; ++mh)
            umma_tf32(tmem_base + mh * 256, desc_mn_32b(a_base + mh * 4 * W_BOX + ks * 1024, W_BOX, 512),
                      desc_mn_32b(b_base + ks * 1024, W_BOX, 512), IDESC, (c | ks) != 0 ? 1u : 0u);
        umma_commit(&empty[stage]);
        if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
    __syncwarp();
  } else {
    // flush: TMEM -> this CTA's partial block (lane = output row)
    mbar_wait(done, 0);
    tc_fence_after();
    float* dst = p.partial + ((int64_t)slice * nball + blk) * 65536;
#pragma unroll 1
    for (int mh = 0; mh < 2; ++mh) {
      float* drow = dst + (int64_t)(mh * 128 + warp * 32 + lane) * 256;
#pragma unroll 1
      for (int cb = 0; cb < 8; ++cb) {
        uint32_t raw[32];
        if (chunks > 0) {
          tmem_ld32_issue(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mh * 256 + cb * 32), raw);
          tmem_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) raw[j] = 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<uint4*>(drow + cb * 32 + 4 * k) = make_uint4(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2], raw[4 * k + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =================================================================================================== 3xTF32 forms
// fp32-accurate products on the TF32 tensor cores for the exact path (precision="fp32", 1e-5 parity): every operand is
// split as x ~= hi + lo with hi = rna_tf32(x) and lo = rna_tf32(x - hi) (both exact TF32 values, 22 significant bits
// together), and  x w ~= hi_x lo_w + lo_x hi_w + hi_x hi_w  is accumulated in fp32 in TMEM: the
// dropped lo lo term and the truncation of the lo parts are ~2^-22 relative, the level of fp32 rounding itself.  Weights are
// split once per parameter version into a stacked [2 rows, K] image (pev_split_tf32); activations are split INSIDE the
// kernel by four extra warps between the TMA landing and the MMA issue, so they cross HBM once.  Replaces the SIMT
// sgemm calls of the fp32 form of EGNLayer.forward (models/en_gnn_decoder.py:61-79).
__device__ __forceinline__ void split_tf32(const float4 v, float4& hi, float4& lo) {
  uint32_t h0, h1, h2, h3;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(v.x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(v.y));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h2) : "f"(v.z));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h3) : "f"(v.w));
  hi = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
  // lo rounded to TF32 here: the tensor core would TRUNCATE it, a bias that adds up coherently over K (measured 3.7e-6 at
  // K = 512 against 8.8e-7 for an fp32 library GEMM)
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h0) : "f"(v.x - hi.x));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h1) : "f"(v.y - hi.y));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h2) : "f"(v.z - hi.z));
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h3) : "f"(v.w - hi.w));
  lo = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(h2), __uint_as_float(h3));
}
// hi in place, lo at +lo_off, over `bytes` of shared memory, by the 128 threads of the split warps
__device__ __forceinline__ void split_region(uint8_t* base, int lo_off, int bytes, int tid) {
#pragma unroll 4
  for (int off = tid * 16; off < bytes; off += 128 * 16) {
    float4 hi, lo;
    split_tf32(*reinterpret_cast<const float4*>(base + off), hi, lo);
    *reinterpret_cast<float4*>(base + off) = hi;
    *reinterpret_cast<float4*>(base + lo_off + off) = lo;
  }
}

namespace x3 {
// Tile 128 x 128 (not 128 x 256): the 512 TMEM columns then hold THREE accumulators per tile beside each other -- the
// hi hi products of the first and of the second half of K, and the two small products -- because the tensor core
// TRUNCATES every addition into its fp32 accumulator: the error of a product grows with the number of MMA instructions
// per accumulator (measured with everything in one: 3.8e-6 at K = 512; small products apart: 1.6e-6; library fp32
// GEMM: 0.9e-6).  In the small accumulator a truncation costs 2^-11 of what it costs in a main one.
constexpr int BN3 = 128;
constexpr int B3_BYTES = BN3 * BK * 4;                     // 16 KB
constexpr int STAGES = 3;
constexpr int A_LO_OFF = A_BYTES;                          // [A hi 16 KB][A lo 16 KB][W hi 16 KB][W lo 16 KB]
constexpr int B_HI_OFF = 2 * A_BYTES;
constexpr int B_LO_OFF = 2 * A_BYTES + B3_BYTES;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B3_BYTES;    // 64 KB
constexpr int STG_OFF = STAGES * STAGE_BYTES;
constexpr int BAR_OFF = STG_OFF + 8 * 4096;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;
constexpr int SPLIT_WARP0 = 10;
constexpr int THREADS = 32 * 14;                           // 0..7 epilogue, 8 TMA, 9 MMA, 10..13 split
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
}  // namespace x3

// out[M, Nout] = A[M, K] W^T (+ bias) (+ res), W given as its split image W3 = [hi ; lo] ([2 Nout, K]).
__global__ void __launch_bounds__(x3::THREADS, 1)
node_gemm3_kernel(const Params p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + x3::BAR_OFF);
  uint64_t* full = bars;                       // [3] TMA -> split warps, MMA
  uint64_t* split = bars + 3;                  // [3] split warps -> MMA
  uint64_t* empty = bars + 6;                  // [3] MMA -> TMA
  uint64_t* tfull = bars + 9;
  uint64_t* tempty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.Nout / x3::BN3;
  const int total = (int)((p.M + BM - 1) / BM) * n_tiles;
  const int kchunks = p.k1_chunks;
  const int khalf = (kchunks + 1) / 2;         // chunks [0, khalf) -> accumulator a, the rest -> b
  if (threadIdx.x == 0) {
    for (int s = 0; s < x3::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == TMA_WARP) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * x3::BN3;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], A_BYTES + 2 * x3::B3_BYTES);
          uint8_t* dst = smem + stage * x3::STAGE_BYTES;
          tma_load_2d(dst, &mapA, kc * BK, m0, &full[stage]);
          tma_load_2d(dst + x3::B_HI_OFF, &mapB, kc * BK, n0, &full[stage]);
          tma_load_2d(dst + x3::B_LO_OFF, &mapB, kc * BK, p.Nout + n0, &full[stage]);
          if (++stage == x3::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_tf32(BM, x3::BN3);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        mbar_wait(tempty, (it & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          mbar_wait(&split[stage], phase);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * x3::STAGE_BYTES);
          const uint32_t a_lo = a_hi + x3::A_LO_OFF, b_hi = a_hi + x3::B_HI_OFF, b_lo = a_hi + x3::B_LO_OFF;
          const bool second = kc >= khalf;
          const uint32_t dmain = tmem_base + (second ? x3::BN3 : 0), dsmall = tmem_base + 2 * x3::BN3;
          const int kfirst = second ? khalf : 0;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            umma_tf32(dsmall, desc_kmajor(a_hi + ks * 32), desc_kmajor(b_lo + ks * 32), IDESC, (kc | ks) != 0 ? 1u : 0u);
            umma_tf32(dsmall, desc_kmajor(a_lo + ks * 32), desc_kmajor(b_hi + ks * 32), IDESC, 1u);
            umma_tf32(dmain, desc_kmajor(a_hi + ks * 32), desc_kmajor(b_hi + ks * 32), IDESC, ((kc - kfirst) | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == x3::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull);
      }
    }
    __syncwarp();
  } else if (warp >= x3::SPLIT_WARP0) {
    const int tid = threadIdx.x - 32 * x3::SPLIT_WARP0;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x)
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&full[stage], phase);
        split_region(smem + stage * x3::STAGE_BYTES, x3::A_LO_OFF, A_BYTES, tid);
        fence_proxy_async();
        mbar_arrive(&split[stage]);
        if (++stage == x3::STAGES) { stage = 0; phase ^= 1; }
      }
  } else {
    const int q = warp & 3, hh = warp >> 2;            // TMEM lane quarter, 64-column half of the tile
    float* stg = reinterpret_cast<float*>(smem + x3::STG_OFF) + warp * 1024;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int m0 = (t / n_tiles) * BM, n0 = (t % n_tiles) * x3::BN3;
      const int64_t row = (int64_t)m0 + q * 32 + lane;
      const bool valid = row < p.M;
      const int64_t wrow0 = (int64_t)m0 + q * 32;
      const int rows_valid = (int)((p.M - wrow0) < 0 ? 0 : ((p.M - wrow0) > 32 ? 32 : (p.M - wrow0)));
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(hh * 64);
      mbar_wait(tfull, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int b = 0; b < 2; ++b) {
        const int c0 = n0 + hh * 64 + 32 * b;
        uint32_t raw[32];
        tmem_ld32_issue(taddr + 2 * x3::BN3 + 32 * b, raw);              // small products first
        float val[32], rsd[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          val[j] = p.bias ? __ldg(p.bias + c0 + j) : 0.f;
          rsd[j] = 0.f;
        }
        if (p.res) {
          const float* arow = p.res + (valid ? row : 0) * (p.ldres ? p.ldres : (int64_t)p.Nout) + c0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 v = *reinterpret_cast<const float4*>(arow + 4 * j4);
            rsd[4 * j4] = v.x; rsd[4 * j4 + 1] = v.y; rsd[4 * j4 + 2] = v.z; rsd[4 * j4 + 3] = v.w;
          }
        }
        tmem_wait();
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = __uint_as_float(raw[j]);
        if (kchunks > 1) {
          tmem_ld32_issue(taddr + x3::BN3 + 32 * b, raw);
          tmem_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(raw[j]);
        }
        tmem_ld32_issue(taddr + 32 * b, raw);
        tmem_wait();
        if (b == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        const uint32_t dbase = drop_base(p, row, c0);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float o = val[j] + (acc[j] + __uint_as_float(raw[j]));
          if (p.relu) o = fmaxf(o, 0.f);
          if (p.drop_thresh) o = drop_elem(p, o, dbase, j);
          val[j] = o + rsd[j];
        }
        {
          const int64_t ldc = p.ldc ? p.ldc : (int64_t)p.Nout;
          store_block32(stg, val, p.out + wrow0 * ldc + c0, ldc, rows_valid, lane);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Weight gradients, 3xTF32: as node_wgrad_kernel with K-chunks of 16 rows, both operands split in shared memory.
namespace w3 {
constexpr int WK = 16;
constexpr int BOX = WK * 128;                    // 2 KB: [16 rows][32 floats]
constexpr int OP = 8 * BOX;                      // 16 KB: 256 columns of one operand
constexpr int LO_OFF = 2 * OP;                   // [G hi][X hi][G lo][X lo]
constexpr int STAGE = 4 * OP;                    // 64 KB
constexpr int STAGES = 3;
constexpr int BAR_OFF = STAGES * STAGE;
constexpr int SMEM = BAR_OFF + 128 + 1024;
constexpr int THREADS = 32 * 10;                 // 0..3 flush, 4 TMA, 5 MMA, 6..9 split
// The tensor core truncates every addition into its accumulator, an error that grows with the number of MMA
// instructions per output: TMEM is flushed into the CTA's fp32 partial (round-to-nearest adds, L2-resident) every SUB chunks.
constexpr int SUB = 32;                          // 512 rows = 192 MMA instructions per output element and flush
}  // namespace w3

__global__ void __launch_bounds__(w3::THREADS, 1)
node_wgrad3_kernel(const WParams p, const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapX) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + w3::BAR_OFF);
  uint64_t* full = bars;
  uint64_t* split = bars + w3::STAGES;
  uint64_t* empty = bars + 2 * w3::STAGES;
  uint64_t* done = bars + 3 * w3::STAGES;          // MMA -> flush warps: a sub-slice is accumulated
  uint64_t* flushed = done + 1;                    // flush warps -> MMA: TMEM may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(flushed + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nbx = p.nblk_x > 0 ? p.nblk_x : 1, nball = p.nblk * nbx;
  const int blk = blockIdx.x % nball, slice = blockIdx.x / nball;
  const int bg = blk / nbx, bxo = (blk % nbx) * 256;
  const int64_t r0 = (int64_t)slice * p.rows_per_slice;
  int64_t r1 = r0 + p.rows_per_slice;
  if (r1 > p.N) r1 = p.N;
  const int chunks = r1 > r0 ? (int)((r1 - r0 + w3::WK - 1) / w3::WK) : 0;
  const int subs = (chunks + w3::SUB - 1) / w3::SUB;
  if (threadIdx.x == 0) {
    for (int s = 0; s < w3::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 128);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    mbar_init(flushed, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < chunks; ++c) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], 2 * w3::OP);
        uint8_t* dst = smem + stage * w3::STAGE;
        const int row = (int)(r0 + (int64_t)c * w3::WK);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          tma_load_2d(dst + b * w3::BOX, &mapG, bg * 256 + 32 * b, row, &full[stage]);
          tma_load_2d(dst + w3::OP + b * w3::BOX, &mapX, bxo + 32 * b, row, &full[stage]);
        }
        if (++stage == w3::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t IDESC = idesc_tf32(128, 256) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int c = 0; c < chunks; ++c) {
        const int cs = c % w3::SUB;                 // position inside the sub-slice
        if (cs == 0 && c > 0) {
          mbar_wait(flushed, ((c / w3::SUB) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait(&full[stage], phase);
        mbar_wait(&split[stage], phase);
        tc_fence_after();
        const uint32_t g_hi = smem_u32(smem + stage * w3::STAGE);
        const uint32_t x_hi = g_hi + w3::OP, g_lo = g_hi + w3::LO_OFF, x_lo = x_hi + w3::LO_OFF;
#pragma unroll
        for (int ks = 0; ks < w3::WK / 8; ++ks)
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint32_t ao = mh * 4 * w3::BOX + ks * 1024, bo = ks * 1024;
            const uint32_t d = tmem_base + mh * 256;
            umma_tf32(d, desc_mn_32b(g_hi + ao, w3::BOX, 512), desc_mn_32b(x_lo + bo, w3::BOX, 512), IDESC, (cs | ks) != 0 ? 1u : 0u);
            umma_tf32(d, desc_mn_32b(g_lo + ao, w3::BOX, 512), desc_mn_32b(x_hi + bo, w3::BOX, 512), IDESC, 1u);
            umma_tf32(d, desc_mn_32b(g_hi + ao, w3::BOX, 512), desc_mn_32b(x_hi + bo, w3::BOX, 512), IDESC, 1u);
          }
        umma_commit(&empty[stage]);
        if (cs == w3::SUB - 1 || c == chunks - 1) umma_commit(done);
        if (++stage == w3::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 6) {
    const int tid = threadIdx.x - 32 * 6;
    int stage = 0;
    uint32_t phase = 0;
    for (int c = 0; c < chunks; ++c) {
      mbar_wait(&full[stage], phase);
      split_region(smem + stage * w3::STAGE, w3::LO_OFF, 2 * w3::OP, tid);
      fence_proxy_async();
      mbar_arrive(&split[stage]);
      if (++stage == w3::STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // flush: after every sub-slice TMEM is ADDED (fp32, round to nearest) into this CTA's partial block, lane = output row
    float* dst = p.partial + ((int64_t)slice * nball + blk) * 65536;
    for (int ss = 0; ss < (subs > 0 ? subs : 1); ++ss) {
      if (chunks > 0) {
        mbar_wait(done, ss & 1);
        tc_fence_after();
      }
#pragma unroll 1
      for (int mh = 0; mh < 2; ++mh) {
        float* drow = dst + (int64_t)(mh * 128 + warp * 32 + lane) * 256;
#pragma unroll 1
        for (int cb = 0; cb < 8; ++cb) {
          uint32_t raw[32];
          if (chunks > 0) {
            tmem_ld32_issue(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mh * 256 + cb * 32), raw);
            tmem_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) raw[j] = 0u;
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float4 v = make_float4(__uint_as_float(raw[4 * k]), __uint_as_float(raw[4 * k + 1]), __uint_as_float(raw[4 * k + 2]),
                                   __uint_as_float(raw[4 * k + 3]));
            float4* q = reinterpret_cast<float4*>(drow + cb * 32 + 4 * k);
            if (ss > 0) {
              const float4 o = *q;
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *q = v;
          }
        }
      }
      if (chunks > 0 && ss + 1 < subs) {
        tc_fence_before();
        mbar_arrive(flushed);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// W fp32 [rows, cols] -> out [2 R, C] = [hi ; lo] of W (transpose = 0: R = rows, C = cols) or of W^T (R = cols, C = rows)
__global__ void split_weight_kernel(const float* __restrict__ W, int rows, int cols, int transpose, float* __restrict__ out) {
  const int R = transpose ? cols : rows, C = transpose ? rows : cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)R * C; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    const float v = transpose ? W[(int64_t)c * cols + r] : W[i];
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    out[i] = __uint_as_float(h);
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - __uint_as_float(h)));
    out[(int64_t)R * C + i] = __uint_as_float(l);
  }
}

// out block (bg, bx) [256 x 256] (leading dimension ldc) = scale * sum over slices (fixed order) of the per-CTA partials
__global__ void __launch_bounds__(256)
wgrad_blocks_reduce_kernel(const float* __restrict__ partial, int slices, int nball, int nbx, float scale, float* __restrict__ out,
                           int64_t ldc, int rows_total) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)nball * 16384) return;
  const int blk = (int)(idx >> 14), e = (int)(idx & 16383);
  const float4* src = reinterpret_cast<const float4*>(partial) + (int64_t)blk * 16384 + e;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < slices; ++s) {
    const float4 v = src[(int64_t)s * nball * 16384];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int r = e >> 6, c4 = e & 63;
  if ((blk / nbx) * 256 + r >= rows_total) return;           // Mo = 128 (mod 256): the last row block is half empty
  float* dst = out + ((int64_t)(blk / nbx) * 256 + r) * ldc + (blk % nbx) * 256 + 4 * c4;
  *reinterpret_cast<float4*>(dst) = make_float4(scale * acc.x, scale * acc.y, scale * acc.z, scale * acc.w);
}

}  // namespace ng
}  // namespace pev

using namespace pev;

extern "C" int64_t pev_node_wgrad_workspace_bytes(void) { return (int64_t)sm_count() * 65536 * (int64_t)sizeof(float); }

extern "C" int pev_node_wgrad(const float* G, int32_t Mo, const float* X, int64_t N, float scale, float* workspace,
                              float* out, int32_t ldc, void* stream) {
  PEV_REQUIRE(G && X && out && workspace && N >= 0 && (Mo == 256 || Mo == 512) && ldc >= 256, "bad argument");
  cudaStream_t st = as_stream(stream);
  const int nblk = Mo / 256;
  if (N == 0) {
    for (int r = 0; r < Mo; ++r) cudaMemsetAsync(out + (int64_t)r * ldc, 0, sizeof(float) * 256, st);
    return 0;
  }
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ng::node_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::W_SMEM);
    if (e != cudaSuccess) return set_error(2, "node_wgrad_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  int slices = sm_count() / nblk;
  const int64_t chunks = (N + ng::WK - 1) / ng::WK;
  if (slices > chunks) slices = (int)chunks;
  ng::WParams p = {};
  p.N = N; p.nblk = nblk; p.partial = workspace;
  p.rows_per_slice = (int)(((chunks + slices - 1) / slices) * ng::WK);
  slices = (int)((N + p.rows_per_slice - 1) / p.rows_per_slice);
  alignas(64) CUtensorMap mG, mX;
  if (int rc = ng::make_f32_map(G, N, Mo, Mo, ng::WK, &mG, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (int rc = ng::make_f32_map(X, N, 256, 256, ng::WK, &mX, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  ng::node_wgrad_kernel<<<slices * nblk, ng::W_THREADS, ng::W_SMEM, st>>>(p, mG, mX);
  if (int rc = after_launch("node_wgrad_kernel")) return rc;
  // partial layout [slice][blk][256*256]; output block blk = rows 256 blk .. of out (leading dimension ldc)
  for (int b = 0; b < nblk; ++b) {
    if (ldc == 256) {
      if (int rc = launch_partial_reduce(workspace + (int64_t)b * 65536, slices, (int64_t)nblk * 65536, 65536, scale,
                                         out + (int64_t)b * 65536, st)) return rc;
    } else {
      if (int rc = launch_partial_reduce_2d(workspace + (int64_t)b * 65536, slices, (int64_t)nblk * 65536, 256, 256, scale,
                                            out + (int64_t)b * 256 * ldc, ldc, st)) return rc;
    }
  }
  return 0;
}

extern "C" int pev_node_gemm(int32_t epilogue, const float* A1, int32_t K1, const float* A2, int32_t K2, const float* W,
                             const float* bias, int64_t M, int32_t Nout, float scale, const float* aux, const float* gamma,
                             const float* beta, float eps, void* out, float* out2, float* mean, float* rstd, void* stream) {
  PEV_REQUIRE(A1 && W && out && M >= 0, "null argument");
  PEV_REQUIRE((Nout == 256 || Nout == 512) && K1 > 0 && K1 % 32 == 0 && K2 >= 0 && K2 % 32 == 0 && (K2 == 0) == (A2 == nullptr),
              "shape: Nout in {256, 512}, K1 / K2 multiples of 32");
  if (M == 0) return 0;
  ng::Params p = {};
  p.M = M; p.Nout = Nout; p.k1_chunks = K1 / 32; p.k2_chunks = K2 / 32; p.bias = bias; p.scale = scale;
  p.res = aux; p.gamma = gamma; p.beta = beta; p.eps = eps; p.out2 = out2; p.mean = mean; p.rstd = rstd;
  cudaStream_t st = as_stream(stream);
  switch (epilogue) {
    case ng::EPI_ABH:
      p.out16 = reinterpret_cast<__half*>(out);
      return ng::launch<ng::EPI_ABH>(p, A1, K1, A2, K2, W, st);
    case ng::EPI_SILU:
      p.out = reinterpret_cast<float*>(out);
      return ng::launch<ng::EPI_SILU>(p, A1, K1, A2, K2, W, st);
    case ng::EPI_RES_LN:
      PEV_REQUIRE(Nout == 256 && aux && gamma && beta && (mean == nullptr) == (rstd == nullptr), "RES_LN: Nout = 256, h / gamma / beta");
      p.out = reinterpret_cast<float*>(out);
      return ng::launch<ng::EPI_RES_LN>(p, A1, K1, A2, K2, W, st);
    case ng::EPI_PLAIN:
      p.out = reinterpret_cast<float*>(out);
      return ng::launch<ng::EPI_PLAIN>(p, A1, K1, A2, K2, W, st);
    case ng::EPI_DSILU:
      PEV_REQUIRE(aux, "DSILU needs p");
      p.out = reinterpret_cast<float*>(out);
      return ng::launch<ng::EPI_DSILU>(p, A1, K1, A2, K2, W, st);
    default:
      return set_error(1, "pev_node_gemm: unknown epilogue %d", epilogue);
  }
}

extern "C" int pev_split_tf32(const float* W, int32_t rows, int32_t cols, int32_t transpose, float* out, void* stream) {
  PEV_REQUIRE(W && out && rows > 0 && cols > 0, "bad argument");
  const int64_t n = (int64_t)rows * cols;
  int grid = (int)((n + 255) / 256);
  if (grid > 4 * sm_count()) grid = 4 * sm_count();
  ng::split_weight_kernel<<<grid, 256, 0, as_stream(stream)>>>(W, rows, cols, transpose, out);
  return after_launch("split_weight_kernel");
}

extern "C" int pev_node_gemm3(const float* A, int32_t K, const float* W3, const float* bias, int64_t M, int32_t Nout,
                              const float* res, float* out, void* stream) {
  PEV_REQUIRE(A && W3 && out && M >= 0, "null argument");
  PEV_REQUIRE((Nout == 256 || Nout == 512) && K > 0 && K % 32 == 0, "shape: Nout in {256, 512}, K a multiple of 32");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ng::node_gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::x3::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(2, "node_gemm3_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  ng::Params p = {};
  p.M = M; p.Nout = Nout; p.k1_chunks = K / 32; p.bias = bias; p.res = res; p.out = out;
  alignas(64) CUtensorMap mA, mB;
  if (int rc = ng::make_f32_map(A, M, K, K, ng::BM, &mA)) return rc;
  if (int rc = ng::make_f32_map(W3, 2 * (int64_t)Nout, K, K, ng::x3::BN3, &mB)) return rc;
  const int total = (int)((M + ng::BM - 1) / ng::BM) * (Nout / ng::x3::BN3);
  const int grid = total < sm_count() ? total : sm_count();
  ng::node_gemm3_kernel<<<grid, ng::x3::THREADS, ng::x3::SMEM_BYTES, st>>>(p, mA, mB);
  return after_launch("node_gemm3_kernel");
}

extern "C" int pev_node_wgrad3(const float* G, int32_t Mo, const float* X, int64_t N, float scale, float* workspace,
                               float* out, int32_t ldc, void* stream) {
  PEV_REQUIRE(G && X && out && workspace && N >= 0 && (Mo == 256 || Mo == 512) && ldc >= 256, "bad argument");
  cudaStream_t st = as_stream(stream);
  const int nblk = Mo / 256;
  if (N == 0) {
    for (int r = 0; r < Mo; ++r) cudaMemsetAsync(out + (int64_t)r * ldc, 0, sizeof(float) * 256, st);
    return 0;
  }
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ng::node_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::w3::SMEM);
    if (e != cudaSuccess) return set_error(2, "node_wgrad3_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  int slices = sm_count() / nblk;
  const int64_t chunks = (N + ng::w3::WK - 1) / ng::w3::WK;
  if (slices > chunks) slices = (int)chunks;
  ng::WParams p = {};
  p.N = N; p.nblk = nblk; p.partial = workspace;
  p.rows_per_slice = (int)(((chunks + slices - 1) / slices) * ng::w3::WK);
  slices = (int)((N + p.rows_per_slice - 1) / p.rows_per_slice);
  alignas(64) CUtensorMap mG, mX;
  if (int rc = ng::make_f32_map(G, N, Mo, Mo, ng::w3::WK, &mG, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (int rc = ng::make_f32_map(X, N, 256, 256, ng::w3::WK, &mX, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  ng::node_wgrad3_kernel<<<slices * nblk, ng::w3::THREADS, ng::w3::SMEM, st>>>(p, mG, mX);
  if (int rc = after_launch("node_wgrad3_kernel")) return rc;
  for (int b = 0; b < nblk; ++b) {
    if (ldc == 256) {
      if (int rc = launch_partial_reduce(workspace + (int64_t)b * 65536, slices, (int64_t)nblk * 65536, 65536, scale,
                                         out + (int64_t)b * 65536, st)) return rc;
    } else {
      if (int rc = launch_partial_reduce_2d(workspace + (int64_t)b * 65536, slices, (int64_t)nblk * 65536, 256, 256, scale,
                                            out + (int64_t)b * 256 * ldc, ldc, st)) return rc;
    }
  }
  return 0;
}

// General linear layer on the node-level GEMM kernels: out[M,Nout] (leading dimension ldc) = act(A[M,K] W^T + bias + res),
// A / res with leading dimensions lda / ldres; precise = 0: TF32 (W fp32 [Nout,K], Nout a multiple of 256), precise = 1:
// 3xTF32 (W = split image [2 Nout, K], Nout a multiple of 128).
extern "C" int pev_linear(int32_t precise, const float* A, int64_t lda, int32_t K, const float* W, const float* bias, int64_t M,
                          int32_t Nout, int32_t relu, float p_drop, uint32_t seed, const float* res, int64_t ldres, float* out,
                          int64_t ldc, void* stream) {
  PEV_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "p_drop in [0, 1)");
  PEV_REQUIRE(A && W && out && M >= 0 && K > 0 && K % 32 == 0 && lda >= K && ldc >= Nout && (!res || ldres >= Nout), "bad argument");
  PEV_REQUIRE(Nout > 0 && Nout % (precise ? 128 : 256) == 0, "Nout: a multiple of 256 (TF32) / 128 (3xTF32)");
  PEV_REQUIRE(lda % 4 == 0 && ldc % 4 == 0 && (!res || ldres % 4 == 0), "leading dimensions: multiples of 4 floats");
  if (M == 0) return 0;
  cudaStream_t st = as_stream(stream);
  ng::Params p = {};
  p.M = M; p.Nout = Nout; p.k1_chunks = K / 32; p.bias = bias; p.res = res; p.out = out; p.ldc = ldc; p.ldres = ldres; p.relu = relu;
  p.drop_thresh = p_drop > 0.f ? (uint32_t)(p_drop * 4294967296.0) : 0u;
  p.drop_seed = seed;
  p.drop_scale = 1.0f / (1.0f - p_drop);
  if (!precise) return ng::launch<ng::EPI_PLAIN>(p, A, K, nullptr, 0, W, st, lda);
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ng::node_gemm3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::x3::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(2, "node_gemm3_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  alignas(64) CUtensorMap mA, mB;
  if (int rc = ng::make_f32_map(A, M, K, lda, ng::BM, &mA)) return rc;
  if (int rc = ng::make_f32_map(W, 2 * (int64_t)Nout, K, K, ng::x3::BN3, &mB)) return rc;
  const int total = (int)((M + ng::BM - 1) / ng::BM) * (Nout / ng::x3::BN3);
  const int grid = total < sm_count() ? total : sm_count();
  ng::node_gemm3_kernel<<<grid, ng::x3::THREADS, ng::x3::SMEM_BYTES, st>>>(p, mA, mB);
  return after_launch("node_gemm3_kernel");
}

// Weight gradient of a linear layer in ONE launch: out[Mo, Kx] (leading dimension ldc) = scale * G^T X for G [N, Mo] (ldg) and
// X [N, Kx] (ldx), Mo and Kx multiples of 256: every 256 x 256 output block gets sm_count / blocks row slices, all blocks run
// concurrently (the CTAs that share a row slice of G or X meet in L2: HBM sees each operand once), one fixed-order reduction.
extern "C" int pev_linear_wgrad(int32_t precise, const float* G, int64_t ldg, int32_t Mo, const float* X, int64_t ldx, int32_t Kx,
                                int64_t N, float scale, float* workspace, float* out, int64_t ldc, void* stream) {
  PEV_REQUIRE(G && X && out && workspace && N >= 0 && Mo > 0 && Mo % 128 == 0 && Kx > 0 && Kx % 256 == 0 && ldc >= Kx &&
              ldg >= Mo && ldx >= Kx && ldg % 4 == 0 && ldx % 4 == 0 && ldc % 4 == 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  const int nbg = (Mo + 255) / 256, nbx = Kx / 256, nball = nbg * nbx;   // columns of G past Mo are zero-filled by TMA
  PEV_REQUIRE(nball <= sm_count(), "too many output blocks for one launch");
  if (N == 0) {
    for (int r = 0; r < Mo; ++r) cudaMemsetAsync(out + (int64_t)r * ldc, 0, sizeof(float) * Kx, st);
    return 0;
  }
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ng::node_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::W_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ng::node_wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ng::w3::SMEM);
    if (e != cudaSuccess) return set_error(2, "node_wgrad kernels: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int wk = precise ? ng::w3::WK : ng::WK;
  int slices = sm_count() / nball;
  const int64_t chunks = (N + wk - 1) / wk;
  if (slices > chunks) slices = (int)chunks;
  ng::WParams p = {};
  p.N = N; p.nblk = nbg; p.nblk_x = nbx; p.partial = workspace;
  p.rows_per_slice = (int)(((chunks + slices - 1) / slices) * wk);
  slices = (int)((N + p.rows_per_slice - 1) / p.rows_per_slice);
  alignas(64) CUtensorMap mG, mX;
  if (int rc = ng::make_f32_map(G, N, Mo, ldg, wk, &mG, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (int rc = ng::make_f32_map(X, N, Kx, ldx, wk, &mX, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return rc;
  if (precise) ng::node_wgrad3_kernel<<<slices * nball, ng::w3::THREADS, ng::w3::SMEM, st>>>(p, mG, mX);
  else ng::node_wgrad_kernel<<<slices * nball, ng::W_THREADS, ng::W_SMEM, st>>>(p, mG, mX);
  if (int rc = after_launch("node_wgrad_kernel")) return rc;
  ng::wgrad_blocks_reduce_kernel<<<(nball * 16384 + 255) / 256, 256, 0, st>>>(workspace, slices, nball, nbx, scale, out, ldc, Mo);
  return after_launch("wgrad_blocks_reduce_kernel");
}
