// Per-conformer multi-head attention of the encoder (models/encoder.py:125-140: nn.MultiheadAttention with a key-padding
// mask) on tcgen05, as three batched TF32 GEMM forms over the PACKED rows of a batch plus two row kernels:
//
//   NT  S[h, i, j]  = scale * <A[i, h hd : (h+1) hd], B[j, ...]>          scores Q K^T, and dP = dO V^T in the backward pass
//   NN  O[i, h hd+c] = sum_j P[h, i, j] B[j, h hd + c]                      P V, and dQ = dS K
//   TN  O[j, h hd+c] = sum_i P[h, i, j] B[i, h hd + c]                      dV = P^T dO, and dK = dS^T Q
//   softmax rows of S in place (keys of the row's own conformer only), and its backward dS = P (dP - sum_j P dP)
//
// A / B / O are row-major fp32 tensors over the N packed residues (Q, K, V are column blocks of the [N, 3 d] projection,
// never copied apart).  The score buffer S / P is [H, Np, Lpad]: every conformer's rows padded to a multiple of 128 (so an
// M tile never straddles two conformers) and written as zeros there and beyond the conformer's length, which is what
// makes the ragged edges safe: a K-chunk that runs past a conformer's end multiplies finite foreign rows by exact zeros.
// Tiles: 128 rows x (256 | hd) columns, K-chunks of 32, operands by TMA (K-major SWIZZLE_128B; MN-major operands in the
// 32-bit layout SWIZZLE_128B_BASE32B, see node_gemm_kernels.cu), 3-stage ring, two TMEM accumulator stages, 16 epilogue
// warps (lane quarter x 4 column groups).  The softmax and its backward run inside the epilogues of the two NT GEMMs (row
// statistics exchanged through shared memory; dS = P (dP - <dO, O>) needs no row pass at all), so raw scores and dP never
// reach HBM: per layer the probability buffer is written once and read three times, the dS buffer written once and read twice.
// Rows longer than 256 keys fall back to the separate softmax kernel in the forward pass.
#include <cuda.h>

#include <cstring>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "tc_common.cuh"

namespace pev {
namespace at {
using namespace tcx;

constexpr int BM = 128, BK = 32;
constexpr int STAGES = 3;
constexpr int A_BYTES = BM * BK * 4;             // 16 KB
constexpr int B_BYTES = 256 * BK * 4;            // 32 KB (NT); NN / TN use hd / 32 boxes of 4 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BOX = 32 * 128;                    // [32 rows][32 floats]
constexpr int STG_OFF = STAGES * STAGE_BYTES;
constexpr int BAR_OFF = STG_OFF + 16 * 4096;
constexpr int EX_OFF = BAR_OFF + 256;            // [4 column groups][128 rows] row statistics of the fused softmax
constexpr int SMEM_BYTES = EX_OFF + 2048 + 1024;
constexpr int EPI_WARPS = 16, TMA_WARP = 16, MMA_WARP = 17;     // 16 epilogue warps: lane quarter x 4 column groups
constexpr int THREADS = 32 * 18;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

enum Form { NT = 0, NN = 1, TN = 2 };

struct Params {
  const int32_t* tile_conf;   // [m_tiles] conformer of each 128-row tile of the padded layout
  const int32_t* cu;          // [B+1] packed row offsets
  const int32_t* cup;         // [B+1] padded row offsets (multiples of 128)
  int m_tiles, H, hd, Lpad;
  int64_t Np;
  int a_col0, b_col0;         // first column of the A / B operand inside its packed tensor (head 0)
  float scale;
  float* out;
  int64_t ldo;
  int out_col0;
  // NT only.  mode 1: softmax fused into the epilogue (needs Lpad <= 256: one key tile per row): out = probabilities, out2 =
  // dropout-kept probabilities / (1 - p) or null.  mode 2: softmax backward fused: out = dS = P (keep dP / (1 - p) - delta),
  // P = stored probabilities, delta[h Np + padded row] = <dO, O> of that row and head.
  int mode;
  float* out2;
  const float* P;
  const float* delta;
  uint32_t drop_thresh, drop_seed;
  float drop_scale;
};

__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint64_t desc_mn_32b(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
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

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {      // counter-based dropout bits (lowbias32 integer hash)
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool keep_bit(uint32_t seed, int64_t idx, uint32_t thresh) {
  if (thresh == 0u) return true;                             // no dropout: no hashing (the row kernels are issue-bound on it)
  return mix32(seed + (uint32_t)idx * 0x9e3779b9U + (uint32_t)(idx >> 32)) >= thresh;
}

// One tile = (head, 128-row tile of the padded layout[, 256-column tile of the scores]); every role walks the same list.
struct Tile {
  int h, b, m0, n0, Lb, kchunks;
  int64_t arow, brow, prow;   // first packed row of the conformer (A / B operands), first padded row of the tile
  bool skip;
};
template <int FORM>
__device__ __forceinline__ Tile get_tile(const Params& p, int t, int n_tiles) {
  Tile T;
  const int nt = FORM == NT ? t % n_tiles : 0;
  const int r = FORM == NT ? t / n_tiles : t;
  const int mt = r % p.m_tiles;
  T.h = r / p.m_tiles;
  T.b = p.tile_conf[mt];
  const int c0 = p.cu[T.b], cp0 = p.cup[T.b];
  T.Lb = p.cu[T.b + 1] - c0;
  T.m0 = mt * BM - cp0;
  T.n0 = nt * 256;
  T.arow = c0;
  T.brow = c0;
  T.prow = (int64_t)T.h * p.Np + (int64_t)mt * BM;
  T.kchunks = FORM == NT ? p.hd / BK : (T.Lb + BK - 1) / BK;
  // NT: key tiles past the conformer's length hold nothing; NN: query rows past it produce nothing; TN: key columns likewise
  T.skip = FORM == NT ? (T.n0 >= T.Lb || T.m0 >= T.Lb) : (T.m0 >= T.Lb);
  return T;
}

template <int FORM>
__global__ void __launch_bounds__(THREADS, 1)
attn_gemm_kernel(const Params p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = FORM == NT ? (p.Lpad + 255) / 256 : 1;
  const int total = p.H * p.m_tiles * n_tiles;
  const int nb = p.hd / 32;                                  // 32-column boxes of an hd-wide MN-major operand
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
        const Tile T = get_tile<FORM>(p, t, n_tiles);
        if (T.skip) continue;
        for (int kc = 0; kc < T.kchunks; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = smem + stage * STAGE_BYTES;
          if (FORM == NT) {
            mbar_arrive_expect_tx(&full[stage], A_BYTES + B_BYTES);
            tma_load_2d(dst, &mapA, p.a_col0 + T.h * p.hd + kc * BK, (int)(T.arow + T.m0), &full[stage]);
            tma_load_2d(dst + A_BYTES, &mapB, p.b_col0 + T.h * p.hd + kc * BK, (int)(T.brow + T.n0), &full[stage]);
          } else if (FORM == NN) {
            mbar_arrive_expect_tx(&full[stage], A_BYTES + nb * BOX);
            tma_load_2d(dst, &mapA, kc * BK, (int)T.prow, &full[stage]);
            for (int bx = 0; bx < nb; ++bx)
              tma_load_2d(dst + A_BYTES + bx * BOX, &mapB, p.b_col0 + T.h * p.hd + 32 * bx, (int)(T.brow + kc * BK), &full[stage]);
          } else {
            // contraction over the conformer's (padded) query rows; A columns = this tile's 128 keys
            mbar_arrive_expect_tx(&full[stage], 4 * BOX + nb * BOX);
            const int64_t prow0 = (int64_t)T.h * p.Np + p.cup[T.b] + kc * BK;
            for (int bx = 0; bx < 4; ++bx) tma_load_2d(dst + bx * BOX, &mapA, T.m0 + 32 * bx, (int)prow0, &full[stage]);
            for (int bx = 0; bx < nb; ++bx)
              tma_load_2d(dst + A_BYTES + bx * BOX, &mapB, p.b_col0 + T.h * p.hd + 32 * bx, (int)(T.brow + kc * BK), &full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == MMA_WARP) {
    if (lane == 0) {
      const uint32_t idesc = FORM == NT ? idesc_tf32(BM, 256, false, false)
                                        : (FORM == NN ? idesc_tf32(BM, p.hd, false, true) : idesc_tf32(BM, p.hd, true, true));
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const Tile T = get_tile<FORM>(p, t, n_tiles);
        if (T.skip) continue;
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int kc = 0; kc < T.kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_base = a_base + A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t ad = FORM == TN ? desc_mn_32b(a_base + ks * 1024, BOX, 512) : desc_kmajor(a_base + ks * 32);
            const uint64_t bd = FORM == NT ? desc_kmajor(b_base + ks * 32) : desc_mn_32b(b_base + ks * 1024, BOX, 512);
            umma_tf32(d, ad, bd, idesc, (kc | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        ++it;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, cq = warp >> 2;                  // TMEM lane quarter; 32-column blocks cq, cq + 4, ...
    float* stg = reinterpret_cast<float*>(smem + STG_OFF) + warp * 1024;
    const int ncols = FORM == NT ? 256 : p.hd;               // accumulator columns in use
    const int nblocks = ncols / 32;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const Tile T = get_tile<FORM>(p, t, n_tiles);
      if (T.skip) continue;
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      float* obase;
      int64_t ld;
      int rows_valid;
      if (FORM == NT) {
        obase = p.out + (T.prow + q * 32) * p.Lpad + T.n0;
        ld = p.Lpad;
        rows_valid = 32;
      } else {
        obase = p.out + (T.arow + T.m0 + q * 32) * p.ldo + p.out_col0 + T.h * p.hd;
        ld = p.ldo;
        const int left = T.Lb - T.m0 - q * 32;
        rows_valid = left < 0 ? 0 : (left > 32 ? 32 : left);
      }
      if (FORM == NT && p.mode == 1) {
        // ---- softmax of the whole score row in the epilogue: the four warps of a lane quarter hold 64 columns each
        const int r = T.m0 + q * 32 + lane;
        const bool rv = r < T.Lb;
        const int64_t prow = T.prow + q * 32 + lane;
        float* ex = reinterpret_cast<float*>(smem + EX_OFF);
        float* mine = ex + cq * BM + q * 32 + lane;
        const float* col = ex + q * 32 + lane;                 // + g * BM: the value of column group g
        float mx = -INFINITY;
#pragma unroll 1
        for (int b = cq; b < nblocks; b += 4) {
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          tmem_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (32 * b + j < T.Lb) mx = fmaxf(mx, p.scale * __uint_as_float(raw[j]));
        }
        *mine = mx;
        asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory");
        mx = fmaxf(fmaxf(col[0], col[BM]), fmaxf(col[2 * BM], col[3 * BM]));
        asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory");
        float sum = 0.f;
#pragma unroll 1
        for (int b = cq; b < nblocks; b += 4) {
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          tmem_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (32 * b + j < T.Lb) sum += __expf(p.scale * __uint_as_float(raw[j]) - mx);
        }
        *mine = sum;
        asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory");
        sum = (col[0] + col[BM]) + (col[2 * BM] + col[3 * BM]);   // the same order in all four warps
        asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory");
        const float inv = 1.0f / sum;
#pragma unroll 1
        for (int b = cq; b < nblocks; b += 4) {
          const int c0 = 32 * b;
          if (c0 >= p.Lpad) break;
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          tmem_wait();
          float val[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            val[j] = (rv && c0 + j < T.Lb) ? __expf(p.scale * __uint_as_float(raw[j]) - mx) * inv : 0.f;
          store_block32(stg, val, obase + c0, ld, 32, lane);
          if (p.out2) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              val[j] = keep_bit(p.drop_seed, prow * p.Lpad + c0 + j, p.drop_thresh) ? val[j] * p.drop_scale : 0.f;
            store_block32(stg, val, p.out2 + (obase - p.out) + c0, ld, 32, lane);
          }
        }
      } else if (FORM == NT && p.mode == 2) {
        // ---- softmax backward in the epilogue: dS = P (keep dP / (1 - p) - delta)
        const int r = T.m0 + q * 32 + lane;
        const bool rv = r < T.Lb;
        const int64_t prow = T.prow + q * 32 + lane;
        const float dl = p.delta[prow];
#pragma unroll 1
        for (int b = cq; b < nblocks; b += 4) {
          const int c0 = T.n0 + 32 * b;
          if (c0 >= p.Lpad) break;
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          float pv[32];
          const float* prw = p.P + prow * p.Lpad + c0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 v = *reinterpret_cast<const float4*>(prw + 4 * j4);
            pv[4 * j4] = v.x; pv[4 * j4 + 1] = v.y; pv[4 * j4 + 2] = v.z; pv[4 * j4 + 3] = v.w;
          }
          tmem_wait();
          float val[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float g = keep_bit(p.drop_seed, prow * p.Lpad + c0 + j, p.drop_thresh) ? __uint_as_float(raw[j]) * p.drop_scale : 0.f;
            val[j] = (rv && c0 + j < T.Lb) ? pv[j] * (g - dl) : 0.f;
          }
          store_block32(stg, val, obase + 32 * b, ld, 32, lane);
        }
      } else {
#pragma unroll 1
        for (int b = cq; b < nblocks; b += 4) {
          uint32_t raw[32];
          tmem_ld32_issue(taddr + 32 * b, raw);
          tmem_wait();
          float val[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = p.scale * __uint_as_float(raw[j]);
          const bool in_range = FORM != NT || (T.n0 + 32 * b) < p.Lpad;
          if (in_range) store_block32(stg, val, obase + 32 * b, ld, rows_valid, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------- row kernels
// One warp per (head, padded row): softmax over the keys of the row's conformer, zeros elsewhere; with dropout the kept and
// rescaled probabilities go to Pd (the operand of the P V GEMM), the plain ones stay in S (needed by the backward pass).
__global__ void __launch_bounds__(256)
softmax_fwd_kernel(float* __restrict__ S, float* __restrict__ Pd, const int32_t* __restrict__ tile_conf,
                   const int32_t* __restrict__ cu, const int32_t* __restrict__ cup, int64_t Np, int H, int Lpad, float p_drop,
                   uint32_t seed) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (int64_t)H * Np) return;
  const int64_t rp = row % Np;
  const int b = tile_conf[rp / BM];
  const int Lb = cu[b + 1] - cu[b];
  const int i = (int)(rp - cup[b]);
  float* s = S + row * Lpad;
  float* pd = Pd ? Pd + row * Lpad : nullptr;
  if (i >= Lb) {
    for (int j = lane; j < Lpad; j += 32) {
      s[j] = 0.f;
      if (pd) pd[j] = 0.f;
    }
    return;
  }
  const uint32_t thresh = p_drop > 0.f ? (uint32_t)(p_drop * 4294967296.0) : 0u;
  const float keep_scale = 1.0f / (1.0f - p_drop);
  if (Lpad <= 512) {
    // the row lives in registers (16 values per lane): one read and one write of the score buffer
    float v[16];
    float mx = -INFINITY;
    const int nk = Lpad >> 5;                    // live 32-column groups (uniform)
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = lane + 32 * k;
      v[k] = (k < nk && j < Lb) ? s[j] : -INFINITY;
      mx = fmaxf(mx, v[k]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < nk) {
        v[k] = __expf(v[k] - mx);                // exp(-inf) = 0 past the conformer's length
        sum += v[k];
      }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = lane + 32 * k;
      if (k < nk) {
        const float pj = v[k] * inv;
        s[j] = pj;
        if (pd) pd[j] = keep_bit(seed, row * Lpad + j, thresh) ? pj * keep_scale : 0.f;
      }
    }
    return;
  }
  float mx = -INFINITY;
  for (int j = lane; j < Lb; j += 32) mx = fmaxf(mx, s[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < Lb; j += 32) sum += __expf(s[j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < Lpad; j += 32) {
    const float pj = j < Lb ? __expf(s[j] - mx) * inv : 0.f;
    s[j] = pj;
    if (pd) pd[j] = keep_bit(seed, row * Lpad + j, thresh) ? pj * keep_scale : 0.f;
  }
}

// dS = P (dP - sum_j P dP) in place on dP (= G), dropout mask re-derived from the counter hash
__global__ void __launch_bounds__(256)
softmax_bwd_kernel(const float* __restrict__ P, float* __restrict__ G, const int32_t* __restrict__ tile_conf,
                   const int32_t* __restrict__ cu, const int32_t* __restrict__ cup, int64_t Np, int H, int Lpad, float p_drop,
                   uint32_t seed) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (int64_t)H * Np) return;
  const int64_t rp = row % Np;
  const int b = tile_conf[rp / BM];
  const int Lb = cu[b + 1] - cu[b];
  const int i = (int)(rp - cup[b]);
  const float* pr = P + row * Lpad;
  float* g = G + row * Lpad;
  if (i >= Lb) {
    for (int j = lane; j < Lpad; j += 32) g[j] = 0.f;
    return;
  }
  const uint32_t thresh = p_drop > 0.f ? (uint32_t)(p_drop * 4294967296.0) : 0u;
  const float keep_scale = 1.0f / (1.0f - p_drop);
  if (Lpad <= 512) {
    float pv[16], gv[16];
    float dot = 0.f;
    const int nk = Lpad >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int j = lane + 32 * k;
      const bool live = k < nk && j < Lb;
      pv[k] = live ? pr[j] : 0.f;
      gv[k] = live ? g[j] : 0.f;                 // both loads issued before any hashing
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < nk) {
        gv[k] = keep_bit(seed, row * Lpad + lane + 32 * k, thresh) ? gv[k] * keep_scale : 0.f;
        dot = fmaf(pv[k], gv[k], dot);
      }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (k < nk) g[lane + 32 * k] = pv[k] * (gv[k] - dot);
    return;
  }
  float dot = 0.f;
  for (int j = lane; j < Lb; j += 32) {
    const float gj = keep_bit(seed, row * Lpad + j, thresh) ? g[j] * keep_scale : 0.f;
    dot = fmaf(pr[j], gj, dot);
  }
  dot = warp_sum(dot);
  for (int j = lane; j < Lpad; j += 32) {
    float out = 0.f;
    if (j < Lb) {
      const float gj = keep_bit(seed, row * Lpad + j, thresh) ? g[j] * keep_scale : 0.f;
      out = pr[j] * (gj - dot);
    }
    g[j] = out;
  }
}

// delta[h Np + padded row of i] = sum_c dO[i, h hd + c] O[i, h hd + c]  (= sum_j P_ij dP_ij, with or without dropout)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ dO, const float* __restrict__ O, int64_t N, int H, int hd,
                  const int32_t* __restrict__ row_pad, int64_t Np, float* __restrict__ delta) {
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= N) return;
  const int d = H * hd;
  const float* a = dO + i * d;
  const float* b = O + i * d;
  const int64_t rp = row_pad[i];
  for (int h = 0; h < H; ++h) {
    float s = 0.f;
    for (int c = lane; c < hd; c += 32) s = fmaf(a[h * hd + c], b[h * hd + c], s);
    s = warp_sum(s);
    if (lane == 0) delta[(int64_t)h * Np + rp] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_map(const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows, bool mn, CUtensorMap* out) {
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
  const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, mn ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(2, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

template <int FORM>
static int launch(const Params& p, const CUtensorMap& mA, const CUtensorMap& mB, cudaStream_t st) {
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(attn_gemm_kernel<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error(2, "attn_gemm_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int n_tiles = FORM == NT ? (p.Lpad + 255) / 256 : 1;
  const int64_t total = (int64_t)p.H * p.m_tiles * n_tiles;
  const int grid = (int)(total < sm_count() ? total : sm_count());
  attn_gemm_kernel<FORM><<<grid, THREADS, SMEM_BYTES, st>>>(p, mA, mB);
  return after_launch("attn_gemm_kernel");
}

}  // namespace at
}  // namespace pev

using namespace pev;

// form 0 (NT): out = score buffer [H, Np, Lpad] = scale * A B^T per conformer and head, A / B packed tensors [N, lda / ldb]
// form 1 (NN): out[N, ldo] (columns out_col0 + h hd ..) = P B, A = score buffer; form 2 (TN): out = P^T B
extern "C" int pev_attn_gemm(int32_t form, const float* A, int64_t lda, int32_t a_col0, const float* Bm, int64_t ldb,
                             int32_t b_col0, const int32_t* tile_conf, const int32_t* cu, const int32_t* cup, int32_t m_tiles,
                             int32_t H, int32_t hd, int32_t Lpad, int64_t N, float scale, float* out, int64_t ldo,
                             int32_t out_col0, void* stream) {
  PEV_REQUIRE(A && Bm && tile_conf && cu && cup && out, "null argument");
  PEV_REQUIRE((hd == 64 || hd == 128) && Lpad > 0 && Lpad % 32 == 0 && m_tiles >= 0 && H > 0 && lda % 4 == 0 && ldb % 4 == 0,
              "shape: head dim 64 or 128, Lpad a multiple of 32");
  if (m_tiles == 0 || N == 0) return 0;
  cudaStream_t st = as_stream(stream);
  at::Params p = {};
  p.tile_conf = tile_conf; p.cu = cu; p.cup = cup; p.m_tiles = m_tiles; p.H = H; p.hd = hd; p.Lpad = Lpad;
  p.Np = (int64_t)m_tiles * at::BM; p.a_col0 = a_col0; p.b_col0 = b_col0; p.scale = scale; p.out = out; p.ldo = ldo;
  p.out_col0 = out_col0;
  alignas(64) CUtensorMap mA, mB;
  if (form == at::NT) {
    if (int rc = at::make_map(A, N, lda, lda, at::BM, false, &mA)) return rc;
    if (int rc = at::make_map(Bm, N, ldb, ldb, 256, false, &mB)) return rc;
    return at::launch<at::NT>(p, mA, mB, st);
  }
  PEV_REQUIRE(ldo % 4 == 0, "ldo: a multiple of 4");
  if (form == at::NN) {
    if (int rc = at::make_map(A, (int64_t)H * p.Np, Lpad, Lpad, at::BM, false, &mA)) return rc;
    if (int rc = at::make_map(Bm, N, ldb, ldb, 32, true, &mB)) return rc;
    return at::launch<at::NN>(p, mA, mB, st);
  }
  if (form == at::TN) {
    if (int rc = at::make_map(A, (int64_t)H * p.Np, Lpad, Lpad, 32, true, &mA)) return rc;
    if (int rc = at::make_map(Bm, N, ldb, ldb, 32, true, &mB)) return rc;
    return at::launch<at::TN>(p, mA, mB, st);
  }
  return set_error(1, "pev_attn_gemm: unknown form %d", form);
}

extern "C" int pev_attn_softmax(int32_t backward, float* S, float* G, const int32_t* tile_conf, const int32_t* cu,
                                const int32_t* cup, int32_t m_tiles, int32_t H, int32_t Lpad, float p_drop, uint32_t seed,
                                void* stream) {
  PEV_REQUIRE(S && tile_conf && cu && cup && (!backward || G) && p_drop >= 0.f && p_drop < 1.f, "bad argument");
  if (m_tiles == 0) return 0;
  const int64_t Np = (int64_t)m_tiles * at::BM, rows = (int64_t)H * Np;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  cudaStream_t st = as_stream(stream);
  if (backward) at::softmax_bwd_kernel<<<grid, 256, 0, st>>>(S, G, tile_conf, cu, cup, Np, H, Lpad, p_drop, seed);
  else at::softmax_fwd_kernel<<<grid, 256, 0, st>>>(S, G, tile_conf, cu, cup, Np, H, Lpad, p_drop, seed);
  return after_launch("attn_softmax_kernel");
}

// Scores with a fused epilogue (form NT): mode 1 -- out = softmax probabilities (requires Lpad <= 256), out2 = dropout-kept
// probabilities / (1 - p_drop) or NULL; mode 2 -- A = dO, B = V: out = dS = P (keep dP / (1 - p_drop) - delta).
extern "C" int pev_attn_scores(int32_t mode, const float* A, int64_t lda, int32_t a_col0, const float* Bm, int64_t ldb,
                               int32_t b_col0, const int32_t* tile_conf, const int32_t* cu, const int32_t* cup, int32_t m_tiles,
                               int32_t H, int32_t hd, int32_t Lpad, int64_t N, float scale, float* out, float* out2, const float* P,
                               const float* delta, float p_drop, uint32_t seed, void* stream) {
  PEV_REQUIRE(A && Bm && tile_conf && cu && cup && out && (mode == 1 || mode == 2), "bad argument");
  PEV_REQUIRE((hd == 64 || hd == 128) && Lpad > 0 && Lpad % 32 == 0 && m_tiles >= 0 && H > 0 && lda % 4 == 0 && ldb % 4 == 0,
              "shape: head dim 64 or 128, Lpad a multiple of 32");
  PEV_REQUIRE(mode != 1 || Lpad <= 256, "the fused softmax needs the whole row in one key tile (Lpad <= 256)");
  PEV_REQUIRE(mode != 2 || (P && delta), "mode 2 needs P and delta");
  PEV_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "p_drop in [0, 1)");
  if (m_tiles == 0 || N == 0) return 0;
  at::Params p = {};
  p.tile_conf = tile_conf; p.cu = cu; p.cup = cup; p.m_tiles = m_tiles; p.H = H; p.hd = hd; p.Lpad = Lpad;
  p.Np = (int64_t)m_tiles * at::BM; p.a_col0 = a_col0; p.b_col0 = b_col0; p.scale = scale; p.out = out;
  p.mode = mode; p.out2 = out2; p.P = P; p.delta = delta;
  p.drop_thresh = p_drop > 0.f ? (uint32_t)(p_drop * 4294967296.0) : 0u;
  p.drop_seed = seed;
  p.drop_scale = 1.0f / (1.0f - p_drop);
  alignas(64) CUtensorMap mA, mB;
  if (int rc = at::make_map(A, N, lda, lda, at::BM, false, &mA)) return rc;
  if (int rc = at::make_map(Bm, N, ldb, ldb, 256, false, &mB)) return rc;
  return at::launch<at::NT>(p, mA, mB, as_stream(stream));
}

extern "C" int pev_attn_delta(const float* dO, const float* O, int64_t N, int32_t H, int32_t hd, const int32_t* row_pad,
                              int64_t Np, float* delta, void* stream) {
  PEV_REQUIRE(dO && O && row_pad && delta && N >= 0 && H > 0 && hd > 0, "bad argument");
  if (N == 0) return 0;
  at::attn_delta_kernel<<<(unsigned)((N + 7) / 8), 256, 0, as_stream(stream)>>>(dO, O, N, H, hd, row_pad, Np, delta);
  return after_launch("attn_delta_kernel");
}
