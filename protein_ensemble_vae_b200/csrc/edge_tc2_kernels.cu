// K1, bf16 tensor-core form, second generation ("v2"): the EGNN edge MLP (models/en_gnn_decoder.py:60-79) and its
// backward on tcgen05, organised so that the CUDA-core roles stay near the MUFU floor (one tanh per activation):
//
//   * half domain: every pre-activation is carried as h = z/2 (the 1/2 is folded into the packed weights, the biases
//     and the node-level A|B projections), so silu(z) = h + h tanh(h) and silu'(z) = (1 + r)/2 need no extra scaling;
//   * orientation by epilogue: a GEMM whose epilogue reduces over EDGES (segment sums agg / column sums) is issued
//     transposed, D^T[feature, edge] = W . X^T, so that a TMEM lane is a feature and the 32 registers of a tcgen05.ld
//     are 32 consecutive edges -- segment sums become in-thread adds, stores and node loads are coalesced across lanes;
//     a GEMM whose epilogue reduces over FEATURES (w = t.w6, gd2 = gu.wd) keeps lane = edge;
//   * per-edge tensors that are produced by a feature-lane epilogue are kept in HBM as "tile images": per 128-edge tile
//     a 64 KB block [fq 4][eh 2][fg 8][r 8][128 B] (feature f = 64 fq + 8 fg + r, edge e = 64 eh + 8 c + i stored at
//     16-byte chunk c ^ r) -- exactly the SWIZZLE_128B shared-memory image of the tile, which the tensor core reads
//     either as an MN-major operand (M = edges) or as a K-major operand (K = edges), so consumers copy it verbatim.
//
//   fwd1  hu = Ah_i + Bh_j + wdh d2 ; a = silu  --GEMM W2h (transposed)-->  hv (+b2h), m = silu(hv) -> HBM tile images
//         (hv only when a backward pass follows) ; agg[row] += m (in-thread segment sums, one RED per segment per feature)
//   fwd2  m (TMA, A operand MN-major)  --GEMM W5h-->  hs (+b5h) [-> HBM, training], t = silu, w[e] = t.w6 + b6
//   bwd2  ghs = gw w6 (1 + r(hs))  --GEMM W5h^T (transposed)-->  ghv = (. + gagg[row]) (1 + r(hv)) -> HBM tile image
//         (hv arrives by TMA into the warp's staging buffer and is overwritten in place) ; db2 += sum_e ghv
//   bwd1  ghv (TMA, A operand MN-major)  --GEMM W2h^T-->  ghu = . (1 + r(hu)), hu rebuilt from ABh / d2 -> HBM rows,
//         gd2[e] = ghu . wdh
//   wgrad5 / wgrad2  dW5 = ghs^T m, dW2 = ghv^T a : split-K over edges, the whole 256 x 256 fp32 matrix accumulated
//         in TMEM over a CTA's tiles, per-CTA partials + fixed-order reduction
//   sums  row / column segment sums of ghu (gA | gB) and sum_e d2 ghu (gwd), one warp per node
#include <cuda.h>          // CUtensorMap types only: the encoder is resolved at run time (no libcuda link dependency)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "tc_common.cuh"
#include "tc_edge_common.cuh"

namespace pev {
namespace tc2 {
using namespace tcx;
using namespace tce;

constexpr int H = 256;
constexpr int TILE_M = 128;             // edges per tile
constexpr int KCHUNK = 64;              // bf16 per 128-byte swizzle row
constexpr int NUM_KCHUNKS = H / KCHUNK;
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 4;           // operand ring: up to 4 x 16 KB K-chunks in flight
constexpr int STAGE_BYTES = TILE_M * KCHUNK * 2;
constexpr int W_BYTES = H * H * 2;
constexpr int TILE_IMG_BYTES = TILE_M * H * 2;   // 64 KB tile image
// Warp roles, aligned to warpgroups so that setmaxnreg can move registers between them:
//   warps 0..7 epilogue (TMEM lane quarter = warp % 4), warp 8 MMA issue (+ 9..11 register donors),
//   warps 12..27 producers.  896 threads x 72 registers at launch; then MMA group 40, producers 64, epilogue 104.
constexpr int NUM_EPI_WARPS = 8;
constexpr int MMA_WARP = 8;
constexpr int PROD_WARP0 = 12;
constexpr int NUM_PROD_WARPS = 16;
constexpr int NUM_THREADS = 32 * (PROD_WARP0 + NUM_PROD_WARPS);   // 896
constexpr int NUM_PROD_THREADS = 32 * NUM_PROD_WARPS;             // 512
constexpr int NUM_EPI_THREADS = 32 * NUM_EPI_WARPS;
constexpr int REGS_MMA_WG = 40;
constexpr int REGS_PROD = 64;
constexpr int REGS_EPI = 104;
constexpr int TMEM_COLS = 512;

// Dynamic shared memory: resident weight image | operand ring | 4 x 256 floats | per-warp staging | barriers.
template <int STAGES, int VEC_FLOATS, int STG_BYTES>
struct SmemL {
  static constexpr int NSTAGE = STAGES;
  static constexpr int W_OFF = 0;
  static constexpr int A_OFF = W_BYTES;
  static constexpr int VEC_OFF = A_OFF + STAGES * STAGE_BYTES;          // per-feature vectors (bias, wd, w6)
  static constexpr int STG_OFF = VEC_OFF + VEC_FLOATS * 4;              // per-warp staging of tile-image stores
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256;
  static constexpr int BYTES = TOTAL + 1024;                            // slack for manual 1024-byte alignment
  static_assert(BYTES <= 232448, "shared memory budget");
};
using SmemT = SmemL<2, 256, 8 * 8192>;   // fwd1 (training): 2-stage ring + 8 x (4 KB hv rows | 4 KB m rows) image staging
using SmemTI = SmemL<4, 256, 8 * 4096>;  // fwd1 (inference, hv not written): 4-stage ring + 8 x 4 KB m rows
using SmemB2 = SmemL<2, 256, 8 * 8192>;  // bwd2: 2-stage ring + 8 x 2 x 4 KB image buffers (hv in by TMA, ghv out in place)
using SmemB1 = SmemL<2, 256, 16 * 4096>;  // bwd1: 2-stage ring (TMA-fed) + 16 x 2 x 2 KB row-box staging for the store of ghu

struct Bars {
  uint64_t* full;     // [NUM_STAGES] producers -> MMA
  uint64_t* empty;    // [NUM_STAGES] MMA -> producers
  uint64_t* tfull;    // [2] MMA -> epilogue
  uint64_t* tempty;   // [2] epilogue -> MMA
  uint64_t* w;        // weight image landed
  uint32_t* tmem_slot;
};
__device__ __forceinline__ Bars make_bars(uint8_t* bar_base) {
  uint64_t* b = reinterpret_cast<uint64_t*>(bar_base);
  return Bars{b, b + MAX_STAGES, b + 2 * MAX_STAGES, b + 2 * MAX_STAGES + 2, b + 2 * MAX_STAGES + 4,
              reinterpret_cast<uint32_t*>(b + 2 * MAX_STAGES + 5)};
}
__device__ __forceinline__ void init_bars(const Bars& B, int full_count) {
  for (int s = 0; s < MAX_STAGES; ++s) {
    mbar_init(&B.full[s], full_count);
    mbar_init(&B.empty[s], 1);
  }
  for (int a = 0; a < 2; ++a) {
    mbar_init(&B.tfull[a], 1);
    mbar_init(&B.tempty[a], NUM_EPI_THREADS);
  }
  mbar_init(B.w, 1);
  fence_barrier_init();
}
__device__ __forceinline__ void load_weight_image(uint8_t* sW, const void* Wp, uint64_t* bar) {
  mbar_arrive_expect_tx(bar, W_BYTES);
  for (int i = 0; i < W_BYTES / 16384; ++i)
    bulk_g2s(sW + i * 16384, reinterpret_cast<const uint8_t*>(Wp) + i * 16384, 16384, bar);
  mbar_wait(bar, 0);
}

// Common prologue: barriers, TMEM, role register budgets.  Returns the TMEM base address.
#define PEV_TC2_PROLOGUE(SM, FULL_COUNT)                                                      \
  constexpr int NUM_STAGES = SM::NSTAGE;                                                      \
  extern __shared__ __align__(1024) uint8_t smem_raw[];                                       \
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);                \
  uint8_t* sW = smem + SM::W_OFF;                                                             \
  uint8_t* sA = smem + SM::A_OFF;                                                             \
  float* sVec = reinterpret_cast<float*>(smem + SM::VEC_OFF);                                 \
  const Bars B = make_bars(smem + SM::BAR_OFF);                                               \
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;                                 \
  if (threadIdx.x == 0) init_bars(B, FULL_COUNT);                                             \
  if (warp == MMA_WARP) tmem_alloc(B.tmem_slot, TMEM_COLS);

#define PEV_TC2_SYNC_ROLES()                                                                  \
  tc_fence_before();                                                                          \
  __syncthreads();                                                                            \
  tc_fence_after();                                                                           \
  const uint32_t tmem_base = *B.tmem_slot;

#define PEV_TC2_EPILOGUE()                                                                    \
  tc_fence_before();                                                                          \
  __syncthreads();                                                                            \
  if (warp == MMA_WARP) {                                                                     \
    tc_fence_after();                                                                         \
    tmem_dealloc(tmem_base, TMEM_COLS);                                                       \
  }

// =================================================================================================== fwd1
struct Fwd1Params {
  const __half* ABh;           // [N,512] fp16, half domain: 0.5 (h Wa^T + b1) | 0.5 h Wb^T
  const float* d2;            // [E] squared edge lengths
  const int32_t* row;         // [E]
  const int32_t* col;         // [E]
  const float* wd;            // [256] (full domain; halved on load)
  const float* b2;            // [256] (full domain; halved on load)
  const void* W2hp;           // packed image of 0.5 W2
  uint8_t* hvT;               // [num_tiles] x 64 KB tile images of hv = v/2 (bf16); null when no backward follows
  uint8_t* mT;                // [num_tiles] x 64 KB tile images of m = silu(v) (bf16): the operand of fwd2 / wgrad5
  float* agg;                 // [N,256] (+=)
  int64_t E;
  int num_tiles;
  int dbg;                    // PEV_TC2_DEBUG bit mask (profiling experiments only; 0 in production)
};

template <int DBG, bool TRAIN>   // TRAIN: hv is written too (p.hvT != null)
__global__ void __launch_bounds__(NUM_THREADS, 1) fwd1_kernel(const Fwd1Params p) {
  const int dbg = DBG ? p.dbg : 0;
  using SM = std::conditional_t<TRAIN, SmemT, SmemTI>;
  constexpr int STG_PER_WARP = TRAIN ? 8192 : 4096;
  constexpr int M_STG_OFF = TRAIN ? 4096 : 0;
  PEV_TC2_PROLOGUE(SM, NUM_PROD_THREADS)
  float* sWd = sVec;
  for (int k = threadIdx.x; k < H; k += NUM_THREADS) sWd[vec_slot(k)] = 0.5f * p.wd[k];
  PEV_TC2_SYNC_ROLES()

  if (warp >= MMA_WARP && warp < PROD_WARP0) {
    // ------------------------------------------------------------------ MMA issue: D^T[f, e] = W2h[f, :] . a[e, :]
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA_WG));
    if (warp == MMA_WARP && lane == 0) {
      load_weight_image(sW, p.W2hp, B.w);
      constexpr uint32_t IDESC = idesc_bf16(128, 128, false, false);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&B.tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * H;
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&B.full[stage], phase);
          tc_fence_after();
          const uint32_t x_base = smem_u32(sA + stage * STAGE_BYTES);           // activations: [128 e][64 k]
          const uint32_t w_base = smem_u32(sW + kc * (H * KCHUNK * 2));          // weights:     [256 f][64 k]
#pragma unroll
          for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)
              if (!(dbg & 32))
                umma_bf16(d0 + mh * 128, desc_kmajor(w_base + mh * 16384 + ks * UMMA_K * 2),
                          desc_kmajor(x_base + ks * UMMA_K * 2), IDESC, (kc | ks) != 0 ? 1u : 0u);
          umma_commit(&B.empty[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&B.tfull[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= PROD_WARP0) {
    // ------------------------------------------------------------------ producers: a = silu(hu) -> K-major ring
    // Warps are independent (no block barrier): warp pw owns rows 8 pw .. 8 pw + 7 of every tile; lane -> row
    // (lane >> 3) + 4 i (i < 2), 16-byte column chunk lane & 7.  Row metadata (row, col, d2) is fetched one tile ahead,
    // the A|B operand chunks one K-chunk ahead.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PROD));
    const int pw = warp - PROD_WARP0;
    const int chunk = lane & 7;
    constexpr int RPT = 2;
    const int r0 = pw * 8 + (lane >> 3);
    int stage = 0;
    uint32_t phase = 0;
    // Row metadata is kept lane-distributed (lane l: edge 8 pw + (l & 7) of the current / the next tile, three registers
    // each) and handed to its users by shuffles at every use: kept per user lane it was spilled right after the loads,
    // which parked every producer warp for a full global-load latency once per tile.
    struct Meta { int r, c; float d; };
    auto load_meta = [&](int tile, Meta& m) {
      int64_t e = (int64_t)tile * TILE_M + pw * 8 + (lane & 7);
      e = e < p.E ? e : p.E - 1;                   // tail rows recompute the last edge; the epilogue drops them
      m.r = __ldg(p.row + e);
      m.c = __ldg(p.col + e);
      m.d = __ldg(p.d2 + e);
    };
    uint4 pfA[2][RPT], pfB[2][RPT];
    auto issue = [&](const Meta& m, int kc, int slot) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int nr = __shfl_sync(0xffffffffu, m.r, (lane >> 3) + 4 * i);
        const int nc = __shfl_sync(0xffffffffu, m.c, (lane >> 3) + 4 * i);
        if (dbg & 1) { pfA[slot][i] = pfB[slot][i] = make_uint4(0u, 0u, 0u, 0u); continue; }
        pfA[slot][i] = __ldg(reinterpret_cast<const uint4*>(p.ABh + (int64_t)nr * 2 * H + kc * KCHUNK + chunk * 8));
        pfB[slot][i] = __ldg(reinterpret_cast<const uint4*>(p.ABh + (int64_t)nc * 2 * H + H + kc * KCHUNK + chunk * 8));
      }
    };
    Meta mc, mn;
    load_meta(blockIdx.x, mc);
    mn = mc;
    issue(mc, 0, 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int next_tile = tile + gridDim.x;
      if (dbg & 64) {                                  // role ablation: ring handshake only
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&B.empty[stage], phase ^ 1);
          fence_proxy_async();
          mbar_arrive(&B.full[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (next_tile < p.num_tiles) load_meta(next_tile, mn);
      float d2r[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) d2r[i] = __shfl_sync(0xffffffffu, mc.d, (lane >> 3) + 4 * i);
#pragma unroll
      for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
        if (kc + 1 < NUM_KCHUNKS) issue(mc, kc + 1, (kc + 1) & 1);
        else if (next_tile < p.num_tiles) issue(mn, 0, 0);
        const int k0 = kc * KCHUNK + chunk * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(sWd + vec_slot(k0)), w1 = *reinterpret_cast<const float4*>(sWd + vec_slot(k0 + 4));
        const float wd8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint4 out[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          float s8[8], o8[8];
          add_f16x8_to_f32(pfA[kc & 1][i], pfB[kc & 1][i], s8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float hu = fmaf(wd8[j], d2r[i], s8[j]);
            o8[j] = (dbg & 16) ? hu : silu_h(hu);
          }
          out[i] = pack8(o8);
        }
        mbar_wait(&B.empty[stage], phase ^ 1);
        uint8_t* st = sA + stage * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) *reinterpret_cast<uint4*>(st + sw128_offset(r0 + 4 * i, chunk)) = out[i];
        fence_proxy_async();
        mbar_arrive(&B.full[stage]);
        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
      }
      mc = mn;
    }
  } else {
    // ------------------------------------------------------------------ epilogue: lane = feature, registers = edges
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    const int q = warp & 3, mh = warp >> 2;
    const int f = mh * 128 + q * 32 + lane;
    const float bias = 0.5f * __ldg(p.b2 + f);
    float* aggcol = p.agg + f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * 128);
    // tile-image offset of this warp's 32 rows (features f - lane .. + 31), edge half 0: 4 KB contiguous
    const uint32_t img_blk = (uint32_t)((f >> 6) * 16384 + (q & 1) * 4096);
    const int sw = f & 7;
    uint8_t* stg = smem + SM::STG_OFF + warp * STG_PER_WARP; // 4 KB buffers: [hv rows |] m rows
    auto flush = [&](int r, float s) {
      if (r >= 0 && !(dbg & 4)) atomicAdd(aggcol + (int64_t)r * H, s);
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int64_t e0 = (int64_t)tile * TILE_M;
      int myrow[4];
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const int64_t e = e0 + cb * 32 + lane;
        myrow[cb] = e < p.E ? __ldg(p.row + e) : -1;
      }
      mbar_wait(&B.tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      if (dbg & 128) {                                 // role ablation: accumulator handshake only
        tc_fence_before();
        mbar_arrive(&B.tempty[acc]);
        continue;
      }
      int cur = __shfl_sync(0xffffffffu, myrow[0], 0);
      float seg = 0.f;
      uint32_t raw[32];
      tmem_ld32_issue(lane_addr + acc * H, raw);
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        tmem_wait();
        float val[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) val[j] = __uint_as_float(raw[j]) + bias;
        if (cb + 1 < 4) tmem_ld32_issue(lane_addr + acc * H + (cb + 1) * 32, raw);
        else {
          tc_fence_before();
          mbar_arrive(&B.tempty[acc]);            // accumulator stage fully read
        }
        // hv and m = silu(2 hv) -> HBM tile images (bf16).  This warp's 32 feature rows x 64 edges are 4 KB contiguous
        // in an image: the rows are staged in shared memory (swizzled chunk positions: conflict-free) and written by
        // one TMA bulk copy per edge half and image -- full lines, no per-lane global stores.
        float m[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) m[j] = (dbg & 8) ? val[j] : silu_h(val[j]);
        if (!(dbg & 2)) {
          if ((cb & 1) == 0) {
            if (lane == 0) bulk_wait_read();          // the previous bulk copies have read the staging buffers
            __syncwarp();
          }
          uint8_t* srow = stg + lane * 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t pos = (uint32_t)((((cb & 1) * 4 + k) ^ sw) << 4);
            if (TRAIN) {
              const float o8[8] = {val[8 * k], val[8 * k + 1], val[8 * k + 2], val[8 * k + 3],
                                   val[8 * k + 4], val[8 * k + 5], val[8 * k + 6], val[8 * k + 7]};
              *reinterpret_cast<uint4*>(srow + pos) = pack8(o8);
            }
            const float q8[8] = {m[8 * k], m[8 * k + 1], m[8 * k + 2], m[8 * k + 3],
                                 m[8 * k + 4], m[8 * k + 5], m[8 * k + 6], m[8 * k + 7]};
            *reinterpret_cast<uint4*>(srow + M_STG_OFF + pos) = pack8(q8);
          }
          if (cb & 1) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int64_t off = (int64_t)tile * TILE_IMG_BYTES + img_blk + (cb >> 1) * 8192;
              if (TRAIN) bulk_s2g(p.hvT + off, stg, 4096);
              bulk_s2g(p.mT + off, stg + M_STG_OFF, 4096);
            }
          }
        }
        // segment sums over the 32 edges (uniform control flow: every lane sees the same edges)
        int prev = __shfl_up_sync(0xffffffffu, myrow[cb], 1);
        if (lane == 0) prev = cur;
        uint32_t bm = __ballot_sync(0xffffffffu, myrow[cb] != prev);
        if (bm == 0u) {
          seg += sum32(m);
        } else {
          int lo = 0;
          while (bm) {
            const int jb = __ffs(bm) - 1;
            bm &= bm - 1u;
            flush(cur, seg + masked_sum32(m, ((1u << jb) - 1u) & ~((1u << lo) - 1u)));
            seg = 0.f;
            cur = __shfl_sync(0xffffffffu, myrow[cb], jb);
            lo = jb;
          }
          seg = masked_sum32(m, ~((1u << lo) - 1u));
        }
      }
      flush(cur, seg);
    }
    if (lane == 0) bulk_wait_all();
  }
  PEV_TC2_EPILOGUE()
}

// =================================================================================================== fwd2
//   m (tile images, TMA -> MN-major A operand)  --GEMM (W5/2)-->  hs (+b5h) [-> HBM rows, training], t = silu,
//   w[e] = t . w6 + b6.   Block: 16 epilogue warps (lane = edge; TMEM lane quarter = warp % 4, column quarter =
//   warp / 4), warp 16 = MMA issue, warp 17 = TMA loads.  No CUDA-core producer: fwd1 already wrote m.
constexpr int F2_EPI_WARPS = 16;
constexpr int F2_MMA_WARP = 16;
constexpr int F2_TMA_WARP = 17;
constexpr int F2_THREADS = 32 * (F2_EPI_WARPS + 4);     // 640 -> 96 registers at launch
constexpr int F2_REGS_EPI = 104;
using SmemF2 = SmemL<3, 512 + 768, 16 * 2048>;          // 3-stage ring + b5h | w6 | 2 x 3 x 128 dot partials + 16 x 2 KB row-box staging

struct Fwd2Params {
  const uint8_t* mT;      // tile images of m = silu(v)
  const void* W5hp;       // packed image of 0.5 W5
  const float* b5;        // [256] full domain
  const float* w6;        // [256]
  const float* b6;        // [1]
  __nv_bfloat16* hs;      // [E,256] hs = s/2 (training) or null
  float* w;               // [E]
  int64_t E;
  int num_tiles;
  int dbg;
};

// named barrier of the four warps that share a TMEM lane quarter (ids 1..4; 0 is __syncthreads)
__device__ __forceinline__ void quarter_barrier(int q) { asm volatile("bar.sync %0, 128;" ::"r"(q + 1) : "memory"); }

template <int DBG, bool TRAIN>   // TRAIN: hs is written (p.hs != null)
__global__ void __launch_bounds__(F2_THREADS, 1) fwd2_kernel(const Fwd2Params p, const __grid_constant__ CUtensorMap hs_map) {
  const int dbg = DBG ? p.dbg : 0;
  constexpr int NUM_STAGES = SmemF2::NSTAGE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem + SmemF2::W_OFF;
  uint8_t* sA = smem + SmemF2::A_OFF;
  float* sBias = reinterpret_cast<float*>(smem + SmemF2::VEC_OFF);
  float* sW6 = sBias + H;
  float* sPart = sW6 + H;          // [tile parity 2][column quarter 1..3][128 edges]: partial dots, combined in fixed order
  const Bars B = make_bars(smem + SmemF2::BAR_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < H; k += F2_THREADS) {
    sBias[k] = 0.5f * p.b5[k];
    sW6[k] = p.w6[k];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(&B.full[s], 1);
      mbar_init(&B.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&B.tfull[a], 1);
      mbar_init(&B.tempty[a], 32 * F2_EPI_WARPS);
    }
    mbar_init(B.w, 1);
    fence_barrier_init();
  }
  if (warp == F2_MMA_WARP) tmem_alloc(B.tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *B.tmem_slot;

  if (warp >= F2_EPI_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA_WG));
    if (warp == F2_MMA_WARP && lane == 0) {
      // ---------------------------------------------------------------- MMA issue: D[e, n] = m[e, :] . W5h[n, :]
      load_weight_image(sW, p.W5hp, B.w);
      constexpr uint32_t IDESC = idesc_bf16(128, 256, true, false);             // A (activations) MN-major
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&B.tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * H;
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&B.full[stage], phase);
          tc_fence_after();
          const uint32_t x_base = smem_u32(sA + stage * STAGE_BYTES);           // [eh 2][fg 8][r 8][128 B]
          const uint32_t w_base = smem_u32(sW + kc * (H * KCHUNK * 2));
#pragma unroll
          for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
            if (!(dbg & 32))
              umma_bf16(d0, desc_mnmajor(x_base + ks * 2048, 8192, 1024), desc_kmajor(w_base + ks * UMMA_K * 2), IDESC,
                        (kc | ks) != 0 ? 1u : 0u);
          umma_commit(&B.empty[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&B.tfull[acc]);
      }
    } else if (warp == F2_TMA_WARP && lane == 0) {
      // ---------------------------------------------------------------- TMA: tile image K-chunks -> ring, verbatim
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&B.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&B.full[stage], STAGE_BYTES);
          bulk_g2s(sA + stage * STAGE_BYTES, p.mT + (int64_t)tile * TILE_IMG_BYTES + kc * STAGE_BYTES, STAGE_BYTES,
                   &B.full[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: lane = edge, registers = features
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(F2_REGS_EPI));
    const int q = warp & 3, cq = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
    const float b6 = cq == 0 ? __ldg(p.b6) : 0.f;
    uint8_t* stg = smem + SmemF2::STG_OFF + warp * 2048;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int64_t e = (int64_t)tile * TILE_M + q * 32 + lane;
      const bool valid = e < p.E;
      float dot = 0.f;
      mbar_wait(&B.tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      if (dbg & 128) {
        tc_fence_before();
        mbar_arrive(&B.tempty[acc]);
        continue;
      }
      uint32_t raw[32];
      tmem_ld32_issue(lane_addr + acc * H, raw);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int col0 = cq * 64 + b * 32;
        tmem_wait();
        float val[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(sBias + col0 + 4 * j4);
          val[4 * j4] = __uint_as_float(raw[4 * j4]) + bb.x;
          val[4 * j4 + 1] = __uint_as_float(raw[4 * j4 + 1]) + bb.y;
          val[4 * j4 + 2] = __uint_as_float(raw[4 * j4 + 2]) + bb.z;
          val[4 * j4 + 3] = __uint_as_float(raw[4 * j4 + 3]) + bb.w;
        }
        if (b == 0) tmem_ld32_issue(lane_addr + acc * H + 32, raw);
        else {
          tc_fence_before();
          mbar_arrive(&B.tempty[acc]);
        }
        if (TRAIN && !(dbg & 2)) {
          // hs rows -> HBM through a TMA tensor store: the warp's [32 edges x 32 features] box is staged in shared
          // memory (64-byte rows, SWIZZLE_64B chunk positions: conflict-free) and written by the TMA engine, which
          // also clips the rows past E.
          if (lane == 0) bulk_wait_read();
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float o8[8] = {val[8 * k], val[8 * k + 1], val[8 * k + 2], val[8 * k + 3],
                                 val[8 * k + 4], val[8 * k + 5], val[8 * k + 6], val[8 * k + 7]};
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) = pack8(o8);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) tma_store_2d(&hs_map, col0, (int)((int64_t)tile * TILE_M + q * 32), stg);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(sW6 + col0 + 4 * j4);
          dot = fmaf(silu_h(val[4 * j4]), w.x, dot);
          dot = fmaf(silu_h(val[4 * j4 + 1]), w.y, dot);
          dot = fmaf(silu_h(val[4 * j4 + 2]), w.z, dot);
          dot = fmaf(silu_h(val[4 * j4 + 3]), w.w, dot);
        }
      }
      // w[e] = ((q0 + b6) + q1) + (q2 + q3): the four column quarters of an edge live in four warps; the partials meet
      // in shared memory (two buffers by tile parity: a writer reaches tile n + 2 only after the reader left tile n)
      float* part = sPart + (it & 1) * 3 * TILE_M + q * 32 + lane;
      if (cq != 0) part[(cq - 1) * TILE_M] = dot;
      quarter_barrier(q);
      if (cq == 0 && valid) p.w[e] = ((dot + b6) + part[0]) + (part[TILE_M] + part[2 * TILE_M]);
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == F2_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =================================================================================================== bwd2
//   ghs = gw w6 (1 + r(hs))  --GEMM (W5/2) (transposed)-->  gm ; ghv = (gm + gagg[row]) (1 + r(hv)) -> HBM tile image,
//   db2 += sum_e ghv (in-thread).   [d silu(2h)/dh = 1 + r(h)]
struct Bwd2Params {
  const __nv_bfloat16* hs;    // [E,256] hs = s/2
  const float* gw;            // [E] dL/dw
  const float* w6;            // [256]
  const void* W5thp;          // packed image of 0.5 W5^T
  const float* gagg;          // [N,256] dL/dagg
  const int32_t* row;         // [E]
  const uint8_t* hvT;         // tile images of hv
  uint8_t* ghvT;              // tile images of ghv = dL/dhv (out)
  float* db2_partial;         // [gridDim.x][256]: per-CTA sum_e ghv, reduced in fixed order by the launcher
  int64_t E;
  int num_tiles;
  int dbg;
};

template <int DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1) bwd2_kernel(const Bwd2Params p) {
  const int dbg = DBG ? p.dbg : 0;
  PEV_TC2_PROLOGUE(SmemB2, NUM_PROD_THREADS)
  float* sW6 = sVec;
  for (int k = threadIdx.x; k < H; k += NUM_THREADS) sW6[vec_slot(k)] = p.w6[k];
  PEV_TC2_SYNC_ROLES()

  if (warp >= MMA_WARP && warp < PROD_WARP0) {
    // ------------------------------------------------------------------ MMA issue: D^T[f, e] = W5h^T[f, :] . ghs[e, :]
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA_WG));
    if (warp == MMA_WARP && lane == 0) {
      load_weight_image(sW, p.W5thp, B.w);
      constexpr uint32_t IDESC = idesc_bf16(128, 128, false, false);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&B.tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * H;
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&B.full[stage], phase);
          tc_fence_after();
          const uint32_t x_base = smem_u32(sA + stage * STAGE_BYTES);
          const uint32_t w_base = smem_u32(sW + kc * (H * KCHUNK * 2));
#pragma unroll
          for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
#pragma unroll
            for (int mh = 0; mh < 2; ++mh)
              if (!(dbg & 32))
                umma_bf16(d0 + mh * 128, desc_kmajor(w_base + mh * 16384 + ks * UMMA_K * 2),
                          desc_kmajor(x_base + ks * UMMA_K * 2), IDESC, (kc | ks) != 0 ? 1u : 0u);
          umma_commit(&B.empty[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&B.tfull[acc]);
      }
    }
    __syncwarp();
  } else if (warp >= PROD_WARP0) {
    // ------------------------------------------------------------------ producers: ghs -> K-major ring (rows = edges)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PROD));
    const int pw = warp - PROD_WARP0;
    const int chunk = lane & 7;
    constexpr int RPT = 2;
    const int r0 = pw * 8 + (lane >> 3);
    int stage = 0;
    uint32_t phase = 0;
    uint4 pf[2][RPT];
    float gwc[RPT], gwn[RPT];
    auto load_meta = [&](int tile, float (&g)[RPT]) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int64_t e = (int64_t)tile * TILE_M + r0 + 4 * i;
        g[i] = e < p.E ? __ldg(p.gw + e) : 0.f;          // tail rows contribute exact zeros
      }
    };
    auto issue = [&](int tile, int kc, int slot) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        int64_t e = (int64_t)tile * TILE_M + r0 + 4 * i;
        e = e < p.E ? e : p.E - 1;
        pf[slot][i] = (dbg & 1) ? make_uint4(0u, 0u, 0u, 0u)
                                : __ldg(reinterpret_cast<const uint4*>(p.hs + e * H + kc * KCHUNK + chunk * 8));
      }
    };
    load_meta(blockIdx.x, gwc);
    issue(blockIdx.x, 0, 0);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int next_tile = tile + gridDim.x;
      if (dbg & 64) {
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&B.empty[stage], phase ^ 1);
          fence_proxy_async();
          mbar_arrive(&B.full[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (next_tile < p.num_tiles) load_meta(next_tile, gwn);
#pragma unroll
      for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
        if (kc + 1 < NUM_KCHUNKS) issue(tile, kc + 1, (kc + 1) & 1);
        else if (next_tile < p.num_tiles) issue(next_tile, 0, 0);
        const int k0 = kc * KCHUNK + chunk * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(sW6 + vec_slot(k0)), w1 = *reinterpret_cast<const float4*>(sW6 + vec_slot(k0 + 4));
        const float w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint4 out[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          float h8[8], o8[8];
          unpack8(pf[kc & 1][i], h8);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float2 w2 = make_float2(w8[j], w8[j + 1]);
            const float2 o = fmul2(splat2(gwc[i]), ffma2(w2, silu_grad_r2(make_float2(h8[j], h8[j + 1])), w2));
            o8[j] = o.x;
            o8[j + 1] = o.y;
          }
          out[i] = pack8(o8);
        }
        mbar_wait_backoff<64>(&B.empty[stage], phase ^ 1);
        uint8_t* st = sA + stage * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) *reinterpret_cast<uint4*>(st + sw128_offset(r0 + 4 * i, chunk)) = out[i];
        fence_proxy_async();
        mbar_arrive(&B.full[stage]);
        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) gwc[i] = gwn[i];
    }
  } else {
    // ------------------------------------------------------------------ epilogue: lane = feature, registers = edges
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
    const int q = warp & 3, mh = warp >> 2;
    const int f = mh * 128 + q * 32 + lane;
    const float* gaggcol = p.gagg + f;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * 128);
    const int sw = f & 7;
    float2 dbacc2 = make_float2(0.f, 0.f);
    // The CTA's tiles are walked as a flat sequence of 32-edge batches (4 per tile), each in two steps of 16 edges
    // (one 32-byte pair of the thread's image row, 16 TMEM columns).  Row ids are fetched two batches ahead, the
    // gagg values of a batch's first / last row one batch ahead.
    // hv arrives by TMA: this warp's 32 feature rows x 64 edges ("half" h of the flat sequence: tile h / 2, edge half
    // h & 1) are 4 KB contiguous in the image and land in one of the warp's two 4 KB buffers; ghv overwrites them in
    // place (every thread rewrites exactly the chunks it read) and leaves by one TMA bulk store per half.  A buffer is
    // refilled one and a half batches before it is needed, after its store has been read out.  No per-lane global
    // loads of hv, no registers held across the prefetch distance.
    const int nb = 4 * ((p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);
    const int nh = nb >> 1;
    auto tile_of = [&](int bi) { return (int)blockIdx.x + (bi >> 2) * (int)gridDim.x; };
    auto batch_row = [&](int bi) -> int {
      const int64_t e = (int64_t)tile_of(bi) * TILE_M + (bi & 3) * 32 + lane;
      return e < p.E ? __ldg(p.row + e) : -1;
    };
    auto batch_ga = [&](int myrow, float& gF, float& gL) {
      const int rF = __shfl_sync(0xffffffffu, myrow, 0), rL = __shfl_sync(0xffffffffu, myrow, 31);
      gF = rF >= 0 ? __ldg(gaggcol + (int64_t)rF * H) : 0.f;
      gL = rL >= 0 ? __ldg(gaggcol + (int64_t)rL * H) : 0.f;
    };
    uint8_t* buf0 = smem + SmemB2::STG_OFF + warp * 8192;
    uint64_t* hvbar = reinterpret_cast<uint64_t*>(smem + SmemB2::BAR_OFF + 128) + warp * 2;
    const uint32_t img_blk = (uint32_t)((f >> 6) * 16384 + (q & 1) * 4096);
    auto arm = [&](int h) {                        // lane 0: fetch half h into buffer h & 1
      uint64_t* bar = hvbar + (h & 1);
      mbar_arrive_expect_tx(bar, 4096);
      bulk_g2s_stream(buf0 + (h & 1) * 4096, p.hvT + (int64_t)tile_of(2 * h) * TILE_IMG_BYTES + img_blk + (h & 1) * 8192, 4096, bar);
    };
    if (lane == 0) {
      mbar_init(hvbar, 1);
      mbar_init(hvbar + 1, 1);
      fence_barrier_init();
      fence_proxy_async();
      if (!(dbg & 4)) {
        if (nh > 0) arm(0);
        if (nh > 1) arm(1);
      }
    }
    __syncwarp();
    int row_c = batch_row(0), row_n = nb > 1 ? batch_row(1) : -1;
    float gaF, gaL, gaFn = 0.f, gaLn = 0.f;
    batch_ga(row_c, gaF, gaL);
#pragma unroll 1
    for (int bi = 0; bi < nb; ++bi) {
      const int cb = bi & 3, it = bi >> 2, acc = it & 1, h = bi >> 1;
      const int tile = tile_of(bi);
      const int row_nn = bi + 2 < nb ? batch_row(bi + 2) : -1;
      if (bi + 1 < nb) batch_ga(row_n, gaFn, gaLn);
      if (cb == 0) {
        mbar_wait(&B.tfull[acc], (it >> 1) & 1);
        tc_fence_after();
      }
      if (dbg & 128) {
        if (cb == 3) {
          tc_fence_before();
          mbar_arrive(&B.tempty[acc]);
        }
        row_c = row_n; row_n = row_nn; gaF = gaFn; gaL = gaLn;
        continue;
      }
      // segments of the 32-edge batch (uniform control flow: every lane sees the same edges)
      const int prev = __shfl_up_sync(0xffffffffu, row_c, 1);
      const uint32_t bm = __ballot_sync(0xffffffffu, lane != 0 && row_c != prev);
      const bool simple = (bm & (bm - 1u)) == 0u;  // at most one boundary in the batch
      const uint32_t low = bm - 1u;                // edges before the boundary (all 32 when there is none)
      float ga_run = gaF;                          // several boundaries: running value / row
      int r_run = __shfl_sync(0xffffffffu, row_c, 0);
      uint8_t* my = buf0 + (h & 1) * 4096 + lane * 128;
      if ((cb & 1) == 0) {
        if (!(dbg & 4)) {
          mbar_wait(hvbar + (h & 1), (h >> 1) & 1);   // this half of hv has landed
        } else if (!(dbg & 2)) {
          if (lane == 0) bulk_wait_read();
          __syncwarp();
        }
      }
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        uint32_t raw[16];
        tmem_ld16_issue(lane_addr + acc * H + cb * 32 + pr * 16, raw);
        const int c = (cb & 1) * 4 + 2 * pr;
        uint4* p0 = reinterpret_cast<uint4*>(my + ((c ^ sw) << 4));
        uint4* p1 = reinterpret_cast<uint4*>(my + (((c + 1) ^ sw) << 4));
        // r(hv) of the 16 edges while the TMEM load is in flight
        float2 r[8];
        {
          float h8[8];
          unpack8(*p0, h8);
#pragma unroll
          for (int j = 0; j < 4; ++j) r[j] = silu_grad_r2(make_float2(h8[2 * j], h8[2 * j + 1]));
          unpack8(*p1, h8);
#pragma unroll
          for (int j = 0; j < 4; ++j) r[4 + j] = silu_grad_r2(make_float2(h8[2 * j], h8[2 * j + 1]));
        }
        tmem_wait();
        if (cb == 3 && pr == 1) {
          tc_fence_before();
          mbar_arrive(&B.tempty[acc]);
        }
        float gv[16];
        if (simple) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const float2 g = fadd2(make_float2(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1])),
                                   make_float2(((low >> (16 * pr + j)) & 1u) ? gaF : gaL,
                                               ((low >> (16 * pr + j + 1)) & 1u) ? gaF : gaL));
            gv[j] = g.x;
            gv[j + 1] = g.y;
          }
        } else {                                   // several short segments (not a banded graph): edge by edge
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int rj = __shfl_sync(0xffffffffu, row_c, 16 * pr + j);
            if (rj != r_run) {
              r_run = rj;
              ga_run = rj >= 0 ? __ldg(gaggcol + (int64_t)rj * H) : 0.f;
            }
            gv[j] = __uint_as_float(raw[j]) + ga_run;
          }
        }
        {
          float2 g2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 g = make_float2(gv[2 * j], gv[2 * j + 1]);
            g2[j] = ffma2(g, r[j], g);
            gv[2 * j] = g2[j].x;
            gv[2 * j + 1] = g2[j].y;
          }
          const float2 s = fadd2(fadd2(fadd2(g2[0], g2[1]), fadd2(g2[2], g2[3])), fadd2(fadd2(g2[4], g2[5]), fadd2(g2[6], g2[7])));
          dbacc2 = fadd2(dbacc2, s);
        }
        {                                           // ghv over hv, same chunks
          const float lo8[8] = {gv[0], gv[1], gv[2], gv[3], gv[4], gv[5], gv[6], gv[7]};
          const float hi8[8] = {gv[8], gv[9], gv[10], gv[11], gv[12], gv[13], gv[14], gv[15]};
          *p0 = pack8(lo8);
          *p1 = pack8(hi8);
        }
        if (pr == 0 && (cb & 1) == 0 && h >= 1 && h + 1 < nh && !(dbg & 4)) {
          // the other buffer (half h - 1) was handed to its store one batch step ago: refill it with half h + 1
          if (lane == 0) {
            bulk_wait_read();
            arm(h + 1);
          }
        }
      }
      if (cb & 1) {                                 // 32 rows x 64 edges complete: one 4 KB TMA bulk store
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && !(dbg & 2))
          bulk_s2g(p.ghvT + (int64_t)tile * TILE_IMG_BYTES + img_blk + (cb >> 1) * 8192, buf0 + (h & 1) * 4096, 4096);
      }
      row_c = row_n; row_n = row_nn; gaF = gaFn; gaL = gaLn;
    }
    if (lane == 0) bulk_wait_all();
    p.db2_partial[(int64_t)blockIdx.x * H + f] = dbacc2.x + dbacc2.y;
  }
  PEV_TC2_EPILOGUE()
}

// =================================================================================================== bwd1
//   ghv (tile images, TMA -> MN-major A operand)  --GEMM (W2/2)-->  ga ; ghu = ga (1 + r(hu)), hu rebuilt from ABh, d2 ;
//   ghu -> HBM rows, gd2[e] = ghu . (wd/2).   Block: 16 epilogue warps (lane = edge; TMEM lane quarter = warp % 4,
//   column quarter = warp / 4), warp 16 = MMA issue, warp 17 = TMA loads.
constexpr int B1_EPI_WARPS = 16;
constexpr int B1_MMA_WARP = 16;
constexpr int B1_TMA_WARP = 17;
constexpr int B1_THREADS = 32 * (B1_EPI_WARPS + 4);     // 640 -> 96 registers at launch
constexpr int B1_REGS_EPI = 104;

struct Bwd1Params {
  const uint8_t* ghvT;        // tile images of ghv
  const void* W2thp;          // packed image of 0.5 W2^T
  const __half* ABh;           // [N,512] fp16
  const float* d2;            // [E]
  const int32_t* row;         // [E]
  const int32_t* col;         // [E]
  const float* wd;            // [256] (full domain; halved on load)
  __nv_bfloat16* ghu;         // [E,256] (out) dL/dhu
  float* gd2_parts;           // [4][E]: per column quarter, summed in fixed order by sum4_kernel
  int64_t E;
  int num_tiles;
  int dbg;
};

template <int DBG>
__global__ void __launch_bounds__(B1_THREADS, 1) bwd1_kernel(const Bwd1Params p, const __grid_constant__ CUtensorMap ghu_map) {
  const int dbg = DBG ? p.dbg : 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int NUM_STAGES = SmemB1::NSTAGE;
  uint8_t* sW = smem + SmemB1::W_OFF;
  uint8_t* sA = smem + SmemB1::A_OFF;
  float* sWd = reinterpret_cast<float*>(smem + SmemB1::VEC_OFF);
  const Bars B = make_bars(smem + SmemB1::BAR_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < H; k += B1_THREADS) sWd[k] = 0.5f * p.wd[k];
  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(&B.full[s], 1);
      mbar_init(&B.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&B.tfull[a], 1);
      mbar_init(&B.tempty[a], 32 * B1_EPI_WARPS);
    }
    mbar_init(B.w, 1);
    fence_barrier_init();
  }
  if (warp == B1_MMA_WARP) tmem_alloc(B.tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *B.tmem_slot;

  if (warp >= B1_EPI_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA_WG));
    if (warp == B1_MMA_WARP && lane == 0) {
      // ---------------------------------------------------------------- MMA issue: D[e, j] = ghv[e, :] . W2h^T[j, :]
      // (this thread and the TMA thread sleep 32 ns between polls: the 16-warp epilogue is the slow role here, and inside
      // the power-capped training step the quieter waits are worth 2 %: 1.27 -> 1.24 ms)
      load_weight_image(sW, p.W2thp, B.w);
      constexpr uint32_t IDESC = idesc_bf16(128, 256, true, false);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&B.tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * H;
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait_backoff<32>(&B.full[stage], phase);
          tc_fence_after();
          const uint32_t x_base = smem_u32(sA + stage * STAGE_BYTES);
          const uint32_t w_base = smem_u32(sW + kc * (H * KCHUNK * 2));
#pragma unroll
          for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
            if (!(dbg & 32))
              umma_bf16(d0, desc_mnmajor(x_base + ks * 2048, 8192, 1024), desc_kmajor(w_base + ks * UMMA_K * 2), IDESC,
                        (kc | ks) != 0 ? 1u : 0u);
          umma_commit(&B.empty[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&B.tfull[acc]);
      }
    } else if (warp == B1_TMA_WARP && lane == 0) {
      // ---------------------------------------------------------------- TMA: tile image K-chunks -> ring, verbatim
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait_backoff<32>(&B.empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&B.full[stage], STAGE_BYTES);
          bulk_g2s_stream(sA + stage * STAGE_BYTES, p.ghvT + (int64_t)tile * TILE_IMG_BYTES + kc * STAGE_BYTES, STAGE_BYTES,
                   &B.full[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: lane = edge, registers = features
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(B1_REGS_EPI));
    const int q = warp & 3, cq = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 64);
    auto load_meta = [&](int tile, int& nr, int& nc, float& dd) {
      int64_t e = (int64_t)tile * TILE_M + q * 32 + lane;
      e = e < p.E ? e : p.E - 1;
      nr = __ldg(p.row + e);
      nc = __ldg(p.col + e);
      dd = __ldg(p.d2 + e);
    };
    int nr, nc, nrn = 0, ncn = 0;
    float dd, ddn = 0.f;
    load_meta(blockIdx.x, nr, nc, dd);
    uint8_t* stg = smem + SmemB1::STG_OFF + warp * 4096;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int64_t e = (int64_t)tile * TILE_M + q * 32 + lane;
      const bool valid = e < p.E;
      const int next_tile = tile + gridDim.x;
      if (next_tile < p.num_tiles) load_meta(next_tile, nrn, ncn, ddn);
      const __half* arow = p.ABh + (int64_t)nr * 2 * H + cq * 64;
      const __half* brow = p.ABh + (int64_t)nc * 2 * H + H + cq * 64;
      uint4 a4[4], b4[4];
      ld_256(arow, a4[0], a4[1]);
      ld_256(arow + 16, a4[2], a4[3]);
      ld_256(brow, b4[0], b4[1]);
      ld_256(brow + 16, b4[2], b4[3]);
      float dot = 0.f;
      mbar_wait(&B.tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      if (dbg & 128) {
        tc_fence_before();
        mbar_arrive(&B.tempty[acc]);
        nr = nrn; nc = ncn; dd = ddn;
        continue;
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int col0 = cq * 64 + b * 32;
        uint32_t raw[32];
        tmem_ld32_issue(lane_addr + acc * H + b * 32, raw);
        float sab[32];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float s8[8];
          add_f16x8_to_f32(a4[k], b4[k], s8);
#pragma unroll
          for (int j = 0; j < 8; ++j) sab[8 * k + j] = s8[j];
        }
        if (b == 0) {
          ld_256(arow + 32, a4[0], a4[1]);
          ld_256(arow + 48, a4[2], a4[3]);
          ld_256(brow + 32, b4[0], b4[1]);
          ld_256(brow + 48, b4[2], b4[3]);
        }
        float r[32];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(sWd + col0 + 8 * k);
          const float4 w1 = *reinterpret_cast<const float4*>(sWd + col0 + 8 * k + 4);
          const float w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) r[8 * k + j] = silu_grad_r(fmaf(w8[j], dd, sab[8 * k + j]));
        }
        tmem_wait();
        if (b == 1) {
          tc_fence_before();
          mbar_arrive(&B.tempty[acc]);
        }
        float gu[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 w = *reinterpret_cast<const float4*>(sWd + col0 + 4 * k);
          gu[4 * k] = fmaf(__uint_as_float(raw[4 * k]), r[4 * k], __uint_as_float(raw[4 * k]));
          gu[4 * k + 1] = fmaf(__uint_as_float(raw[4 * k + 1]), r[4 * k + 1], __uint_as_float(raw[4 * k + 1]));
          gu[4 * k + 2] = fmaf(__uint_as_float(raw[4 * k + 2]), r[4 * k + 2], __uint_as_float(raw[4 * k + 2]));
          gu[4 * k + 3] = fmaf(__uint_as_float(raw[4 * k + 3]), r[4 * k + 3], __uint_as_float(raw[4 * k + 3]));
          dot = fmaf(gu[4 * k], w.x, dot);
          dot = fmaf(gu[4 * k + 1], w.y, dot);
          dot = fmaf(gu[4 * k + 2], w.z, dot);
          dot = fmaf(gu[4 * k + 3], w.w, dot);
        }
        if (!(dbg & 2)) {                         // ghu rows -> HBM: staged [32 x 32] box + TMA tensor store (see fwd2)
          uint8_t* sbuf = stg + b * 2048;
          if (lane == 0) bulk_wait_read_1();
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float o8[8] = {gu[8 * k], gu[8 * k + 1], gu[8 * k + 2], gu[8 * k + 3],
                                 gu[8 * k + 4], gu[8 * k + 5], gu[8 * k + 6], gu[8 * k + 7]};
            *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) = pack8(o8);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) tma_store_2d(&ghu_map, col0, (int)((int64_t)tile * TILE_M + q * 32), sbuf);
        }
      }
      if (valid) p.gd2_parts[(int64_t)cq * p.E + e] = dot;
      nr = nrn; nc = ncn; dd = ddn;
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == B1_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// =================================================================================================== wgrad
// Weight gradients as split-K GEMMs over the edge dimension, accumulated in TMEM over all tiles of a CTA (2 x 256
// columns = the whole 256 x 256 fp32 matrix) and written once per CTA to a partial buffer; wgrad_reduce_kernel sums
// the partials in a fixed order (deterministic).  Both operands are rebuilt on the fly, nothing but hs / mT / ghvT
// is read from HBM.  The ring unit is a HALF tile (64 edges): 32 KB of each operand, 3 stages.
//   MODE 5:  dW5h[k, f] = sum_e ghs[e, k] m[e, f]       A = ghs (MN-major, rows = edges), B = m (K-major image via TMA)
//            + column sums db5h[k] = sum_e ghs[e, k], dW6[k] = sum_e gw[e] t[e, k]  (in the row producers' registers)
//   MODE 2:  dW2h[f, j] = sum_e ghv[e, f] a[e, j]       A = ghv (K-major image via TMA), B = a = silu(hu) (MN-major, rows = edges)
// Warps: 0..15 row producers (4 per 64-column block, 16 rows each), 16 MMA issue, 17 TMA (the image operand).
// After the main loop warps 0..7 flush TMEM.
constexpr int WG_HALF_BYTES = 64 * H * 2;          // 32 KB: one operand of a half tile
constexpr int WG_STAGE_BYTES = 2 * WG_HALF_BYTES;  // 64 KB
constexpr int WG_STAGES = 3;
constexpr int WG_ROW_WARPS = 16;
constexpr int WG_MMA_WARP = 16;
constexpr int WG_TMA_WARP = 17;
constexpr int WG_VEC_OFF = WG_STAGES * WG_STAGE_BYTES;                 // 256 floats + 2 x [4 row groups][256] column sums
constexpr int WG_BAR_OFF = WG_VEC_OFF + 9 * H * 4;
constexpr int WG_SMEM_BYTES = WG_BAR_OFF + 128 + 1024;
template <int MODE> struct WgCfg {
  static constexpr int WARPS = 20;
  static constexpr int THREADS = 32 * WARPS;                            // 640 (96 registers per thread)
  static constexpr int FULL_COUNT = 32 * WG_ROW_WARPS + 1;              // row producers + the TMA thread's expect_tx
};

struct WgradParams {
  // MODE 5
  const __nv_bfloat16* hs;    // [E,256]
  const float* gw;            // [E]
  const float* w6;            // [256]
  const uint8_t* mT;          // tile images of m = silu(v)
  float* colsum_partial;      // [gridDim.x][2][256]: per-CTA sum_e ghs | sum_e gw t, reduced in fixed order by the launcher
  // MODE 2
  const uint8_t* ghvT;        // tile images of ghv
  const __half* ABh;           // [N,512] fp16
  const float* d2;            // [E]
  const int32_t* row;
  const int32_t* col;
  const float* wd;            // [256] (full domain; halved on load)
  // both
  float* partial;             // [gridDim.x][256*256] fp32
  int64_t E;
  int num_tiles;
};

template <int MODE>
__global__ void __launch_bounds__(WgCfg<MODE>::THREADS, 1) wgrad_kernel(const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* sVec = reinterpret_cast<float*>(smem + WG_VEC_OFF);          // MODE 5: w6 | db5h | dw6 ; MODE 2: wd/2
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_BAR_OFF);
  uint64_t* full = bars;                   // [WG_STAGES]
  uint64_t* empty = bars + WG_STAGES;      // [WG_STAGES]
  uint64_t* done = bars + 2 * WG_STAGES;   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = threadIdx.x; k < H; k += WgCfg<MODE>::THREADS) {
    sVec[k] = MODE == 5 ? p.w6[k] : 0.5f * p.wd[k];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full[s], WgCfg<MODE>::FULL_COUNT);
      mbar_init(&empty[s], 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nh = 2 * ((p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x);   // half tiles of this CTA
  auto tile_of = [&](int hi) { return (int)blockIdx.x + (hi >> 1) * (int)gridDim.x; };

  if (warp < WG_ROW_WARPS) {
    // ------------------------------------------------------------------ row producers: [64 rows][64 columns] blocks
    const int cq = warp & 3;                        // 64-column block of this warp
    const int chunk = lane & 7;
    const int rl0 = (warp >> 2) * 16 + (lane >> 3); // local rows rl0 + 4 i, i < 4
    const int c0 = cq * KCHUNK + chunk * 8;         // first of this thread's 8 columns
    float v8[8];                                    // MODE 5: w6 ; MODE 2: wd/2
#pragma unroll
    for (int j = 0; j < 8; ++j) v8[j] = sVec[c0 + j];
    int stage = 0;
    uint32_t phase = 0;
    if constexpr (MODE == 5) {
      float cs0[8], cs1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cs0[j] = cs1[j] = 0.f;
      uint4 pf[4];
      float gwv[4];
      auto issue = [&](int hi, int g) {
#pragma unroll
        for (int i = 2 * g; i < 2 * g + 2; ++i) {
          const int64_t e = (int64_t)tile_of(hi) * TILE_M + (hi & 1) * 64 + rl0 + 4 * i;
          const int64_t ec = e < p.E ? e : p.E - 1;
          pf[i] = __ldg(reinterpret_cast<const uint4*>(p.hs + ec * H + c0));
          gwv[i] = e < p.E ? __ldg(p.gw + e) : 0.f;
        }
      };
      issue(0, 0);
      issue(0, 1);
      for (int hi = 0; hi < nh; ++hi) {
        uint4 out[4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {                // rows {0,1} then {2,3}: each pair is re-issued for half tile hi + 1
#pragma unroll                                       // as soon as it has been consumed
          for (int i = 2 * g; i < 2 * g + 2; ++i) {
            float h8[8], o8[8];
            unpack8(pf[i], h8);
            const float gwi = gwv[i];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t;
              const float r = silu_grad_r(h8[j], t);
              o8[j] = gwi * fmaf(v8[j], r, v8[j]);
              cs0[j] += o8[j];
              cs1[j] = fmaf(gwi, t, cs1[j]);
            }
            out[i] = pack8(o8);
          }
          if (hi + 1 < nh) issue(hi + 1, g);
        }
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * WG_STAGE_BYTES + cq * 8192;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(st + sw128_offset(rl0 + 4 * i, chunk)) = out[i];
        fence_proxy_async();
        mbar_arrive(&full[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      // column sums: lanes l, l^8, l^16, l^24 share columns -> shared -> one global atomic per column per CTA
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a0 = cs0[j], a1 = cs1[j];
        a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
        if (lane < 8) {                              // slot of this warp's row group: no two warps share one
          sVec[H + (warp >> 2) * H + c0 + j] = a0;
          sVec[5 * H + (warp >> 2) * H + c0 + j] = a1;
        }
      }
    } else {
      uint4 pfA[4], pfB[4];
      float dd[4], ddn[4];
      int nr[4], nc[4];
      auto load_meta = [&](int hi) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int64_t e = (int64_t)tile_of(hi) * TILE_M + (hi & 1) * 64 + rl0 + 4 * i;
          e = e < p.E ? e : p.E - 1;                 // tail rows: finite values, multiplied by ghv = 0
          nr[i] = __ldg(p.row + e);
          nc[i] = __ldg(p.col + e);
          ddn[i] = __ldg(p.d2 + e);
        }
      };
      auto issue = [&](int g) {
#pragma unroll
        for (int i = 2 * g; i < 2 * g + 2; ++i) {
          pfA[i] = __ldg(reinterpret_cast<const uint4*>(p.ABh + (int64_t)nr[i] * 2 * H + c0));
          pfB[i] = __ldg(reinterpret_cast<const uint4*>(p.ABh + (int64_t)nc[i] * 2 * H + H + c0));
        }
      };
      load_meta(0);
      issue(0);
      issue(1);
#pragma unroll
      for (int i = 0; i < 4; ++i) dd[i] = ddn[i];
      if (nh > 1) load_meta(1);
      for (int hi = 0; hi < nh; ++hi) {
        uint4 out[4];
#pragma unroll
        for (int g = 0; g < 2; ++g) {                // rows {0,1} then {2,3}: each pair is re-issued for half tile hi + 1
#pragma unroll                                       // (metadata in nr / nc / ddn) as soon as it has been consumed
          for (int i = 2 * g; i < 2 * g + 2; ++i) {
            float s8[8], o8[8];
            add_f16x8_to_f32(pfA[i], pfB[i], s8);
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = silu_h(fmaf(v8[j], dd[i], s8[j]));
            out[i] = pack8(o8);
          }
          if (hi + 1 < nh) issue(g);
        }
        if (hi + 1 < nh) {
#pragma unroll
          for (int i = 0; i < 4; ++i) dd[i] = ddn[i];
          if (hi + 2 < nh) load_meta(hi + 2);
        }
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * WG_STAGE_BYTES + WG_HALF_BYTES + cq * 8192;
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(st + sw128_offset(rl0 + 4 * i, chunk)) = out[i];
        fence_proxy_async();
        mbar_arrive(&full[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == WG_MMA_WARP) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issue (accumulates over every half tile)
      constexpr uint32_t IDESC = MODE == 5 ? idesc_bf16(128, 256, true, false) : idesc_bf16(128, 256, false, true);
      int stage = 0;
      uint32_t phase = 0;
      for (int hi = 0; hi < nh; ++hi) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + stage * WG_STAGE_BYTES);
        const uint32_t b_base = a_base + WG_HALF_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
#pragma unroll
          for (int mh = 0; mh < 2; ++mh) {
            const uint32_t acc = (hi | ks) != 0 ? 1u : 0u;
            if constexpr (MODE == 5)
              umma_bf16(tmem_base + mh * H, desc_mnmajor(a_base + mh * 16384 + ks * 2048, 8192, 1024),
                        desc_kmajor(b_base + ks * UMMA_K * 2), IDESC, acc);
            else
              umma_bf16(tmem_base + mh * H, desc_kmajor(a_base + mh * 16384 + ks * UMMA_K * 2),
                        desc_mnmajor(b_base + ks * 2048, 8192, 1024), IDESC, acc);
          }
        umma_commit(&empty[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
    __syncwarp();
  } else if (warp == WG_TMA_WARP) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA: image half (4 x 8 KB) -> K-major operand
      // (MODE 2: ghv -> A operand; MODE 5: m -> B operand); the 8 KB pieces [fq] of the edge half become contiguous
      const uint8_t* img = MODE == 5 ? p.mT : p.ghvT;
      int stage = 0;
      uint32_t phase = 0;
      for (int hi = 0; hi < nh; ++hi) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], WG_HALF_BYTES);
        const uint8_t* src = img + (int64_t)tile_of(hi) * TILE_IMG_BYTES + (hi & 1) * 8192;
        uint8_t* dst = smem + stage * WG_STAGE_BYTES + (MODE == 5 ? WG_HALF_BYTES : 0);
#pragma unroll
        for (int fq = 0; fq < 4; ++fq) bulk_g2s_stream(dst + fq * 8192, src + fq * 16384, 8192, &full[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  }

  // -------------------------------------------------------------------- flush: TMEM -> this CTA's partial matrix
  if (warp < 8) {
    mbar_wait(done, 0);
    tc_fence_after();
    const int q = warp & 3, mh = warp >> 2;
    float* dstrow = p.partial + (int64_t)blockIdx.x * H * H + (int64_t)(mh * 128 + q * 32 + lane) * H;
#pragma unroll 1
    for (int cb = 0; cb < 8; ++cb) {
      uint32_t raw[32];
      tmem_ld32_issue(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mh * H + cb * 32), raw);
      tmem_wait();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<uint4*>(dstrow + cb * 32 + 4 * k) = make_uint4(raw[4 * k], raw[4 * k + 1], raw[4 * k + 2], raw[4 * k + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (MODE == 5 && threadIdx.x < H) {
    const int k = threadIdx.x;
    float* dst = p.colsum_partial + (int64_t)blockIdx.x * 2 * H;
    dst[k] = (sVec[H + k] + sVec[2 * H + k]) + (sVec[3 * H + k] + sVec[4 * H + k]);
    dst[H + k] = (sVec[5 * H + k] + sVec[6 * H + k]) + (sVec[7 * H + k] + sVec[8 * H + k]);
  }
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// out[e] = (parts[0][e] + parts[1][e]) + (parts[2][e] + parts[3][e])
__global__ void sum4_kernel(const float* __restrict__ parts, int64_t E, float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) out[e] = (parts[e] + parts[E + e]) + (parts[2 * E + e] + parts[3 * E + e]);
}

// out[i] = scale * sum_g partial[g][i]   (fixed summation order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int G, float scale, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * H) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int g = 0;
  for (; g + 3 < G; g += 4) {
    s0 += partial[(int64_t)g * H * H + i];
    s1 += partial[(int64_t)(g + 1) * H * H + i];
    s2 += partial[(int64_t)(g + 2) * H * H + i];
    s3 += partial[(int64_t)(g + 3) * H * H + i];
  }
  for (; g < G; ++g) s0 += partial[(int64_t)g * H * H + i];
  out[i] = scale * ((s0 + s1) + (s2 + s3));
}

// =================================================================================================== sums
// Segment sums of ghu = dL/dhu (bf16 [E,256]) over CSR rows and CSC columns, one WARP per node: a lane owns 8
// features (one 16-byte chunk of the 512-byte edge row, so every load is a fully coalesced row), fp32 accumulation
// in ascending edge order (deterministic), 4 edge rows in flight.  gA[i] = sum_{row e = i} ghu[e], gB[j] = sum_{col e = j}
// ghu[e]; the (wd/2)-gradient sum_e d2[e] ghu[e] stays in registers across the nodes a warp visits (one atomic per
// feature per warp at the end).  HBM-bound: ghu is read twice (1024 B per edge), gAB written once.
// (256, 3): three resident blocks per SM (see the launcher); without the second argument ptxas settles on 48 registers
// once the kernel contains a barrier and serialises the eight row loads in flight (0.89 -> 1.22 ms)
__global__ void __launch_bounds__(256, 3)
edge_sums_kernel(const uint4* __restrict__ ghu, const float* __restrict__ d2, const int32_t* __restrict__ row_ptr,
                 const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ csc_perm, int64_t N,
                 float* __restrict__ gAB, float* __restrict__ gwd_partial /*[gridDim.x][256]*/) {
  __shared__ float swd[8][H];
  constexpr int UNR = 8;                        // edge rows in flight per warp
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float wacc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) wacc[j] = 0.f;
  for (int64_t i = warp0; i < N; i += nwarps) {
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
    {
      const int64_t e0 = row_ptr[i], e1 = row_ptr[i + 1];
      int64_t e = e0;
      for (; e + UNR <= e1; e += UNR) {
        uint4 v[UNR];
        float dd[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          v[u] = __ldg(ghu + (e + u) * 32 + lane);
          dd[u] = __ldg(d2 + e + u);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          float g8[8];
          unpack8(v[u], g8);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            a[j] += g8[j];
            wacc[j] = fmaf(g8[j], dd[u], wacc[j]);
          }
        }
      }
      for (; e < e1; ++e) {
        float g8[8];
        unpack8(__ldg(ghu + e * 32 + lane), g8);
        const float dd = __ldg(d2 + e);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[j] += g8[j];
          wacc[j] = fmaf(g8[j], dd, wacc[j]);
        }
      }
    }
    {
      const int64_t q0 = col_ptr[i], q1 = col_ptr[i + 1];
      int64_t q = q0;
      for (; q + UNR <= q1; q += UNR) {
        uint4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) v[u] = __ldg(ghu + (int64_t)__ldg(csc_perm + q + u) * 32 + lane);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          float g8[8];
          unpack8(v[u], g8);
#pragma unroll
          for (int j = 0; j < 8; ++j) b[j] += g8[j];
        }
      }
      for (; q < q1; ++q) {
        float g8[8];
        unpack8(__ldg(ghu + (int64_t)__ldg(csc_perm + q) * 32 + lane), g8);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] += g8[j];
      }
    }
    float4* oa = reinterpret_cast<float4*>(gAB + i * 2 * H + lane * 8);
    float4* ob = reinterpret_cast<float4*>(gAB + i * 2 * H + H + lane * 8);
    oa[0] = make_float4(a[0], a[1], a[2], a[3]);
    oa[1] = make_float4(a[4], a[5], a[6], a[7]);
    ob[0] = make_float4(b[0], b[1], b[2], b[3]);
    ob[1] = make_float4(b[4], b[5], b[6], b[7]);
  }
  // sum_e d2 ghu: per-warp sums -> fixed-order sum over the block's eight warps -> per-block partial (no atomics)
#pragma unroll
  for (int j = 0; j < 8; ++j) swd[threadIdx.x >> 5][lane * 8 + j] = wacc[j];
  __syncthreads();
  const int k = threadIdx.x;
  gwd_partial[(int64_t)blockIdx.x * H + k] =
      ((swd[0][k] + swd[1][k]) + (swd[2][k] + swd[3][k])) + ((swd[4][k] + swd[5][k]) + (swd[6][k] + swd[7][k]));
}

// d2[e] = |x[row[e]] - x[col[e]]|^2  (models/en_gnn_decoder.py:61-62), one thread per edge
__global__ void edge_d2_kernel(const float* __restrict__ x, const int32_t* __restrict__ row,
                               const int32_t* __restrict__ col, int64_t E, float* __restrict__ d2) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t i = row[e], j = col[e];
  const float dx = x[3 * i] - x[3 * j], dy = x[3 * i + 1] - x[3 * j + 1], dz = x[3 * i + 2] - x[3 * j + 2];
  d2[e] = dx * dx + dy * dy + dz * dz;
}

// fp32 [256,256] (out,in) -> bf16 image of scale * W (or scale * W^T): 4 K-blocks of [256 rows x 128 B], SWIZZLE_128B
__global__ void pack_weight_scaled_kernel(const float* __restrict__ W, int transpose, float scale,
                                          __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * H) return;
  const int n = idx / H, k = idx % H;
  const float v = scale * (transpose ? W[k * H + n] : W[n * H + k]);
  const int kb = k / KCHUNK, kl = k % KCHUNK;
  const uint32_t byte = kb * (H * KCHUNK * 2) + sw128_offset(n, kl >> 3) + (kl & 7) * 2;
  out[byte >> 1] = __float2bfloat16(v);
}

template <typename K>
static int configure(K kernel, const char* name, int smem_bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return set_error(2, "%s: %s", name, cudaGetErrorString(e));
  return 0;
}
// Role-ablation switches (DBG = 1 instantiations, PEV_TC2_DEBUG bit mask) exist in profiling builds only
// (-DPEV_TC2_ABLATE); the shipped library instantiates the DBG = 0 forms, in which every `dbg &` test is a
// compile-time constant and folds away.
static int debug_mask() {
#ifdef PEV_TC2_ABLATE
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("PEV_TC2_DEBUG"); dbg = e ? atoi(e) : 0; }
  return dbg;
#else
  return 0;
#endif
}
static int grid_for(int num_tiles) {
  const int sms = sm_count();
  return num_tiles < sms ? num_tiles : sms;
}

}  // namespace tc2
}  // namespace pev

using namespace pev;
typedef __nv_bfloat16 bf16_t;

// workspace layout (floats): [sms][256*256] weight-gradient partials | [sms][256] db2 | [sms][2][256] db5, dw6 |
// [3 sms][256] gwd
static int64_t ws_off_db2() { return (int64_t)sm_count() * tc2::H * tc2::H; }
static int64_t ws_off_colsum() { return ws_off_db2() + (int64_t)sm_count() * tc2::H; }
static int64_t ws_off_gwd() { return ws_off_colsum() + (int64_t)sm_count() * 2 * tc2::H; }

template <int MODE>
static int launch_wgrad(tc2::WgradParams& p, float scale, float* dW, cudaStream_t st) {
  static bool configured_dev[pev::kMaxDevices] = {};   // the attribute is per device
  bool& configured = configured_dev[pev::current_device()];
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(tc2::wgrad_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         tc2::WG_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(2, "wgrad_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  p.num_tiles = (int)((p.E + tc2::TILE_M - 1) / tc2::TILE_M);
  const int grid = tc2::grid_for(p.num_tiles);
  tc2::wgrad_kernel<MODE><<<grid, tc2::WgCfg<MODE>::THREADS, tc2::WG_SMEM_BYTES, st>>>(p);
  if (int rc = after_launch(MODE == 5 ? "edge2_wgrad5_kernel" : "edge2_wgrad2_kernel")) return rc;
  tc2::wgrad_reduce_kernel<<<tc2::H * tc2::H / 256, 256, 0, st>>>(p.partial, grid, scale, dW);
  return after_launch("wgrad_reduce_kernel");
}


extern "C" {

int pev_pack_weight_bf16_scaled(const float* W, int32_t transpose, float scale, void* packed, void* stream) {
  PEV_REQUIRE(W && packed, "null argument");
  tc2::pack_weight_scaled_kernel<<<(tc2::H * tc2::H + 255) / 256, 256, 0, as_stream(stream)>>>(
      W, transpose, scale, reinterpret_cast<bf16_t*>(packed));
  return after_launch("pack_weight_scaled_kernel");
}

int64_t pev_edge2_tile_image_bytes(int64_t num_edges) {
  return ((num_edges + tc2::TILE_M - 1) / tc2::TILE_M) * (int64_t)tc2::TILE_IMG_BYTES;
}

int pev_edge_d2(const float* x, const int32_t* row, const int32_t* col, int64_t num_edges, float* d2, void* stream) {
  PEV_REQUIRE(num_edges >= 0, "bad argument");
  if (num_edges == 0) return 0;
  PEV_REQUIRE(x && row && col && d2, "null argument");
  tc2::edge_d2_kernel<<<(unsigned)((num_edges + 255) / 256), 256, 0, as_stream(stream)>>>(x, row, col, num_edges, d2);
  return after_launch("edge_d2_kernel");
}

int pev_edge2_fwd1(const void* ABh, const float* d2, const float* wd, const void* W2hp, const float* b2,
                   const int32_t* row, const int32_t* col, int64_t num_nodes, int64_t num_edges, void* hvT, void* mT,
                   float* agg, void* stream) {
  PEV_REQUIRE(ABh && wd && W2hp && b2 && agg && num_nodes >= 0 && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (num_nodes > 0) cudaMemsetAsync(agg, 0, sizeof(float) * tc2::H * (size_t)num_nodes, st);
  if (num_edges == 0) return 0;
  PEV_REQUIRE(row && col && mT && d2, "edge arrays missing");
  static bool configured_dev[pev::kMaxDevices] = {};   // the attribute is per device
  bool& configured = configured_dev[pev::current_device()];
  if (!configured) {
    if (int rc = tc2::configure(tc2::fwd1_kernel<0, true>, "fwd1_kernel", tc2::SmemT::BYTES)) return rc;
#ifdef PEV_TC2_ABLATE
    if (int rc = tc2::configure(tc2::fwd1_kernel<1, true>, "fwd1_kernel", tc2::SmemT::BYTES)) return rc;
#endif
    if (int rc = tc2::configure(tc2::fwd1_kernel<0, false>, "fwd1_kernel", tc2::SmemTI::BYTES)) return rc;
#ifdef PEV_TC2_ABLATE
    if (int rc = tc2::configure(tc2::fwd1_kernel<1, false>, "fwd1_kernel", tc2::SmemTI::BYTES)) return rc;
#endif
    configured = true;
  }
  tc2::Fwd1Params p = {};
  p.ABh = reinterpret_cast<const __half*>(ABh); p.d2 = d2; p.row = row; p.col = col; p.wd = wd; p.b2 = b2; p.W2hp = W2hp;
  p.hvT = reinterpret_cast<uint8_t*>(hvT); p.mT = reinterpret_cast<uint8_t*>(mT); p.agg = agg; p.E = num_edges;
  p.num_tiles = (int)((num_edges + tc2::TILE_M - 1) / tc2::TILE_M);
  p.dbg = tc2::debug_mask();
  const int grid = tc2::grid_for(p.num_tiles);
  if (hvT) {
#ifdef PEV_TC2_ABLATE
    if (p.dbg) tc2::fwd1_kernel<1, true><<<grid, tc2::NUM_THREADS, tc2::SmemT::BYTES, st>>>(p);
    else
#endif
    tc2::fwd1_kernel<0, true><<<grid, tc2::NUM_THREADS, tc2::SmemT::BYTES, st>>>(p);
  } else {
#ifdef PEV_TC2_ABLATE
    if (p.dbg) tc2::fwd1_kernel<1, false><<<grid, tc2::NUM_THREADS, tc2::SmemTI::BYTES, st>>>(p);
    else
#endif
    tc2::fwd1_kernel<0, false><<<grid, tc2::NUM_THREADS, tc2::SmemTI::BYTES, st>>>(p);
  }
  return after_launch("edge2_fwd1_kernel");
}

int pev_edge2_fwd2(const void* mT, const void* W5hp, const float* b5, const float* w6, const float* b6,
                   int64_t num_edges, float* w_out, void* hs_out, void* stream) {
  PEV_REQUIRE(W5hp && b5 && w6 && b6 && num_edges >= 0, "bad argument");
  if (num_edges == 0) return 0;
  PEV_REQUIRE(mT && w_out, "edge arrays missing");
  cudaStream_t st = as_stream(stream);
  static bool configured_dev[pev::kMaxDevices] = {};   // the attribute is per device
  bool& configured = configured_dev[pev::current_device()];
  if (!configured) {
    if (int rc = tc2::configure(tc2::fwd2_kernel<0, true>, "fwd2_kernel", tc2::SmemF2::BYTES)) return rc;
#ifdef PEV_TC2_ABLATE
    if (int rc = tc2::configure(tc2::fwd2_kernel<1, true>, "fwd2_kernel", tc2::SmemF2::BYTES)) return rc;
#endif
    if (int rc = tc2::configure(tc2::fwd2_kernel<0, false>, "fwd2_kernel", tc2::SmemF2::BYTES)) return rc;
#ifdef PEV_TC2_ABLATE
    if (int rc = tc2::configure(tc2::fwd2_kernel<1, false>, "fwd2_kernel", tc2::SmemF2::BYTES)) return rc;
#endif
    configured = true;
  }
  tc2::Fwd2Params p = {};
  p.mT = reinterpret_cast<const uint8_t*>(mT); p.W5hp = W5hp; p.b5 = b5; p.w6 = w6; p.b6 = b6;
  p.hs = reinterpret_cast<bf16_t*>(hs_out); p.w = w_out; p.E = num_edges;
  p.num_tiles = (int)((num_edges + tc2::TILE_M - 1) / tc2::TILE_M);
  p.dbg = tc2::debug_mask();
  alignas(64) CUtensorMap hs_map;
  memset(&hs_map, 0, sizeof(hs_map));
  if (hs_out)
    if (int rc = tc2::make_rows_map(hs_out, num_edges, &hs_map)) return rc;
  const int grid = tc2::grid_for(p.num_tiles);
  if (hs_out) {
#ifdef PEV_TC2_ABLATE
    if (p.dbg) tc2::fwd2_kernel<1, true><<<grid, tc2::F2_THREADS, tc2::SmemF2::BYTES, st>>>(p, hs_map);
    else
#endif
    tc2::fwd2_kernel<0, true><<<grid, tc2::F2_THREADS, tc2::SmemF2::BYTES, st>>>(p, hs_map);
  } else {
#ifdef PEV_TC2_ABLATE
    if (p.dbg) tc2::fwd2_kernel<1, false><<<grid, tc2::F2_THREADS, tc2::SmemF2::BYTES, st>>>(p, hs_map);
    else
#endif
    tc2::fwd2_kernel<0, false><<<grid, tc2::F2_THREADS, tc2::SmemF2::BYTES, st>>>(p, hs_map);
  }
  return after_launch("edge2_fwd2_kernel");
}

int pev_edge2_bwd2(const void* hs, const float* gw, const float* w6, const void* W5thp, const float* gagg,
                   const int32_t* row, const void* hvT, int64_t num_edges, float* workspace, void* ghvT, float* db2h,
                   void* stream) {
  PEV_REQUIRE(w6 && W5thp && db2h && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (num_edges == 0) {
    cudaMemsetAsync(db2h, 0, sizeof(float) * tc2::H, st);
    return 0;
  }
  PEV_REQUIRE(hs && gw && gagg && row && hvT && ghvT && workspace, "edge arrays missing");
  static bool configured_dev[pev::kMaxDevices] = {};   // the attribute is per device
  bool& configured = configured_dev[pev::current_device()];
  if (!configured) {
    if (int rc = tc2::configure(tc2::bwd2_kernel<0>, "bwd2_kernel", tc2::SmemB2::BYTES)) return rc;
#ifdef PEV_TC2_ABLATE
    if (int rc = tc2::configure(tc2::bwd2_kernel<1>, "bwd2_kernel", tc2::SmemB2::BYTES)) return rc;
#endif
    configured = true;
  }
  tc2::Bwd2Params p = {};
  p.hs = reinterpret_cast<const bf16_t*>(hs); p.gw = gw; p.w6 = w6; p.W5thp = W5thp; p.gagg = gagg; p.row = row;
  p.hvT = reinterpret_cast<const uint8_t*>(hvT); p.ghvT = reinterpret_cast<uint8_t*>(ghvT);
  p.db2_partial = workspace + ws_off_db2();
  p.E = num_edges;
  p.num_tiles = (int)((num_edges + tc2::TILE_M - 1) / tc2::TILE_M);
  p.dbg = tc2::debug_mask();
#ifdef PEV_TC2_ABLATE
  if (p.dbg) tc2::bwd2_kernel<1><<<tc2::grid_for(p.num_tiles), tc2::NUM_THREADS, tc2::SmemB2::BYTES, st>>>(p);
  else
#endif
  tc2::bwd2_kernel<0><<<tc2::grid_for(p.num_tiles), tc2::NUM_THREADS, tc2::SmemB2::BYTES, st>>>(p);
  if (int rc = after_launch("edge2_bwd2_kernel")) return rc;
  return launch_partial_reduce(p.db2_partial, tc2::grid_for(p.num_tiles), tc2::H, tc2::H, 1.0f, db2h, st);
}

int pev_edge2_bwd1(const void* ghvT, const void* W2thp, const void* ABh, const float* d2, const int32_t* row,
                   const int32_t* col, const float* wd, int64_t num_edges, void* ghu, float* gd2, float* gd2_parts,
                   void* stream) {
  PEV_REQUIRE(W2thp && wd && num_edges >= 0, "bad argument");
  if (num_edges == 0) return 0;
  PEV_REQUIRE(ghvT && ABh && d2 && row && col && ghu && gd2 && gd2_parts, "edge arrays missing");
  cudaStream_t st = as_stream(stream);
  static bool configured_dev[pev::kMaxDevices] = {};   // the attribute is per device
  bool& configured = configured_dev[pev::current_device()];
  if (!configured) {
    if (int rc = tc2::configure(tc2::bwd1_kernel<0>, "bwd1_kernel", tc2::SmemB1::BYTES)) return rc;
#ifdef PEV_TC2_ABLATE
    if (int rc = tc2::configure(tc2::bwd1_kernel<1>, "bwd1_kernel", tc2::SmemB1::BYTES)) return rc;
#endif
    configured = true;
  }
  tc2::Bwd1Params p = {};
  p.ghvT = reinterpret_cast<const uint8_t*>(ghvT); p.W2thp = W2thp; p.ABh = reinterpret_cast<const __half*>(ABh);
  p.d2 = d2; p.row = row; p.col = col; p.wd = wd; p.ghu = reinterpret_cast<bf16_t*>(ghu); p.gd2_parts = gd2_parts;
  p.E = num_edges;
  p.num_tiles = (int)((num_edges + tc2::TILE_M - 1) / tc2::TILE_M);
  p.dbg = tc2::debug_mask();
  alignas(64) CUtensorMap ghu_map;
  if (int rc = tc2::make_rows_map(ghu, num_edges, &ghu_map)) return rc;
#ifdef PEV_TC2_ABLATE
  if (p.dbg) tc2::bwd1_kernel<1><<<tc2::grid_for(p.num_tiles), tc2::B1_THREADS, tc2::SmemB1::BYTES, st>>>(p, ghu_map);
  else
#endif
  tc2::bwd1_kernel<0><<<tc2::grid_for(p.num_tiles), tc2::B1_THREADS, tc2::SmemB1::BYTES, st>>>(p, ghu_map);
  if (int rc = after_launch("edge2_bwd1_kernel")) return rc;
  tc2::sum4_kernel<<<(unsigned)((num_edges + 255) / 256), 256, 0, st>>>(gd2_parts, num_edges, gd2);
  return after_launch("sum4_kernel");
}

int pev_edge2_sums(const void* ghu, const float* d2, const int32_t* row_ptr, const int32_t* col_ptr,
                   const int32_t* csc_perm, int64_t num_nodes, int64_t num_edges, float* workspace, float* gAB,
                   float* gwdh, void* stream) {
  PEV_REQUIRE(row_ptr && col_ptr && gAB && gwdh && workspace && num_nodes >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (num_nodes == 0) {
    cudaMemsetAsync(gwdh, 0, sizeof(float) * tc2::H, st);
    return 0;
  }
  PEV_REQUIRE(num_edges == 0 || (ghu && d2 && csc_perm), "edge arrays missing");
  int64_t grid = (num_nodes + 7) / 8;
  // 3 blocks (24 warps) per SM: consecutive nodes go to consecutive warps, so the nodes in flight span ~3500 x 40 KB
  // of ghu, which stays in L2 between a row's first read (row pass of node i) and its second (column pass of the
  // neighbours i +- W); with all 32 resident warps per SM the window outgrows L2 (0.91 ms against 0.77 ms)
  const int64_t cap = (int64_t)sm_count() * 3;
  if (grid > cap) grid = cap;
  float* part = workspace + ws_off_gwd();
  tc2::edge_sums_kernel<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const uint4*>(ghu), d2, row_ptr, col_ptr,
                                                        csc_perm, num_nodes, gAB, part);
  if (int rc = after_launch("edge2_sums_kernel")) return rc;
  return launch_partial_reduce(part, (int)grid, tc2::H, tc2::H, 1.0f, gwdh, st);
}

int64_t pev_edge2_wgrad_workspace_bytes(void) {
  return (ws_off_gwd() + (int64_t)sm_count() * 3 * tc2::H) * (int64_t)sizeof(float);
}

int pev_edge2_wgrad5(const void* hs, const float* gw, const float* w6, const void* mT, int64_t num_edges,
                     float* workspace, float* dW5, float* db5, float* dw6, void* stream) {
  PEV_REQUIRE(w6 && dW5 && db5 && dw6 && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (num_edges == 0) {
    cudaMemsetAsync(db5, 0, sizeof(float) * tc2::H, st);
    cudaMemsetAsync(dw6, 0, sizeof(float) * tc2::H, st);
    cudaMemsetAsync(dW5, 0, sizeof(float) * tc2::H * tc2::H, st);
    return 0;
  }
  PEV_REQUIRE(hs && gw && mT && workspace, "edge arrays missing");
  tc2::WgradParams p = {};
  p.hs = reinterpret_cast<const bf16_t*>(hs); p.gw = gw; p.w6 = w6; p.mT = reinterpret_cast<const uint8_t*>(mT);
  p.colsum_partial = workspace + ws_off_colsum(); p.partial = workspace; p.E = num_edges;
  if (int rc = launch_wgrad<5>(p, 0.5f, dW5, st)) return rc;
  const int grid = tc2::grid_for((int)((num_edges + tc2::TILE_M - 1) / tc2::TILE_M));
  if (int rc = launch_partial_reduce(p.colsum_partial, grid, 2 * tc2::H, tc2::H, 1.0f, db5, st)) return rc;
  return launch_partial_reduce(p.colsum_partial + tc2::H, grid, 2 * tc2::H, tc2::H, 1.0f, dw6, st);
}

int pev_edge2_wgrad2(const void* ghvT, const void* ABh, const float* d2, const int32_t* row, const int32_t* col,
                     const float* wd, int64_t num_edges, float* workspace, float* dW2, void* stream) {
  PEV_REQUIRE(wd && dW2 && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (num_edges == 0) {
    cudaMemsetAsync(dW2, 0, sizeof(float) * tc2::H * tc2::H, st);
    return 0;
  }
  PEV_REQUIRE(ghvT && ABh && d2 && row && col && workspace, "edge arrays missing");
  tc2::WgradParams p = {};
  p.ghvT = reinterpret_cast<const uint8_t*>(ghvT); p.ABh = reinterpret_cast<const __half*>(ABh); p.d2 = d2;
  p.row = row; p.col = col; p.wd = wd; p.partial = workspace; p.E = num_edges;
  return launch_wgrad<2>(p, 0.5f, dW2, st);
}

}  // extern "C"
