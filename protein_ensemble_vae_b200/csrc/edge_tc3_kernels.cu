// K1, bf16 tensor-core form, third generation ("v3"): the edge MLP of one EGNN layer (models/en_gnn_decoder.py:60-79)
// as ONE kernel per direction, run by CTA PAIRS (tcgen05 cta_group::2):
//
//   fwd   hu = Ah_i + Bh_j + wdh d2 ; a = silu          (16 producer warps, gathers of the fp16 node projection)
//         --GEMM1: hv = a . W2h^T (+b2h)-->  TMEM  --E1: m = silu(hv) -> bf16 operand chunks in SHARED MEMORY-->
//         --GEMM2: hs = m . W5h^T (+b5h)-->  TMEM  --E2: t = silu(hs), w[e] = t . w6 + b6
//         agg[row] += m (two helper warps re-read the bf16 operand chunks from shared memory, one RED pair per segment).
//         The intermediate m never leaves the SM.  When a backward pass follows, hv and hs
//         are written as plain bf16 rows [E,256] through TMA tensor stores; m is NOT stored (silu(hv) is rebuilt where
//         it is needed: one tanh that the consumer needs for silu' anyway).
//
// Why a pair: both 256 x 256 bf16 weights (2 x 128 KB) must be resident for a fused kernel and do not fit one SM.  In
// cta_group::2 the B operand (the weight) is split by N across the two CTAs -- each holds output features
// [128 r, 128 r + 128) of BOTH weights (2 x 64 KB) -- while each CTA keeps its own 128 edges (M) through both GEMMs,
// so nothing but barrier arrivals crosses between the two SMs.  One tcgen05.mma covers 256 edges x 256 features x 16.
//
// Orientation: lane = edge everywhere (TMEM lane = edge, registers = features).  The per-edge outputs are rows, the
// feature reductions (w = t . w6) are in-thread; the edge reductions (agg) run on the operand copy in shared memory.
//
// TMEM: columns [0,256) = GEMM1 accumulator, [256,512) = GEMM2 accumulator (single-buffered).  One pool of eight
// epilogue warps alternates: E1 drains GEMM1's accumulator chunk by chunk while GEMM2 already consumes the finished
// chunks, then E2 drains GEMM2's accumulator under the next tile's GEMM1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"
#include "tc_common.cuh"
#include "tc_edge_common.cuh"

namespace pev {
namespace tc3 {
using namespace tcx;
using namespace tce;

constexpr int H = 256;
constexpr int TILE_M = 128;                 // edges per CTA and tile (a pair tile is 256 edges)
constexpr int KCHUNK = 64;                  // bf16 per 128-byte swizzle row
constexpr int NUM_KCHUNKS = H / KCHUNK;
constexpr int UMMA_K = 16;
constexpr int CHUNK_BYTES = TILE_M * KCHUNK * 2;   // 16 KB: [128 edges][64 features], K-major SWIZZLE_128B
constexpr int WHALF_BYTES = 128 * H * 2;           // 64 KB: 128 output features of one weight
constexpr int WFULL_KC_BYTES = H * KCHUNK * 2;     // 32 KB: one K-chunk of a full packed image (pev_pack_weight_bf16_scaled)
constexpr int TMEM_COLS = 512;

// Warp roles (warpgroup-aligned for setmaxnreg):
//   0..7   epilogue pool: warp (q = warp % 4: TMEM lane quarter, h = warp / 4: column half).  Per tile a pool warp first
//          runs E1 on GEMM1's accumulator (batch b: columns 64 b + 32 h .. + 31, i.e. all eight warps complete operand
//          chunk b together), then E2 on GEMM2's accumulator (columns 128 h + 32 b) -- E2 of tile n runs under GEMM1 of
//          tile n + 1, so neither the tensor pipe nor the pool idles while the other works
//   8      MMA issue (leader CTA only)      9   TMEM allocation + weight load
//   10,11  agg: segment sums of m read back from the bf16 operand chunks in shared memory (chunk parity = warp parity)
//   12..27 producers
constexpr int POOL_WARPS = 8;
constexpr int MMA_WARP = 8;
constexpr int AUX_WARP = 9;
constexpr int AGG_WARP0 = 10;
constexpr int PROD_WARP0 = 12;
constexpr int NUM_PROD_WARPS = 16;
constexpr int NUM_THREADS = 32 * (PROD_WARP0 + NUM_PROD_WARPS);   // 896 -> 72 registers at launch
constexpr int REGS_MMA_WG = 40;
constexpr int REGS_PROD = 64;
constexpr int REGS_POOL = 104;       // 40 + 4 x 64 + 2 x 104 = 504 <= 512 per warpgroup lane

template <bool TRAIN>
struct SmemF {
  static constexpr int A1_STAGES = 2;
  static constexpr int A2_SLOTS = TRAIN ? 2 : 3;
  static constexpr int W2_OFF = 0;
  static constexpr int W5_OFF = WHALF_BYTES;
  static constexpr int A1_OFF = 2 * WHALF_BYTES;
  static constexpr int A2_OFF = A1_OFF + A1_STAGES * CHUNK_BYTES;
  static constexpr int VEC_OFF = A2_OFF + A2_SLOTS * CHUNK_BYTES;     // b2h | b5h | w6 | wdh (vec_slot layout)
  static constexpr int STG_OFF = VEC_OFF + 4 * H * 4;                 // TRAIN: one [32 x 32] bf16 box (2 KB) per pool warp
  static constexpr int BAR_OFF = STG_OFF + (TRAIN ? POOL_WARPS * 2048 : 0);
  static constexpr int TOTAL = BAR_OFF + 256;
  static constexpr int BYTES = TOTAL + 1024;                          // slack for manual 1024-byte alignment
  static_assert(BYTES <= 232448, "shared memory budget");
};

struct BarsF {
  uint64_t* a1_full;    // [2] producers (both CTAs) -> MMA; waited at the leader only
  uint64_t* a1_empty;   // [2] MMA -> producers (multicast commit)
  uint64_t* a2_full;    // [3] pool (both CTAs) -> MMA (leader's barrier)
  uint64_t* a2_ready;   // [3] pool (this CTA) -> agg warps (local)
  uint64_t* a2_empty;   // [3] MMA (multicast commit) + this CTA's agg warp -> pool
  uint64_t* acc1_full;  // MMA -> pool
  uint64_t* acc1_empty; // pool (both CTAs) -> MMA
  uint64_t* acc2_full;  // MMA -> pool
  uint64_t* acc2_empty; // pool (both CTAs) -> MMA
  uint64_t* w;          // weight halves landed (local)
  uint64_t* wready;     // both CTAs' weights landed (leader's barrier)
  uint32_t* tmem_slot;
};
__device__ __forceinline__ BarsF make_bars_f(uint8_t* base) {
  uint64_t* b = reinterpret_cast<uint64_t*>(base);
  return BarsF{b, b + 2, b + 4, b + 7, b + 10, b + 13, b + 14, b + 15, b + 16, b + 17, b + 18,
               reinterpret_cast<uint32_t*>(b + 19)};
}

struct FwdParams {
  const __half* ABh;           // [N,512] fp16, half domain: 0.5 (h Wa^T + b1) | 0.5 h Wb^T
  const float* d2;            // [E] squared edge lengths
  const int32_t* row;         // [E]
  const int32_t* col;         // [E]
  const float* wd;            // [256] (full domain; halved on load)
  const float* b2;            // [256] (full domain; halved on load)
  const float* b5;            // [256] (full domain; halved on load)
  const float* w6;            // [256]
  const float* b6;            // [1]
  const uint8_t* W2hp;        // packed image of 0.5 W2 (pev_pack_weight_bf16_scaled)
  const uint8_t* W5hp;        // packed image of 0.5 W5
  float* agg;                 // [N,256] (+=, zeroed by the launcher)
  float* w;                   // [E] (+=, zeroed by the launcher: the two column halves of an edge add up)
  int64_t E;
  int num_pair_tiles;         // ceil(E / 256)
};

// ABL: role-ablation bit mask for profiling builds (-DPEV_TC3_ABLATE; production instantiates ABL = 0 only):
//   1 no segment sums, 2 E1 without tanh, 4 producers without gathers, 8 producers without tanh,
//   16 E2 without tanh / dot, 32 no MMAs (handshakes only)
template <bool TRAIN, int ABL = 0>   // TRAIN: hv and hs rows are written (a backward pass follows)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    fwd_kernel(const FwdParams p, const __grid_constant__ CUtensorMap hv_map, const __grid_constant__ CUtensorMap hs_map) {
  using SM = SmemF<TRAIN>;
  constexpr int A1_STAGES = SM::A1_STAGES;
  constexpr int A2_SLOTS = SM::A2_SLOTS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW2 = smem + SM::W2_OFF;
  uint8_t* sW5 = smem + SM::W5_OFF;
  uint8_t* sA1 = smem + SM::A1_OFF;
  uint8_t* sA2 = smem + SM::A2_OFF;
  float* sB2 = reinterpret_cast<float*>(smem + SM::VEC_OFF);
  float* sB5 = sB2 + H;
  float* sW6 = sB5 + H;
  float* sWd = sW6 + H;
  const BarsF B = make_bars_f(smem + SM::BAR_OFF);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = (int)cluster_id_x(), ncl = (int)num_clusters_x();
  const int n_it = (p.num_pair_tiles - cid + ncl - 1) / ncl;          // pair tiles of this cluster
  auto tile_of = [&](int it) { return 2 * (cid + it * ncl) + (int)rank; };   // 128-edge tile of this CTA

  for (int k = threadIdx.x; k < H; k += NUM_THREADS) {
    sB2[k] = 0.5f * p.b2[k];
    sB5[k] = 0.5f * p.b5[k];
    sW6[k] = p.w6[k];
    sWd[vec_slot(k)] = 0.5f * p.wd[k];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < A1_STAGES; ++s) {
      mbar_init(&B.a1_full[s], 2 * NUM_PROD_WARPS);
      mbar_init(&B.a1_empty[s], 1);
    }
    for (int s = 0; s < 3; ++s) {
      mbar_init(&B.a2_full[s], 2 * POOL_WARPS);
      mbar_init(&B.a2_ready[s], POOL_WARPS);
      mbar_init(&B.a2_empty[s], 2);
    }
    mbar_init(B.acc1_full, 1);
    mbar_init(B.acc1_empty, 2 * POOL_WARPS);
    mbar_init(B.acc2_full, 1);
    mbar_init(B.acc2_empty, 2 * POOL_WARPS);
    mbar_init(B.w, 1);
    mbar_init(B.wready, 2);
    fence_barrier_init();
  }
  if (warp == AUX_WARP) tmem_alloc_pair(B.tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *B.tmem_slot;

  if (warp >= MMA_WARP && warp < PROD_WARP0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA_WG));
    if (warp == AUX_WARP) {
      if (lane == 0) {
        // -------------------------------------------------------------- this CTA's halves of both weights
        mbar_arrive_expect_tx(B.w, 2 * WHALF_BYTES);
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          bulk_g2s(sW2 + kc * CHUNK_BYTES, p.W2hp + kc * WFULL_KC_BYTES + rank * CHUNK_BYTES, CHUNK_BYTES, B.w);
          bulk_g2s(sW5 + kc * CHUNK_BYTES, p.W5hp + kc * WFULL_KC_BYTES + rank * CHUNK_BYTES, CHUNK_BYTES, B.w);
        }
        mbar_wait(B.w, 0);
        mbar_arrive_cluster(mapa_bar(B.wready, 0));
      }
    } else if (warp == MMA_WARP) {
      if (lane == 0 && rank == 0) {
        // -------------------------------------------------------------- MMA issue (leader): D[e, n] over 256 edges
        mbar_wait_cluster(B.wready, 0);
        constexpr uint32_t IDESC = idesc_bf16(256, 256, false, false);
        int s1 = 0;
        uint32_t ph1 = 0;
        for (int it = 0; it < n_it; ++it) {
          // GEMM1: hv = a . W2h^T
          mbar_wait_cluster(B.acc1_empty, (it & 1) ^ 1);
          tc_fence_after();
          for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
            mbar_wait_cluster(&B.a1_full[s1], ph1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sA1 + s1 * CHUNK_BYTES);
            const uint32_t w_base = smem_u32(sW2 + kc * CHUNK_BYTES);
#pragma unroll
            for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
              if (!(ABL & 32))
                umma2_bf16(tmem_base, desc_kmajor(a_base + ks * UMMA_K * 2), desc_kmajor(w_base + ks * UMMA_K * 2), IDESC,
                           (kc | ks) != 0 ? 1u : 0u);
            umma2_commit_mc(&B.a1_empty[s1]);
            if (++s1 == A1_STAGES) { s1 = 0; ph1 ^= 1; }
          }
          umma2_commit_mc(B.acc1_full);
          // GEMM2: hs = m . W5h^T, chunk by chunk as the pool delivers m
          mbar_wait_cluster(B.acc2_empty, (it & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < NUM_KCHUNKS; ++c) {
            const int g = it * NUM_KCHUNKS + c, slot = g % A2_SLOTS;
            mbar_wait_cluster(&B.a2_full[slot], (g / A2_SLOTS) & 1);
            tc_fence_after();
            const uint32_t a_base = smem_u32(sA2 + slot * CHUNK_BYTES);
            const uint32_t w_base = smem_u32(sW5 + c * CHUNK_BYTES);
#pragma unroll
            for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
              if (!(ABL & 32))
                umma2_bf16(tmem_base + H, desc_kmajor(a_base + ks * UMMA_K * 2), desc_kmajor(w_base + ks * UMMA_K * 2), IDESC,
                           (c | ks) != 0 ? 1u : 0u);
            umma2_commit_mc(&B.a2_empty[slot]);
          }
          umma2_commit_mc(B.acc2_full);
        }
      }
    } else {
      // ---------------------------------------------------------------- agg[row] += m, from the operand chunks
      // Warp a takes the chunks of its parity.  A lane owns two features (one 4-byte column of the [128 edges x 64
      // features] chunk) and walks the 128 edges in order: fp32 accumulation of the bf16 operand values, one pair of
      // REDs per row segment.  A chunk's slot goes back to the pool when this warp AND the MMA commit have arrived.
      const int a = warp - AGG_WARP0;
      for (int it = 0; it < n_it; ++it) {
        const int tile = tile_of(it);
        int rows4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t e = (int64_t)tile * TILE_M + 32 * k + lane;
          rows4[k] = e < p.E ? __ldg(p.row + e) : -1;
        }
        for (int c = a; c < NUM_KCHUNKS; c += 2) {
          const int g = it * NUM_KCHUNKS + c, slot = g % A2_SLOTS;
          mbar_wait(&B.a2_ready[slot], (g / A2_SLOTS) & 1);
          if (!(ABL & 1)) {
            const uint8_t* src = sA2 + slot * CHUNK_BYTES + (lane & 3) * 4;
            const int ch = lane >> 2;
            float* aggcol = p.agg + c * KCHUNK + 2 * lane;
            // segment by segment (boundaries from the row ids, warp-uniform), branch-free inner loops with four
            // independent accumulator pairs
            int e0 = 0;
            while (e0 < TILE_M) {
              const int r = __shfl_sync(0xffffffffu, rows4[e0 >> 5], e0 & 31);
              int e1 = TILE_M;                       // first edge of the next segment
#pragma unroll
              for (int k = 3; k >= 0; --k) {
                const uint32_t diff = __ballot_sync(0xffffffffu, rows4[k] != r && 32 * k + lane > e0);
                if (diff) e1 = 32 * k + __ffs(diff) - 1;
              }
              float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
              int er = e0;
              for (; er + 4 <= e1; er += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t v = *reinterpret_cast<const uint32_t*>(src + sw128_offset(er + u, ch));
                  s0[u] += bf16_lo(v);
                  s1[u] += bf16_hi(v);
                }
              }
              for (; er < e1; ++er) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(src + sw128_offset(er, ch));
                s0[0] += bf16_lo(v);
                s1[0] += bf16_hi(v);
              }
              if (r >= 0) {
                atomicAdd(aggcol + (int64_t)r * H, (s0[0] + s0[1]) + (s0[2] + s0[3]));
                atomicAdd(aggcol + (int64_t)r * H + 1, (s1[0] + s1[1]) + (s1[2] + s1[3]));
              }
              e0 = e1;
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&B.a2_empty[slot]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= PROD_WARP0) {
    // ------------------------------------------------------------------ producers: a = silu(hu) -> K-major ring
    // Warp pw owns rows 8 pw .. 8 pw + 7 of every tile; lane -> row (lane >> 3) + 4 i (i < 2), 16-byte column chunk
    // lane & 7.  Row metadata (row, col, d2) is fetched one tile ahead and kept lane-distributed, the A | B operand
    // pieces one K-chunk ahead (two register slots).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PROD));
    const int pw = warp - PROD_WARP0;
    const int chunk = lane & 7;
    constexpr int RPT = 2;
    const int r0 = pw * 8 + (lane >> 3);
    int stage = 0;
    uint32_t phase = 0;
    struct Meta { int r, c; float d; };
    auto load_meta = [&](int tile, Meta& m) {
      int64_t e = (int64_t)tile * TILE_M + pw * 8 + (lane & 7);
      e = e < p.E ? e : p.E - 1;                   // rows past E recompute the last edge; the epilogues drop them
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
        if (ABL & 4) { pfA[slot][i] = pfB[slot][i] = make_uint4(nr, nc, 0u, 0u); continue; }
        pfA[slot][i] = __ldg(reinterpret_cast<const uint4*>(p.ABh + (int64_t)nr * 2 * H + kc * KCHUNK + chunk * 8));
        pfB[slot][i] = __ldg(reinterpret_cast<const uint4*>(p.ABh + (int64_t)nc * 2 * H + H + kc * KCHUNK + chunk * 8));
      }
    };
    const uint32_t full_leader = mapa_bar(B.a1_full, 0);
    Meta mc, mn;
    if (n_it > 0) {
      load_meta(tile_of(0), mc);
      mn = mc;
      issue(mc, 0, 0);
    }
    for (int it = 0; it < n_it; ++it) {
      const bool has_next = it + 1 < n_it;
      if (has_next) load_meta(tile_of(it + 1), mn);
      float d2r[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i) d2r[i] = __shfl_sync(0xffffffffu, mc.d, (lane >> 3) + 4 * i);
#pragma unroll
      for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
        if (kc + 1 < NUM_KCHUNKS) issue(mc, kc + 1, (kc + 1) & 1);
        else if (has_next) issue(mn, 0, 0);
        const int k0 = kc * KCHUNK + chunk * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(sWd + vec_slot(k0)), w1 = *reinterpret_cast<const float4*>(sWd + vec_slot(k0 + 4));
        const float wd8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        uint4 out[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          float s8[8], o8[8];
          add_f16x8_to_f32(pfA[kc & 1][i], pfB[kc & 1][i], s8);
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = (ABL & 8) ? fmaf(wd8[j], d2r[i], s8[j]) : silu_h(fmaf(wd8[j], d2r[i], s8[j]));
          out[i] = pack8(o8);
        }
        mbar_wait(&B.a1_empty[stage], phase ^ 1);
        uint8_t* st = sA1 + stage * CHUNK_BYTES;
#pragma unroll
        for (int i = 0; i < RPT; ++i) *reinterpret_cast<uint4*>(st + sw128_offset(r0 + 4 * i, chunk)) = out[i];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(full_leader + stage * 8);
        if (++stage == A1_STAGES) { stage = 0; phase ^= 1; }
      }
      mc = mn;
    }
  } else {
    // ------------------------------------------------------------------ epilogue pool
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_POOL));
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int erow = q * 32 + lane;                                  // this thread's row of the tile
    const uint32_t a2full_leader = mapa_bar(B.a2_full, 0);
    const uint32_t acc1empty_leader = mapa_bar(B.acc1_empty, 0);
    const uint32_t acc2empty_leader = mapa_bar(B.acc2_empty, 0);
    const float b6 = h == 0 ? __ldg(p.b6) : 0.f;
    uint8_t* stg = smem + SM::STG_OFF + warp * 2048;                 // TRAIN: [32 edges x 32 features] bf16 box
    // TMA tensor store of the warp's box (64-byte rows, SWIZZLE_64B chunk positions: conflict-free); the engine clips
    // the rows past E.  One buffer per warp: the previous store has read it a whole batch ago.
    auto store_box = [&](const CUtensorMap* map, const float (&val)[32], int c0, int tile) {
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
      if (lane == 0) tma_store_2d(map, c0, (int)((int64_t)tile * TILE_M + q * 32), stg);
    };
    for (int it = 0; it < n_it; ++it) {
      const int tile = tile_of(it);
      const int64_t e = (int64_t)tile * TILE_M + erow;
      // ---- E1: hv (+b2h) -> m = silu -> bf16 operand chunk b (this warp: its 32 edges x columns 64 b + 32 h .. + 31)
      mbar_wait(B.acc1_full, it & 1);
      tc_fence_after();
      uint32_t raw[32];
      tmem_ld32_issue(lane_base + 32 * h, raw);
#pragma unroll
      for (int b = 0; b < NUM_KCHUNKS; ++b) {
        const int c0 = 64 * b + 32 * h;
        const int g = it * NUM_KCHUNKS + b, slot = g % A2_SLOTS;
        tmem_wait();
        float val[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(sB2 + c0 + 4 * j4);
          val[4 * j4] = __uint_as_float(raw[4 * j4]) + bb.x;
          val[4 * j4 + 1] = __uint_as_float(raw[4 * j4 + 1]) + bb.y;
          val[4 * j4 + 2] = __uint_as_float(raw[4 * j4 + 2]) + bb.z;
          val[4 * j4 + 3] = __uint_as_float(raw[4 * j4 + 3]) + bb.w;
        }
        if (b + 1 < NUM_KCHUNKS) tmem_ld32_issue(lane_base + c0 + 64, raw);
        else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc1empty_leader);      // GEMM1's accumulator is drained (this warp's part)
        }
        if (TRAIN) store_box(&hv_map, val, c0, tile);
        uint4 o4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float o8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = (ABL & 2) ? val[8 * k + j] : silu_h(val[8 * k + j]);
          o4[k] = pack8(o8);
        }
        mbar_wait(&B.a2_empty[slot], ((g / A2_SLOTS) & 1) ^ 1);     // GEMM2 and the agg warp are done with the slot
        uint8_t* dst = sA2 + slot * CHUNK_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(dst + sw128_offset(erow, 4 * h + k)) = o4[k];
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(a2full_leader + slot * 8);
          mbar_arrive(&B.a2_ready[slot]);
        }
      }
      // ---- E2: hs (+b5h), t = silu, w += t . w6 (+ b6) over this warp's column half
      mbar_wait(B.acc2_full, it & 1);
      tc_fence_after();
      float dot = 0.f;
      tmem_ld32_issue(lane_base + H + 128 * h, raw);
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int c0 = 128 * h + 32 * b;
        tmem_wait();
        float val[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 bb = *reinterpret_cast<const float4*>(sB5 + c0 + 4 * j4);
          val[4 * j4] = __uint_as_float(raw[4 * j4]) + bb.x;
          val[4 * j4 + 1] = __uint_as_float(raw[4 * j4 + 1]) + bb.y;
          val[4 * j4 + 2] = __uint_as_float(raw[4 * j4 + 2]) + bb.z;
          val[4 * j4 + 3] = __uint_as_float(raw[4 * j4 + 3]) + bb.w;
        }
        if (b + 1 < 4) tmem_ld32_issue(lane_base + H + c0 + 32, raw);
        else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc2empty_leader);      // GEMM2's accumulator is drained (this warp's part)
        }
        if (TRAIN) store_box(&hs_map, val, c0, tile);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 w = *reinterpret_cast<const float4*>(sW6 + c0 + 4 * j4);
          if (ABL & 16) { dot += val[4 * j4] + w.x; continue; }
          dot = fmaf(silu_h(val[4 * j4]), w.x, dot);
          dot = fmaf(silu_h(val[4 * j4 + 1]), w.y, dot);
          dot = fmaf(silu_h(val[4 * j4 + 2]), w.z, dot);
          dot = fmaf(silu_h(val[4 * j4 + 3]), w.w, dot);
        }
      }
      if (e < p.E) atomicAdd(p.w + e, dot + b6);
    }
    if (TRAIN && lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's MMAs / multicast commits are done with this CTA's memory
  if (warp == AUX_WARP) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

template <typename K>
static int configure(K kernel, const char* name, int smem_bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return set_error(2, "%s: %s", name, cudaGetErrorString(e));
  return 0;
}
static int clusters_for(int num_pair_tiles) {
  const int pairs = sm_count() / 2;
  return num_pair_tiles < pairs ? num_pair_tiles : pairs;
}

}  // namespace tc3
}  // namespace pev

using namespace pev;
typedef __nv_bfloat16 bf16_t;

extern "C" {

int pev_edge3_fwd(const void* ABh, const float* d2, const float* wd, const void* W2hp, const float* b2, const void* W5hp,
                  const float* b5, const float* w6, const float* b6, const int32_t* row, const int32_t* col,
                  int64_t num_nodes, int64_t num_edges, void* hv_rows, void* hs_rows, float* agg, float* w_out,
                  void* stream) {
  PEV_REQUIRE(ABh && wd && W2hp && b2 && W5hp && b5 && w6 && b6 && agg && num_nodes >= 0 && num_edges >= 0, "bad argument");
  PEV_REQUIRE((hv_rows == nullptr) == (hs_rows == nullptr), "hv_rows and hs_rows go together");
  cudaStream_t st = as_stream(stream);
  if (num_nodes > 0) cudaMemsetAsync(agg, 0, sizeof(float) * tc3::H * (size_t)num_nodes, st);
  if (num_edges == 0) return 0;
  PEV_REQUIRE(row && col && d2 && w_out, "edge arrays missing");
  cudaMemsetAsync(w_out, 0, sizeof(float) * (size_t)num_edges, st);
  static bool configured_dev[pev::kMaxDevices] = {};   // the attribute is per device
  bool& configured = configured_dev[pev::current_device()];
  if (!configured) {
    if (int rc = tc3::configure(tc3::fwd_kernel<true>, "edge3_fwd_kernel", tc3::SmemF<true>::BYTES)) return rc;
    if (int rc = tc3::configure(tc3::fwd_kernel<false>, "edge3_fwd_kernel", tc3::SmemF<false>::BYTES)) return rc;
    configured = true;
  }
  tc3::FwdParams p = {};
  p.ABh = reinterpret_cast<const __half*>(ABh); p.d2 = d2; p.row = row; p.col = col; p.wd = wd; p.b2 = b2; p.b5 = b5;
  p.w6 = w6; p.b6 = b6; p.W2hp = reinterpret_cast<const uint8_t*>(W2hp); p.W5hp = reinterpret_cast<const uint8_t*>(W5hp);
  p.agg = agg; p.w = w_out; p.E = num_edges;
  p.num_pair_tiles = (int)((num_edges + 2 * tc3::TILE_M - 1) / (2 * tc3::TILE_M));
  alignas(64) CUtensorMap hv_map, hs_map;
  memset(&hv_map, 0, sizeof(hv_map));
  memset(&hs_map, 0, sizeof(hs_map));
  if (hv_rows) {
    if (int rc = tce::make_rows_map(hv_rows, num_edges, &hv_map)) return rc;
    if (int rc = tce::make_rows_map(hs_rows, num_edges, &hs_map)) return rc;
  }
  const int grid = 2 * tc3::clusters_for(p.num_pair_tiles);
#ifdef PEV_TC3_ABLATE
  {
    const char* env = getenv("PEV_TC3_ABL");
    const int abl = env ? atoi(env) : 0;
#define PEV_ABL_CASE(A)                                                                                              \
  if (abl == A && !hv_rows) {                                                                                        \
    tc3::configure(tc3::fwd_kernel<false, A>, "edge3_fwd_kernel", tc3::SmemF<false>::BYTES);                         \
    tc3::fwd_kernel<false, A><<<grid, tc3::NUM_THREADS, tc3::SmemF<false>::BYTES, st>>>(p, hv_map, hs_map);          \
    return after_launch("edge3_fwd_kernel");                                                                         \
  }
    PEV_ABL_CASE(1) PEV_ABL_CASE(2) PEV_ABL_CASE(3) PEV_ABL_CASE(4) PEV_ABL_CASE(12) PEV_ABL_CASE(16) PEV_ABL_CASE(19)
    PEV_ABL_CASE(32) PEV_ABL_CASE(31) PEV_ABL_CASE(63) PEV_ABL_CASE(15) PEV_ABL_CASE(28)
#undef PEV_ABL_CASE
  }
#endif
  if (hv_rows) tc3::fwd_kernel<true><<<grid, tc3::NUM_THREADS, tc3::SmemF<true>::BYTES, st>>>(p, hv_map, hs_map);
  else tc3::fwd_kernel<false><<<grid, tc3::NUM_THREADS, tc3::SmemF<false>::BYTES, st>>>(p, hv_map, hs_map);
  return after_launch("edge3_fwd_kernel");
}

}  // extern "C"
