// K1, bf16 tensor-core form: the EGNN edge MLP (models/en_gnn_decoder.py:60-79) on tcgen05.
//
// One persistent, warp-specialised kernel template, instantiated four times (SURVEY.md 8a math):
//   STAGE 1  a = silu(A_i + B_j + wd d2)  --GEMM W2-->  v = . + b2 ; v -> HBM (bf16), m = silu(v),
//            agg[row] += m   (segmented butterfly reduction per warp, RED.ADD.F32 per segment)
//   STAGE 2  m = silu(v)                  --GEMM W5-->  s = . + b5 ; t = silu(s), w[e] = t . w6 + b6
//   STAGE 3  gs = gw w6 silu'(s)          --GEMM W5^T-> gm = . + gagg[row] ; gv = gm silu'(v) -> HBM
//            (column sums db5 = sum gs, dW6 = sum gw t accumulated in the producers' registers)
//   STAGE 4  gv                           --GEMM W2^T-> ga ; gu = ga silu'(u) -> HBM, gd2[e] = gu . wd
//            (column sum db2 = sum gv in the producers)
// Tile = 128 consecutive edges (rows of the GEMM) x N = 256 x K = 256.
//   * the 256x256 bf16 weight stays resident in shared memory for the whole kernel (128 KB,
//     K-major, 128-byte swizzle), fetched once per CTA with cp.async.bulk (TMA, no tensor map);
//   * producer warps build the A operand on the fly (gather + SiLU -> bf16) straight into a ring of
//     [128 x 64] K-chunks in the UMMA canonical K-major SWIZZLE_128B layout;
//   * one elected thread issues tcgen05.mma (M=128, N=256, K=16; fp32 accumulators in TMEM),
//     two accumulator stages (2 x 256 TMEM columns) so the epilogue of tile t overlaps the MMAs of t+1;
//   * epilogue warps read TMEM with tcgen05.ld (32x32b.x32), apply bias / SiLU and write results.
// Synchronisation is mbarrier-only (full/empty per ring stage, full/empty per accumulator stage).
#include <cuda_bf16.h>

#include <cstdlib>

#include "../../include/pev_b200.h"
#include "pev_common.cuh"

namespace pev {
namespace tc {

constexpr int H = 256;                 // node_dim == hidden_dim of the reference decoder (F1)
constexpr int TILE_M = 128;            // edges per tile
constexpr int KCHUNK = 64;             // bf16 elements per 128-byte swizzle row
constexpr int NUM_KCHUNKS = H / KCHUNK;
constexpr int UMMA_K = 16;
constexpr int NUM_STAGES = 4;          // A-operand ring (16 KB each): one full tile in flight
constexpr int NA = 4;                  // destination rows of a tile staged in shared memory (stage 1)
constexpr int EPI_ROW = 80;            // bytes per row of an epilogue warp's transposition scratch (64 + 16 pad)
constexpr int EPI_SCRATCH = 32 * EPI_ROW;
constexpr int STAGE_BYTES = TILE_M * KCHUNK * 2;
constexpr int W_BYTES = H * H * 2;
// Roles are aligned to warpgroups (4 warps) so that setmaxnreg can move registers between them:
constexpr int NUM_EPI_WARPS = 8;       // warps 0..7   (TMEM lane quarter = warp % 4, column half = warp / 4)
constexpr int MMA_WARP = 8;            // warp 8       (TMEM alloc, weight load, MMA issue); warps 9..11 idle
constexpr int PROD_WARP0 = 12;         // warps 12..19 producers
constexpr int NUM_PROD_WARPS = 8;
constexpr int NUM_THREADS = 32 * (PROD_WARP0 + NUM_PROD_WARPS);   // 640 -> 96 registers per thread at launch
constexpr int REGS_MMA_WG = 32;        // setmaxnreg.dec in the MMA warpgroup frees 128 x 64 registers ...
constexpr int REGS_PROD = 128;         // ... which the 256 producer threads take (96 -> 128)
constexpr int NUM_PROD_THREADS = 32 * NUM_PROD_WARPS;
constexpr int NUM_EPI_THREADS = 32 * NUM_EPI_WARPS;
constexpr int TMEM_COLS = 512;

struct __align__(16) SmemLayout {
  // offsets into dynamic shared memory (base aligned to 1024)
  static constexpr int W_OFF = 0;
  static constexpr int A_OFF = W_BYTES;
  static constexpr int VEC_OFF = A_OFF + NUM_STAGES * STAGE_BYTES;      // 4 x 256 floats
  static constexpr int META_OFF = VEC_OFF + 4 * H * 4;                  // 2 x 128 floats (d2 per tile row)
  static constexpr int AROW_OFF = META_OFF + 2 * TILE_M * 4;            // 2 x NA x 256 floats (A rows of the tile)
  static constexpr int EPI_OFF = AROW_OFF + 2 * NA * H * 4;             // 8 warps x [32 rows x 80 B]
  static constexpr int BAR_OFF = EPI_OFF + NUM_EPI_WARPS * EPI_SCRATCH;
  static constexpr int TOTAL = BAR_OFF + 256;
};
constexpr int SMEM_BYTES = SmemLayout::TOTAL + 1024;   // slack for manual 1024-byte alignment

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), LBO = 1
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(H >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ------------------------------------------------------------------------------------------ math
// silu(z) = z sigmoid(z) = h + h tanh(h), h = z/2: one MUFU.TANH + 2 FMA-pipe ops
__device__ __forceinline__ float silu_fast(float z) {
  float h = 0.5f * z, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// byte offset of 16-byte chunk `chunk` (0..7) of row `r` inside a K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int chunk) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}

struct Params {
  // graph / node inputs
  const float* AB;           // [N, 512] fp32: A (+b1) | B            (stages 1, 4)
  const float* x;            // [N, 3]                                (stages 1, 4)
  const int32_t* row;        // [E]                                   (stages 1, 3, 4)
  const int32_t* col;        // [E]                                   (stages 1, 4)
  const float* vec1;         // [256]: wd (stages 1, 4), w6 (stages 2, 3)
  const float* bias;         // [256]: b2 (stage 1), b5 (stage 2), unused otherwise
  const void* Wp;            // packed weight image (W2, W5, W5^T, W2^T)
  // per-edge streams
  const __nv_bfloat16* in0;  // stage 2: v; stage 3: s; stage 4: gv
  const __nv_bfloat16* in1;  // epilogue stream: stage 3: dm = silu'(v); stage 4: da = silu'(u)
  const float* ein;          // stage 3: gw[E]
  const float* nin;          // stage 3: gagg[N,256]
  __nv_bfloat16* out0;       // stage 1: v; stage 2: s (or null); stage 3: gv; stage 4: gu
  __nv_bfloat16* out1;       // stage 1: a (or null); stage 2: m (or null); stage 3: gs
  __nv_bfloat16* out2;       // stage 1: da = silu'(u) (or null); stage 2: dm = silu'(v) (or null)
  float* eout;               // stage 2: w[E] (+=); stage 4: gd2[E] (+=)
  float* nout;               // stage 1: agg[N,256] (+=)
  float* csum0;              // [256] (+=): stage 3 db5; stage 4 db2
  float* csum1;              // [256] (+=): stage 3 dW6
  const float* b6;           // [1] stage 2
  int64_t E;
  int num_tiles;
  int dbg;                   // PEV_TC_DEBUG bit mask (profiling experiments only; 0 in production)
};

// silu(z) and d silu / dz from one tanh: sg = sigmoid(z) = 0.5 + 0.5 tanh(z/2)
__device__ __forceinline__ void silu_and_grad(float z, float& f, float& df) {
  float h = 0.5f * z, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  const float sg = fmaf(0.5f, t, 0.5f);
  f = z * sg;
  df = fmaf(f, 1.0f - sg, sg);
}
__device__ __forceinline__ float silu_grad(float z) {
  float f, df;
  silu_and_grad(z, f, df);
  return df;
}
__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&val)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    d[j] = make_uint4(pack_bf16(val[8 * j], val[8 * j + 1]), pack_bf16(val[8 * j + 2], val[8 * j + 3]),
                      pack_bf16(val[8 * j + 4], val[8 * j + 5]), pack_bf16(val[8 * j + 6], val[8 * j + 7]));
}
__device__ __forceinline__ void load_f32x8(const float* src, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* src, float (&o)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
  o[0] = bf16_lo(v.x); o[1] = bf16_hi(v.x); o[2] = bf16_lo(v.y); o[3] = bf16_hi(v.y);
  o[4] = bf16_lo(v.z); o[5] = bf16_hi(v.z); o[6] = bf16_lo(v.w); o[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&o)[8]) {
  return make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
}

// ---- epilogue transposition through a per-warp scratch ([32 rows][80 B], conflict-free both ways).
// In the TMEM layout a lane owns a ROW of the tile, so a direct global access touches 32 rows x 16 B per
// instruction (32 L2 transactions).  Going through the scratch, 4 lanes cover 64 contiguous bytes of a row:
// 8 rows x 64 B per instruction -- 4x fewer, full-sector transactions.
__device__ __forceinline__ void warp_store_rows(uint8_t* sc, int lane, const uint4 (&mine)[4], __nv_bfloat16* gbase,
                                                int nrows) {
#pragma unroll
  for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(sc + lane * EPI_ROW + k * 16) = mine[k];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = (lane >> 2) + 8 * j;
    const uint4 v = *reinterpret_cast<const uint4*>(sc + row * EPI_ROW + (lane & 3) * 16);
    if (row < nrows) *reinterpret_cast<uint4*>(gbase + (int64_t)row * H + (lane & 3) * 8) = v;
  }
  __syncwarp();
}
__device__ __forceinline__ void warp_issue_rows(const __nv_bfloat16* gbase, int lane, int nrows, uint4 (&buf)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = (lane >> 2) + 8 * j;
    buf[j] = (row < nrows) ? __ldg(reinterpret_cast<const uint4*>(gbase + (int64_t)row * H + (lane & 3) * 8))
                           : make_uint4(0u, 0u, 0u, 0u);
  }
}
__device__ __forceinline__ void warp_deposit_rows(uint8_t* sc, int lane, const uint4 (&buf)[4], uint4 (&mine)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(sc + ((lane >> 2) + 8 * j) * EPI_ROW + (lane & 3) * 16) = buf[j];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 4; ++k) mine[k] = *reinterpret_cast<const uint4*>(sc + lane * EPI_ROW + k * 16);
  __syncwarp();
}
__device__ __forceinline__ void pack32(const float (&val)[32], uint4 (&o)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    o[k] = make_uint4(pack_bf16(val[8 * k], val[8 * k + 1]), pack_bf16(val[8 * k + 2], val[8 * k + 3]),
                      pack_bf16(val[8 * k + 4], val[8 * k + 5]), pack_bf16(val[8 * k + 6], val[8 * k + 7]));
}

template <int STAGE>
__global__ void __launch_bounds__(NUM_THREADS, 1) edge_mlp_kernel(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B.  Plain pointer arithmetic on the __shared__ array (no integer round
  // trip) so the compiler keeps the shared address space and emits LDS/STS instead of generic LD/ST.
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW = smem + SmemLayout::W_OFF;
  uint8_t* sA = smem + SmemLayout::A_OFF;
  float* sBias = reinterpret_cast<float*>(smem + SmemLayout::VEC_OFF);   // b2 / b5 (stages 1, 2)
  float* sVec1 = sBias + H;                                               // wd / w6
  float* sRed = sBias + 2 * H;                                            // column-sum scratch (stages 3, 4)
  float* sRed2 = sBias + 3 * H;                                           // second column sum (stage 3)
  float* sMeta = reinterpret_cast<float*>(smem + SmemLayout::META_OFF);   // [2][128]
  float* sArow = reinterpret_cast<float*>(smem + SmemLayout::AROW_OFF);   // [2][NA][256]
  uint8_t* sEpi = smem + SmemLayout::EPI_OFF;                             // [8][32][80]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SmemLayout::BAR_OFF);
  uint64_t* full_bar = bars;                         // [NUM_STAGES]
  uint64_t* empty_bar = bars + NUM_STAGES;           // [NUM_STAGES]
  uint64_t* tfull_bar = bars + 2 * NUM_STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * NUM_STAGES + 2;  // [2]
  uint64_t* w_bar = bars + 2 * NUM_STAGES + 4;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NUM_STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int k = threadIdx.x; k < H; k += NUM_THREADS) {
    sBias[k] = (STAGE <= 2) ? p.bias[k] : 0.f;
    sVec1[k] = p.vec1[k];
    sRed[k] = 0.f;
    sRed2[k] = 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) {
      mbar_init(&full_bar[s], NUM_PROD_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], NUM_EPI_THREADS);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= MMA_WARP && warp < PROD_WARP0) {
    // ===================================================================== weight load + MMA issue
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_MMA_WG));
    if (warp == MMA_WARP && lane == 0) {
      mbar_arrive_expect_tx(w_bar, W_BYTES);
      for (int i = 0; i < W_BYTES / 16384; ++i)
        bulk_g2s(sW + i * 16384, reinterpret_cast<const uint8_t*>(p.Wp) + i * 16384, 16384, w_bar);
      mbar_wait(w_bar, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * H;
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * STAGE_BYTES);
          const uint32_t b_base = smem_u32(sW + kc * (H * KCHUNK * 2));
          if (!(p.dbg & 32))
#pragma unroll
          for (int ks = 0; ks < KCHUNK / UMMA_K; ++ks)
            umma_bf16(d_tmem, umma_desc(a_base + ks * UMMA_K * 2), umma_desc(b_base + ks * UMMA_K * 2),
                      (kc | ks) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[stage]);          // frees the ring slot once these MMAs have read it
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);              // accumulator ready for the epilogue
      }
    }
    __syncwarp();
  } else if (warp >= PROD_WARP0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_PROD));
    // ===================================================================== producers
    // Thread pt owns the 16-byte column chunk (pt & 7) of rows (pt >> 3) + 32 i, i < 4, of every K-chunk.
    // All global loads are software-pipelined: edge indices one tile ahead, the tile's gathers (x, A rows,
    // first B chunk) issued together at the top of the tile, later chunks one K-chunk ahead of their use.
    const int pt = threadIdx.x - 32 * PROD_WARP0;           // 0..255
    const int chunk = pt & 7;
    constexpr int RPT = (TILE_M * 8) / NUM_PROD_THREADS;    // rows per thread per K-chunk (4)
    constexpr int RSTEP = NUM_PROD_THREADS / 8;             // 32
    int stage = 0, it = 0;
    uint32_t phase = 0;
    if constexpr (STAGE == 1) {
      struct TileIdx { int mr, mc, rf, rl, nr[RPT], nc[RPT]; };
      auto load_idx = [&](int tile, TileIdx& t) {
        const int64_t e0 = (int64_t)tile * TILE_M;
        const int64_t rem = p.E - e0;
        const int nvalid = rem < TILE_M ? (int)rem : TILE_M;
        t.mr = t.mc = -1;
        if (pt < nvalid) { t.mr = __ldg(p.row + e0 + pt); t.mc = __ldg(p.col + e0 + pt); }
        t.rf = __ldg(p.row + e0);
        t.rl = __ldg(p.row + e0 + nvalid - 1);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = (pt >> 3) + RSTEP * i;
          t.nr[i] = t.nc[i] = -1;
          if (r < nvalid) { t.nr[i] = __ldg(p.row + e0 + r); t.nc[i] = __ldg(p.col + e0 + r); }
        }
      };
      TileIdx nxt;
      load_idx(blockIdx.x, nxt);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const TileIdx cur = nxt;
        if (tile + (int)gridDim.x < p.num_tiles) load_idx(tile + gridDim.x, nxt);
        const int64_t e0 = (int64_t)tile * TILE_M;
        float* meta = sMeta + (it & 1) * TILE_M;
        float* arow = sArow + (it & 1) * NA * H;
        const int span = cur.rl - cur.rf + 1;
        const bool staged = span <= NA;
        // ---- issue every gather of the tile prologue at once
        float xr[3] = {0.f, 0.f, 0.f}, xc[3] = {0.f, 0.f, 0.f};
        if (cur.mr >= 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            xr[k] = __ldg(p.x + 3 * (int64_t)cur.mr + k);
            xc[k] = __ldg(p.x + 3 * (int64_t)cur.mc + k);
          }
        }
        float4 ar[(NA * H / 4) / NUM_PROD_THREADS];
#pragma unroll
        for (int j = 0; j < (NA * H / 4) / NUM_PROD_THREADS; ++j) {
          const int idx = pt + NUM_PROD_THREADS * j;
          ar[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (staged && idx < span * (H / 4))
            ar[j] = __ldg(reinterpret_cast<const float4*>(p.AB + (int64_t)(cur.rf + idx / (H / 4)) * 2 * H) + idx % (H / 4));
        }
        uint4 pf[2][RPT][2];
        auto issue_b = [&](int kc, int buf) {
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            pf[buf][i][0] = pf[buf][i][1] = make_uint4(0u, 0u, 0u, 0u);
            if (cur.nc[i] >= 0 && !(p.dbg & 1)) {
              const uint4* src = reinterpret_cast<const uint4*>(p.AB + (int64_t)cur.nc[i] * 2 * H + H + kc * KCHUNK + chunk * 8);
              pf[buf][i][0] = __ldg(src);
              pf[buf][i][1] = __ldg(src + 1);
            }
          }
        };
        issue_b(0, 0);
        // ---- publish d2 and the staged A rows
        if (pt < TILE_M) {
          const float dx = xr[0] - xc[0], dy = xr[1] - xc[1], dz = xr[2] - xc[2];
          meta[pt] = dx * dx + dy * dy + dz * dz;
        }
#pragma unroll
        for (int j = 0; j < (NA * H / 4) / NUM_PROD_THREADS; ++j)
          reinterpret_cast<float4*>(arow)[pt + NUM_PROD_THREADS * j] = ar[j];
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_PROD_THREADS) : "memory");
        float d2[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) d2[i] = meta[(pt >> 3) + RSTEP * i];
#pragma unroll
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          if (kc + 1 < NUM_KCHUNKS) issue_b(kc + 1, (kc + 1) & 1);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = sA + stage * STAGE_BYTES;
          const int k0 = kc * KCHUNK + chunk * 8;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = (pt >> 3) + RSTEP * i;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (cur.nr[i] >= 0) {
              float a8[8];
              if (staged) {
                const float4* src = reinterpret_cast<const float4*>(arow + (cur.nr[i] - cur.rf) * H + k0);
                const float4 a0 = src[0], a1 = src[1];
                a8[0] = a0.x; a8[1] = a0.y; a8[2] = a0.z; a8[3] = a0.w;
                a8[4] = a1.x; a8[5] = a1.y; a8[6] = a1.z; a8[7] = a1.w;
              } else {
                load_f32x8(p.AB + (int64_t)cur.nr[i] * 2 * H + k0, a8);
              }
              const uint4 b0 = pf[kc & 1][i][0], b1 = pf[kc & 1][i][1];
              const float b8[8] = {__uint_as_float(b0.x), __uint_as_float(b0.y), __uint_as_float(b0.z),
                                   __uint_as_float(b0.w), __uint_as_float(b1.x), __uint_as_float(b1.y),
                                   __uint_as_float(b1.z), __uint_as_float(b1.w)};
              if (p.out2 && !(p.dbg & 2)) {          // training: keep a and silu'(u) for the backward pass
                float g8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) silu_and_grad(a8[j] + b8[j] + sVec1[k0 + j] * d2[i], a8[j], g8[j]);
                out = pack8(a8);
                *reinterpret_cast<uint4*>(p.out1 + (e0 + r) * H + k0) = out;
                *reinterpret_cast<uint4*>(p.out2 + (e0 + r) * H + k0) = pack8(g8);
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = silu_fast(a8[j] + b8[j] + sVec1[k0 + j] * d2[i]);
                out = pack8(a8);
              }
            }
            *reinterpret_cast<uint4*>(st + sw128_offset(r, chunk)) = out;
          }
          fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core
          mbar_arrive(&full_bar[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else {
      // streams (stages 2-4): one bf16 [E,256] input, no shared metadata, no producer barrier.
      // Loads run PD K-chunks ahead of their use (one register buffer per K-chunk, refilled across tile
      // boundaries) so that ~48-64 KB per SM are in flight -- what HBM latency x bandwidth asks for.
      constexpr int PD = (STAGE == 3) ? 1 : 3;             // stage 3 also carries 64 column-sum registers
      uint4 pf[NUM_KCHUNKS][RPT];
      float rw_cur[RPT], rw_nxt[RPT];                       // stage 3: gw of this thread's rows
      // column sums (stage 3: db5, dW6; stage 4: db2) live in registers for the whole persistent loop
      float cs0[STAGE >= 3 ? NUM_KCHUNKS : 1][8], cs1[STAGE == 3 ? NUM_KCHUNKS : 1][8];
#pragma unroll
      for (int a = 0; a < (STAGE >= 3 ? NUM_KCHUNKS : 1); ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) cs0[a][j] = 0.f;
#pragma unroll
      for (int a = 0; a < (STAGE == 3 ? NUM_KCHUNKS : 1); ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) cs1[a][j] = 0.f;
      auto issue = [&](int tile, int kc) {
        const int64_t e0 = (int64_t)tile * TILE_M;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int64_t e = e0 + (pt >> 3) + RSTEP * i;
          pf[kc][i] = (e < p.E && !(p.dbg & 1)) ? __ldg(reinterpret_cast<const uint4*>(p.in0 + e * H + kc * KCHUNK + chunk * 8))
                                : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      auto issue_rw = [&](int tile) {
        const int64_t e0 = (int64_t)tile * TILE_M;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int64_t e = e0 + (pt >> 3) + RSTEP * i;
          rw_nxt[i] = (STAGE == 3 && e < p.E) ? __ldg(p.ein + e) : 0.f;
        }
      };
      issue_rw(blockIdx.x);
#pragma unroll
      for (int kc = 0; kc < PD; ++kc) issue(blockIdx.x, kc);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int64_t e0 = (int64_t)tile * TILE_M;
        const int next_tile = tile + gridDim.x;
#pragma unroll
        for (int i = 0; i < RPT; ++i) rw_cur[i] = rw_nxt[i];
        if (next_tile < p.num_tiles) issue_rw(next_tile);
#pragma unroll
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc) {
          if (kc + PD < NUM_KCHUNKS) issue(tile, kc + PD);
          else if (next_tile < p.num_tiles) issue(next_tile, kc + PD - NUM_KCHUNKS);
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = sA + stage * STAGE_BYTES;
          const int k0 = kc * KCHUNK + chunk * 8;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            const int r = (pt >> 3) + RSTEP * i;
            const int64_t e = e0 + r;
            uint4 out = make_uint4(0u, 0u, 0u, 0u);
            if (e < p.E) {
              const uint4 in = pf[kc][i];
              float v8[8] = {bf16_lo(in.x), bf16_hi(in.x), bf16_lo(in.y), bf16_hi(in.y),
                             bf16_lo(in.z), bf16_hi(in.z), bf16_lo(in.w), bf16_hi(in.w)};
              if (STAGE == 2) {
                if (p.out2 && !(p.dbg & 2)) {
                  float g8[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) silu_and_grad(v8[j], v8[j], g8[j]);
                  out = pack8(v8);
                  *reinterpret_cast<uint4*>(p.out1 + e * H + k0) = out;
                  *reinterpret_cast<uint4*>(p.out2 + e * H + k0) = pack8(g8);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) v8[j] = silu_fast(v8[j]);
                  out = pack8(v8);
                }
              } else if (STAGE == 3) {
                const float gw = rw_cur[i];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float t, dt;
                  silu_and_grad(v8[j], t, dt);
                  const float gs = gw * sVec1[k0 + j] * dt;
                  cs0[STAGE >= 3 ? kc : 0][j] += gs;
                  cs1[STAGE == 3 ? kc : 0][j] = fmaf(gw, t, cs1[STAGE == 3 ? kc : 0][j]);
                  v8[j] = gs;
                }
                out = pack8(v8);
                if (!(p.dbg & 2)) *reinterpret_cast<uint4*>(p.out1 + e * H + k0) = out;
              } else {
                out = in;
#pragma unroll
                for (int j = 0; j < 8; ++j) cs0[STAGE >= 3 ? kc : 0][j] += v8[j];
              }
            }
            *reinterpret_cast<uint4*>(st + sw128_offset(r, chunk)) = out;
          }
          fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core
          mbar_arrive(&full_bar[stage]);
          if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (STAGE >= 3) {
        // column sums: registers -> (lanes l, l^8, l^16, l^24 share columns) -> shared -> one global atomic
        // per column per CTA
#pragma unroll
        for (int kc = 0; kc < NUM_KCHUNKS; ++kc)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a0 = cs0[STAGE >= 3 ? kc : 0][j];
            a0 += __shfl_xor_sync(0xffffffffu, a0, 8);
            a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
            if (lane < 8) atomicAdd(&sRed[kc * KCHUNK + chunk * 8 + j], a0);
            if (STAGE == 3) {
              float a1 = cs1[STAGE == 3 ? kc : 0][j];
              a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
              a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
              if (lane < 8) atomicAdd(&sRed2[kc * KCHUNK + chunk * 8 + j], a1);
            }
          }
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_PROD_THREADS) : "memory");
        atomicAdd(p.csum0 + pt, sRed[pt]);
        if (STAGE == 3) atomicAdd(p.csum1 + pt, sRed2[pt]);
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ===================================================================== epilogue
    // warp (q, half): TMEM lanes 32q..32q+31 (= tile rows), columns half*128 .. +127, in 4 batches of 32.
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    uint8_t* sc = sEpi + warp * EPI_SCRATCH;
    int it = 0;
    // operand stream of the backward epilogues (dm / da), fetched one batch ahead, across tile boundaries
    uint4 nb[4];
    auto rows_of = [&](int tile) {                 // valid rows of this warp's 32-row slice in `tile`
      const int64_t r0 = (int64_t)tile * TILE_M + q * 32;
      const int64_t left = p.E - r0;
      return left >= 32 ? 32 : (left > 0 ? (int)left : 0);
    };
    if (STAGE >= 3) {
      if (p.dbg & 4) { nb[0] = nb[1] = nb[2] = nb[3] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u); }
      else warp_issue_rows(p.in1 + ((int64_t)blockIdx.x * TILE_M + q * 32) * H + half * 128, lane, rows_of(blockIdx.x), nb);
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const int64_t e0w = (int64_t)tile * TILE_M + q * 32;      // first row of this warp's slice
      const int64_t e = e0w + lane;
      const int nrows = rows_of(tile);
      const bool valid = lane < nrows;
      const int next_tile = tile + gridDim.x;
      float dot = 0.f;
      int dest = -1;
      if ((STAGE == 1 || STAGE == 3) && valid) dest = __ldg(p.row + e);
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const int col0 = half * 128 + cb * 32;
        uint4 dmine[4];
        if (STAGE >= 3) {
          warp_deposit_rows(sc, lane, nb, dmine);
          if (!(p.dbg & 4)) {                      // next batch of the operand stream (next tile after batch 3)
            if (cb + 1 < 4) warp_issue_rows(p.in1 + e0w * H + col0 + 32, lane, nrows, nb);
            else if (next_tile < p.num_tiles)
              warp_issue_rows(p.in1 + ((int64_t)next_tile * TILE_M + q * 32) * H + half * 128, lane, rows_of(next_tile), nb);
          }
        }
        uint32_t raw[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)(acc * H + col0), raw);
        float val[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) val[j] = __uint_as_float(raw[j]) + (STAGE <= 2 ? sBias[col0 + j] : 0.f);
        if (STAGE == 3 && !(p.dbg & 4)) {          // + gagg[dest, col0..col0+31]: mostly one row per warp (broadcast)
          const float4* gsrc = reinterpret_cast<const float4*>(p.nin + (int64_t)(valid ? dest : 0) * H + col0);
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float4 g4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) g4[k] = __ldg(gsrc + 4 * hh + k);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              val[16 * hh + 4 * k] += g4[k].x; val[16 * hh + 4 * k + 1] += g4[k].y;
              val[16 * hh + 4 * k + 2] += g4[k].z; val[16 * hh + 4 * k + 3] += g4[k].w;
            }
          }
        }
        if (STAGE >= 3) {
          const uint32_t dw[16] = {dmine[0].x, dmine[0].y, dmine[0].z, dmine[0].w, dmine[1].x, dmine[1].y, dmine[1].z, dmine[1].w,
                                   dmine[2].x, dmine[2].y, dmine[2].z, dmine[2].w, dmine[3].x, dmine[3].y, dmine[3].z, dmine[3].w};
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            val[2 * j] *= bf16_lo(dw[j]);
            val[2 * j + 1] *= bf16_hi(dw[j]);
          }
          if (STAGE == 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) dot = fmaf(val[j], sVec1[col0 + j], dot);
          }
        }
        if ((STAGE != 2 || p.out0) && !(p.dbg & 8)) {       // v (stage 1), s (stage 2), gv (stage 3), gu (stage 4)
          uint4 packed[4];
          pack32(val, packed);
          warp_store_rows(sc, lane, packed, p.out0 + e0w * H + col0, nrows);
        }
        if (STAGE == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) val[j] = silu_fast(val[j]);
          // segmented sum over the warp's 32 edges: one butterfly transpose-reduce per distinct destination
          uint32_t todo = (p.dbg & 16) ? 0u : __ballot_sync(0xffffffffu, valid);
          while (todo) {
            const int leader = __ffs(todo) - 1;
            const int d0 = __shfl_sync(0xffffffffu, dest, leader);
            const bool mine = valid && dest == d0;
            const uint32_t seg = __ballot_sync(0xffffffffu, mine);
            float w[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = mine ? val[j] : 0.f;
#pragma unroll
            for (int ofs = 16; ofs >= 1; ofs >>= 1) {
              const bool up = (lane & ofs) != 0;
#pragma unroll
              for (int j = 0; j < ofs; ++j) {
                const float keep = up ? w[j + ofs] : w[j];
                const float send = up ? w[j] : w[j + ofs];
                w[j] = keep + __shfl_xor_sync(0xffffffffu, send, ofs);
              }
            }
            atomicAdd(p.nout + (int64_t)d0 * H + col0 + lane, w[0]);   // lane l holds column col0 + l
            todo &= ~seg;
          }
        } else if (STAGE == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) dot = fmaf(silu_fast(val[j]), sVec1[col0 + j], dot);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);               // accumulator stage drained
      if (STAGE == 2 && valid) atomicAdd(p.eout + e, dot + (half == 0 ? __ldg(p.b6) : 0.f));
      if (STAGE == 4 && valid) atomicAdd(p.eout + e, dot);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// fp32 [256,256] (out,in) -> bf16 image of the resident B operand: 4 K-blocks of [256 rows x 128 B],
// 128-byte swizzle.  transpose != 0 packs W^T (for the dgrad GEMMs).
__global__ void pack_weight_kernel(const float* __restrict__ W, int transpose, __nv_bfloat16* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * H) return;
  const int n = idx / H, k = idx % H;
  const float v = transpose ? W[k * H + n] : W[n * H + k];
  const int kb = k / KCHUNK, kl = k % KCHUNK;
  const uint32_t byte = kb * (H * KCHUNK * 2) + sw128_offset(n, kl >> 3) + (kl & 7) * 2;
  out[byte >> 1] = __float2bfloat16(v);
}

template <int STAGE>
static int launch(Params& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(edge_mlp_kernel<STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error(2, "edge_mlp_kernel: %s", cudaGetErrorString(e));
    configured = true;
  }
  p.num_tiles = (int)((p.E + TILE_M - 1) / TILE_M);
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("PEV_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
  }
  const int sms = sm_count();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  edge_mlp_kernel<STAGE><<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
  static const char* names[] = {"", "edge_mlp_kernel<1>", "edge_mlp_kernel<2>", "edge_mlp_kernel<3>", "edge_mlp_kernel<4>"};
  return after_launch(names[STAGE]);
}

}  // namespace tc
}  // namespace pev

using namespace pev;
typedef __nv_bfloat16 bf16_t;

extern "C" {

int pev_pack_weight_bf16(const float* W, int32_t transpose, void* packed, void* stream) {
  PEV_REQUIRE(W && packed, "null argument");
  tc::pack_weight_kernel<<<(tc::H * tc::H + 255) / 256, 256, 0, as_stream(stream)>>>(
      W, transpose, reinterpret_cast<bf16_t*>(packed));
  return after_launch("pack_weight_kernel");
}

int pev_edge_mlp1_fwd_bf16(const float* AB, const float* x, const float* wd, const void* W2p, const float* b2,
                           const int32_t* row, const int32_t* col, int64_t num_nodes, int64_t num_edges,
                           void* v_out, void* a_out, void* da_out, float* agg, void* stream) {
  PEV_REQUIRE(AB && x && wd && W2p && b2 && agg && num_nodes >= 0 && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  if (num_nodes > 0) cudaMemsetAsync(agg, 0, sizeof(float) * tc::H * (size_t)num_nodes, st);
  if (num_edges == 0) return 0;
  PEV_REQUIRE(row && col && v_out, "edge arrays missing");
  PEV_REQUIRE((a_out == nullptr) == (da_out == nullptr), "a_out and da_out go together");
  tc::Params p = {};
  p.AB = AB; p.x = x; p.vec1 = wd; p.row = row; p.col = col; p.nout = agg;
  p.out0 = reinterpret_cast<bf16_t*>(v_out);
  p.out1 = reinterpret_cast<bf16_t*>(a_out);
  p.out2 = reinterpret_cast<bf16_t*>(da_out);
  p.Wp = W2p; p.bias = b2; p.E = num_edges;
  return tc::launch<1>(p, st);
}

int pev_edge_mlp2_fwd_bf16(const void* v, const void* W5p, const float* b5, const float* w6, const float* b6,
                           int64_t num_edges, float* w_out, void* s_out, void* m_out, void* dm_out, void* stream) {
  PEV_REQUIRE(W5p && b5 && w6 && b6 && num_edges >= 0, "bad argument");
  if (num_edges == 0) return 0;
  PEV_REQUIRE(v && w_out, "edge arrays missing");
  PEV_REQUIRE((m_out == nullptr) == (dm_out == nullptr), "m_out and dm_out go together");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(w_out, 0, sizeof(float) * (size_t)num_edges, st);
  tc::Params p = {};
  p.in0 = reinterpret_cast<const bf16_t*>(v);
  p.vec1 = w6; p.b6 = b6; p.eout = w_out;
  p.out0 = reinterpret_cast<bf16_t*>(s_out);
  p.out1 = reinterpret_cast<bf16_t*>(m_out);
  p.out2 = reinterpret_cast<bf16_t*>(dm_out);
  p.Wp = W5p; p.bias = b5; p.E = num_edges;
  return tc::launch<2>(p, st);
}

int pev_edge_mlp2_bwd_bf16(const void* s, const void* dm, const float* gw, const float* w6, const void* W5tp,
                           const float* gagg, const int32_t* row, int64_t num_edges, void* gs_out, void* gv_out,
                           float* db5, float* dw6, void* stream) {
  PEV_REQUIRE(w6 && W5tp && db5 && dw6 && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(db5, 0, sizeof(float) * tc::H, st);
  cudaMemsetAsync(dw6, 0, sizeof(float) * tc::H, st);
  if (num_edges == 0) return 0;
  PEV_REQUIRE(s && dm && gw && gagg && row && gs_out && gv_out, "edge arrays missing");
  tc::Params p = {};
  p.in0 = reinterpret_cast<const bf16_t*>(s);
  p.in1 = reinterpret_cast<const bf16_t*>(dm);
  p.ein = gw; p.nin = gagg; p.row = row; p.vec1 = w6;
  p.out0 = reinterpret_cast<bf16_t*>(gv_out);
  p.out1 = reinterpret_cast<bf16_t*>(gs_out);
  p.csum0 = db5; p.csum1 = dw6;
  p.Wp = W5tp; p.E = num_edges;
  return tc::launch<3>(p, st);
}

int pev_edge_mlp1_bwd_bf16(const void* gv, const void* da, const void* W2tp, const float* wd, int64_t num_edges,
                           void* gu_out, float* gd2, float* db2, void* stream) {
  PEV_REQUIRE(W2tp && wd && db2 && num_edges >= 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  cudaMemsetAsync(db2, 0, sizeof(float) * tc::H, st);
  if (num_edges == 0) return 0;
  PEV_REQUIRE(gv && da && gu_out && gd2, "edge arrays missing");
  cudaMemsetAsync(gd2, 0, sizeof(float) * (size_t)num_edges, st);
  tc::Params p = {};
  p.in0 = reinterpret_cast<const bf16_t*>(gv);
  p.in1 = reinterpret_cast<const bf16_t*>(da);
  p.vec1 = wd;
  p.out0 = reinterpret_cast<bf16_t*>(gu_out);
  p.eout = gd2; p.csum0 = db2;
  p.Wp = W2tp; p.E = num_edges;
  return tc::launch<4>(p, st);
}

}  // extern "C"
