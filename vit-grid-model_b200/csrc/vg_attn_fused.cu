// Fused window / grid attention for sm_100a (maxvit.py:170-219 + the partition / residual code around it, :298-340).
//
// One persistent CTA processes tiles of TWO windows (2 x 64 token slots = the 128 rows of a tcgen05 M=128 MMA).
// Nothing between the residual stream in and the residual stream out touches HBM:
//
//   gather (block/grid partition folded into addressing) + register tokens + LayerNorm + FiLM  -> X tile (fp16, in TMEM:
//                                    the A operand of all 32 QKV projections never touches shared memory)
//   per head h (weights and the per-head tables streamed by TMA, accumulators re-used as operands in TMEM):
//     QKV_h = X * Wqkv_h^T           tcgen05 kind::f16 (fp16 operands: the 10-bit mantissa of tf32 at twice the rate and
//                                    half the shared-memory bytes; X is LayerNorm output, far inside the fp16 range)
//                                    M128 N96  K128                          -> TMEM
//     k RMSNorm in registers: K" = k * (32 gq gk) / |k| (tf32) and V^T (bf16) -> smem; q stays in TMEM, 1/|q| per row
//     S = q K"^T                     tcgen05 kind::tf32  M128 N128 K32, A operand read from TMEM (the q accumulator)
//     S/|q| + relative-position bias (index computed arithmetically), masked softmax in registers;
//     P / rowsum (bf16) -> TMEM (tcgen05.st, 64 columns of bf16 pairs over the q | k columns of the NEXT head's QKV buffer,
//                                    dead by then; the S accumulator is free for the next head as soon as the softmax warps
//                                    hold S in registers)
//     O_h = P V                      tcgen05 kind::f16   M128 N32  K128, A operand = P read from TMEM -> TMEM
//     Out += O_h * Wout_h^T          tcgen05 kind::tf32  M128 N128 K32, A operand read from TMEM (the O accumulator),
//                                    accumulated over heads
//   epilogue: Out + residual, scattered back through the inverse partition map; register-token rows to reg_out.
//
// Warp roles: warp 0 = TMA (weights, tables), warp 1 = MMA issuer, warps 2..9 = softmax warps (two threads per token row =
// TMEM lane; they also gather / LayerNorm the tile and write it back), warps 10..17 = staging warps (K" / V^T operands and
// 1/|q| of the next head, concurrently with the softmax of the current one).  All operand tiles written by threads use the same K-major SWIZZLE_128B layout TMA produces.
#include <stdlib.h>

#include "vg_common.cuh"
#include "vg_host.h"
#include "vg_rng.cuh"

namespace vg {

namespace fa {
constexpr int C = 128;         // model channels (K of the QKV projection)
constexpr int DH = 32;         // head dim
constexpr int WIN = 7, REG = 4, SEQ = REG + WIN * WIN;   // the kernel is specialised for 7x7 windows + 4 register tokens
// shared memory map (bytes); every operand tile is 1024-B aligned
constexpr int WQ_BYTES = 2 * 12288;              // fp16: 2 k-blocks x [96 rows x 128 B]
constexpr int WQ_OFF = 0;                        // 2 buffers (heads alternate)
constexpr int WO_OFF = WQ_OFF + 2 * WQ_BYTES;    // tf32 [128 rows x 128 B], 2 buffers
constexpr int R1_BYTES = 16384;                  // K" operand: tf32 [128 keys x 128 B]
// K" / V^T / 1/|q| are written by the staging warps two heads ahead of PV: three buffers (head % 3), so that staging head h+2
// does not wait for PV(h) to release the buffers of head h
constexpr int NOPB = 3;
constexpr int R1_OFF = WO_OFF + 2 * 16384;
constexpr int VT_OFF = R1_OFF + NOPB * R1_BYTES; // NOPB x [2 k-blocks x 32 rows x 128 B]
// per-head table: shifted bias rows [bi][row 0..12] (row stride TAB_SR, bi stride TAB_SB floats: with these strides the
// 16-byte reads of the eight tokens of a quarter-warp fall into different banks, 268 wavefronts per head instead of 588 for
// the dense [7][13][8] layout) | table[169] x 8 | 32*gq*gk [32] | unused [32]
constexpr int TAB_SR = 12, TAB_SB = 180, TAB_T169 = 7 * TAB_SB;
constexpr int TAB_FLOATS = TAB_T169 + 8 + 64;
constexpr int TAB_OFF = VT_OFF + NOPB * 8192;    // 2 x TAB_FLOATS floats
constexpr int RED_OFF = TAB_OFF + 2 * TAB_FLOATS * 4;    // softmax pair exchange: 2 (head parity) x float[128][2][2] (max, sum); then LN sum / sq-sum float[128][2] each
constexpr int QINV_OFF = RED_OFF + 6 * 1024;             // NOPB x float[128]: 1/|q| per row (staging warps -> softmax warps)
constexpr int BAR_OFF = QINV_OFF + NOPB * 512;
constexpr int SMEM_BYTES = BAR_OFF + 256 + 1024;          // + barriers + alignment slack
// warp 0 TMA, warp 1 MMA, warps 2..9 softmax, warps 10..17 operand staging + tile prologue / epilogue (2 threads per token row
// each).  (A warpgroup-aligned 640-thread layout with setmaxnreg was tried: ptxas kept every role within the launch-time 96
// registers, so it only added spills.)
constexpr int THREADS = 576;
// TMEM columns (all 512 in use)
constexpr int T_QKV0 = 0;      // 96   q | k | v accumulators of even heads
constexpr int T_QKV1 = 96;     // 96   odd heads
// The buffer of head h+1 is dead between the S product of head h+1 (issued a head ahead of its softmax) and the QKV projection
// of head h+3: in that window its q | k columns hold P(h) (64 columns of bf16 pairs, the A operand of PV(h)) and its v columns
// O_h (the D of PV(h), the A operand of out(h)).
constexpr int T_PQ = 0;        // P inside the other-parity buffer
constexpr int T_OV = 64;       // O_h inside the other-parity buffer
constexpr int T_X = 192;       // 64   X tile: 128 channels as fp16 pairs (A operand of every QKV projection of the tile)
constexpr int T_S = 256;       // 128
constexpr int T_OUT = 384;     // 128
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace fa

struct FusedAttnParams {
  const float* x; float* x_out;
  const float* reg_in; int reg_per_field; float* reg_out;
  const float* film;                   // (N, 2C) gamma | beta
  const float* head_tab;               // [heads][TAB_FLOATS] (see pack_head_tables in maxvit.py)
  AttnGeom g;
  int heads;
  float ln_eps;
  long long n_windows;
  long long* dbg;                      // optional [heads][8] clock64 stamps of CTA 0 / compute thread 0 / first tile
  DropCfg drop;                        // training: dropout on the probabilities and on the to_out output
};

// byte offset of 16-byte chunk `c16` of row `r` inside a [rows x 128 B] K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128(int r, int c16) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t a, unsigned short v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(v) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
// packed fp32 pairs (FADD2 / FMUL2 on sm_100): half the issue slots of the softmax's elementwise passes
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(reinterpret_cast<uint64_t&>(d)) : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand is an fp32 accumulator read in place as tf32 (lanes = rows, one column per k)
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T with 16-bit operands: A = packed bf16 pairs written with tcgen05.st (one 32-bit column per two k)
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 16 registers -> 32 lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// plain (non-tensor) TMA copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// named barriers: 1 = all 256 compute threads, 2..5 = the two warps that share a TMEM lane group
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void pair_sync(int lg) { asm volatile("bar.sync %0, 64;" ::"r"(2 + lg) : "memory"); }

// residual-stream row of tile row t (= token slot i of window `half` of the tile): block / grid partition and the register
// tokens folded into the address (maxvit.py:298 / :322 / :305)
struct TokRow { const float* src; long long pix; long long wdx; int n; };
__device__ __forceinline__ TokRow tok_row(const FusedAttnParams& p, long long tile, int t) {
  using namespace fa;
  const int half = t >> 6, i = t & 63;
  const AttnGeom& g = p.g;
  const int nwin = g.nwin();
  TokRow r;
  r.wdx = tile * 2 + half;
  const bool win_valid = r.wdx < p.n_windows;
  r.n = win_valid ? (int)(r.wdx / nwin) : 0;
  const int wi = win_valid ? (int)(r.wdx - (long long)r.n * nwin) : 0;
  r.src = nullptr; r.pix = -1;
  if (win_valid && i < SEQ) {
    if (i < REG) r.src = p.reg_in + (p.reg_per_field ? (long long)r.n * REG * C : 0) + (long long)i * C;
    else {
      r.pix = (long long)r.n * g.Hl * g.Wl + attn_token_pixel(g, wi, i - REG);     // maxvit.py:298 / :322
      r.src = p.x + r.pix * C;
    }
  }
  return r;
}
__device__ __forceinline__ void pair_sync2(int lg) { asm volatile("bar.sync %0, 64;" ::"r"(6 + lg) : "memory"); }   // staging-warp pairs

// DROP: training instantiation with the dropout code; the inference instantiation is a quarter smaller (the kernel's cold paths
// -- tile prologue / epilogue -- run from an instruction cache shared with four hot role loops)
template <bool DROP>
__global__ void __launch_bounds__(fa::THREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap mapWq, const __grid_constant__ CUtensorMap mapWo,
                  const FusedAttnParams p) {
  using namespace fa;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
  uint64_t* wq_full = bars + 0;  uint64_t* wq_free = bars + 2;    // [2] each: one per QKV weight buffer
  uint64_t* wo_full = bars + 4;  uint64_t* wo_free = bars + 6;    // [2] each
  uint64_t* x_ready = bars + 8;
  uint64_t* qkv_done = bars + 9;                 // [2]
  uint64_t* qk_ready = bars + 11;                // [2]: one per operand buffer
  uint64_t* s_done = bars + 13;
  uint64_t* p_ready = bars + 14;
  uint64_t* pv_done = bars + 15;                 // [2]  R1[r] / VT[r] no longer read by MMAs
  uint64_t* tile_done = bars + 17; uint64_t* out_free = bars + 18;
  uint64_t* tab_full = bars + 19;                // [2]
  uint64_t* tab_free = bars + 21;                // [2]
  uint64_t* s_free = bars + 23;                  // the softmax warps hold S in registers: the next S product may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int heads = p.heads;

  if (warp == 1 && lane == 0) {
    mbar_init(wq_full + 0, 1); mbar_init(wq_full + 1, 1); mbar_init(wq_free + 0, 1); mbar_init(wq_free + 1, 1);
    mbar_init(wo_full + 0, 1); mbar_init(wo_full + 1, 1); mbar_init(wo_free + 0, 1); mbar_init(wo_free + 1, 1);
    mbar_init(s_free, 8);
    mbar_init(x_ready, 8);
    mbar_init(qkv_done + 0, 1); mbar_init(qkv_done + 1, 1);
    mbar_init(qk_ready + 0, 8); mbar_init(qk_ready + 1, 8); mbar_init(s_done, 1); mbar_init(p_ready, 8);
    mbar_init(pv_done + 0, 1); mbar_init(pv_done + 1, 1);
    mbar_init(tile_done, 1); mbar_init(out_free, 8);
    mbar_init(tab_full + 0, 1); mbar_init(tab_full + 1, 1); mbar_init(tab_free + 0, 8); mbar_init(tab_free + 1, 8);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (lane == 0) { tma_prefetch_desc(&mapWq); tma_prefetch_desc(&mapWo); }
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);        // warp-uniform: lets the MMA issue use uniform registers (no per-MMA elect/broadcast loop)

  const long long n_tiles = (p.n_windows + 1) / 2;

  if (warp == 0) {
    // ============================== TMA: stream the per-head weights and bias tables ==============================
    // Program order follows the order in which the MMA warp needs the data: the QKV weights run two heads ahead of the
    // out-projection weights (a load that waited for out(j-1) in front of WQ(j+2) would dead-lock the MMA warp, which issues
    // QKV(j+2) before out(j)).
    if (lane == 0) {
      long long my_tiles = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) ++my_tiles;
      const long long total = my_tiles * heads;
      auto load_wq = [&](long long j) {
        const int h = (int)(j % heads);
        const uint32_t b = (uint32_t)(j & 1);
        mbar_wait_tag(wq_free + b, (uint32_t)(((j >> 1) & 1) ^ 1), 172);    // QKV(j-2) has read this buffer
        mbar_arrive_expect_tx(wq_full + b, WQ_BYTES);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) tma_load_2d(smem + WQ_OFF + b * WQ_BYTES + kb * 12288, &mapWq, wq_full + b, kb * 64, h * 96);
      };
      auto load_tab = [&](long long j) {
        const int h = (int)(j % heads);
        const uint32_t r = (uint32_t)(j & 1);
        mbar_wait_tag(tab_free + r, (uint32_t)(((j >> 1) & 1) ^ 1), 180);   // the compute warps are done with the tables of head j-2
        mbar_arrive_expect_tx(tab_full + r, TAB_FLOATS * 4);
        bulk_load(smem + TAB_OFF + r * TAB_FLOATS * 4, p.head_tab + (long long)h * TAB_FLOATS, TAB_FLOATS * 4, tab_full + r);
      };
      // need-order of the MMA warp: QKV(0..2) at the start, then per head j: out(j) [WO(j)] followed by QKV(j+3) [WQ(j+3)].
      // Every load only waits for MMAs that were issued before the ones that need it (WO(j): out(j-2); WQ(j+3): QKV(j+1);
      // WQ(2): QKV(0), which needs nothing but WQ(0) and the X tile).
      for (long long j = 0; j < 3 && j < total; ++j) { load_wq(j); if (j < 2) load_tab(j); }
      for (long long j = 0; j < total; ++j) {
        const uint32_t b = (uint32_t)(j & 1);
        mbar_wait_tag(wo_free + b, (uint32_t)(((j >> 1) & 1) ^ 1), 187);
        mbar_arrive_expect_tx(wo_full + b, 16384);
        tma_load_2d(smem + WO_OFF + b * 16384, &mapWo, wo_full + b, 0, (int)(j % heads) * 128);
        if (j + 3 < total) load_wq(j + 3);
        if (j + 2 < total) load_tab(j + 2);
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // The whole warp runs this role converged (all lanes poll the barriers); the tcgen05 instructions of a batch are issued
    // by one elected lane.  Written this way nvcc emits the UTCHMMAs of a batch back to back; a `lane == 0` branch makes it
    // wrap each one in an ELECT / BRA.U.ANY loop.
    {
      constexpr uint32_t id_qkv = umma_idesc_f16(128, 96);
      constexpr uint32_t id_s = umma_idesc_tf32(128, 128);
      constexpr uint32_t id_pv = umma_idesc_bf16(128, 32);
      constexpr uint32_t id_out = umma_idesc_tf32(128, 128);
      const uint32_t sWQ = smem_u32(smem + WQ_OFF), sWO = smem_u32(smem + WO_OFF);
      uint32_t it = 0, tl = 0;                               // global head counter, tile counter
      auto issue_qkv = [&](uint32_t hh) {
        const uint32_t b = hh & 1;
        mbar_wait_tag(wq_full + b, (hh >> 1) & 1, 203);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem + (b ? T_QKV1 : T_QKV0);
          const uint64_t db = umma_desc_k128(sWQ + b * WQ_BYTES);
#pragma unroll
          for (int st = 0; st < 8; ++st) {                         // A = X from TMEM: 8 columns (16 fp16) per K step
            const int kb = st >> 2, k = st & 3;
            tc_mma_bf16_ts(d, tmem + T_X + 8 * st, db + kb * (12288 >> 4) + 2 * k, id_qkv, st ? 1u : 0u);
          }
          tc_commit(qkv_done + b);
          tc_commit(wq_free + b);
        }
        __syncwarp();
      };
      auto issue_s = [&](uint32_t hh, bool first) {          // S = q K"^T, A = the q accumulator of head hh read from TMEM
        const uint32_t r = hh & 1;
        if (!first) mbar_wait_tag(s_free, (hh - 1) & 1, 225);   // the softmax warps hold S(hh-1) in registers
        mbar_wait_tag(qk_ready + r, (hh >> 1) & 1, 224);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t ta = tmem + (r ? T_QKV1 : T_QKV0);
          const uint64_t db = umma_desc_k128(smem_u32(smem + R1_OFF + (hh % NOPB) * R1_BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma_tf32_ts(tmem + T_S, ta + 8 * k, db + 2 * k, id_s, k ? 1u : 0u);
          tc_commit(s_done);
        }
        __syncwarp();
      };
      // Issue order per head h:  [p_ready(h)] PV(h)   [pv_done(h)] out(h)   [s_free(h+1), qk_ready(h+2)] S(h+2)   QKV(h+3)
      // The S product runs one head ahead of the softmax (S(h+1) is complete before the softmax of head h ends), the QKV
      // projection three heads ahead: QKV(h+3) re-uses the TMEM buffer of head h+1, whose q columns were read by S(h+1) and
      // whose v columns hold O_h between PV(h) and out(h) (the tensor pipe executes in issue order).
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        mbar_wait_tag(x_ready, tl & 1, 216);
        tc_fence_after();
        issue_qkv(it);
        issue_qkv(it + 1);
        issue_s(it, true);                                   // (the previous tile's last p_ready implied its s_free)
        issue_qkv(it + 2);
        issue_s(it + 1, false);
        for (int h = 0; h < heads; ++h, ++it) {
          const uint32_t r = it & 1;
          const uint32_t sVT = smem_u32(smem + VT_OFF + (it % NOPB) * 8192);
          long long* md = (p.dbg && blockIdx.x == 0 && tl == 0 && lane == 0) ? p.dbg + (heads + h) * 8 : nullptr;   // MMA-warp time stamps
          if (md) md[0] = clock64();
          const bool do_qkv = h + 3 < heads;
          // the weights arrive a head early: poll their barriers while this warp would idle on p_ready anyway
          mbar_wait_tag(wo_full + r, (it >> 1) & 1, 246);
          if (h == 0) mbar_wait_tag(out_free, (tl & 1) ^ 1, 247);     // previous tile's epilogue has drained Out
          if (do_qkv) mbar_wait_tag(wq_full + (r ^ 1), ((it + 3) >> 1) & 1, 203);
          if (md) md[1] = clock64();
          mbar_wait_tag(p_ready, it & 1, 236);               // softmax(h) done: P(h) is in TMEM
          tc_fence_after();
          if (md) md[2] = clock64();
          const uint32_t tqn = tmem + (r ? T_QKV0 : T_QKV1);       // buffer of heads h+1 / h+3
          const uint32_t tO = tqn + T_OV;
          if (elect_one()) {
            const uint64_t dvt = umma_desc_k128(sVT);
#pragma unroll
            for (int st = 0; st < 8; ++st) {                       // O = P V  (A = P from TMEM)
              const int kb = st >> 2, k = st & 3;
              tc_mma_bf16_ts(tO, tqn + T_PQ + 8 * st, dvt + kb * (4096 >> 4) + 2 * k, id_pv, st ? 1u : 0u);
            }
            tc_commit(pv_done + r);
          }
          __syncwarp();
          if (md) md[3] = clock64();
          // out(h) reads O_h as its A operand from TMEM: the tensor pipe orders accumulation into the same columns, NOT a TMEM
          // operand read behind the write of the previous instruction -- wait for PV(h) to retire (back to back, the
          // out-projection read a partly written O_h: non-deterministic results)
          mbar_wait_tag(pv_done + r, (it >> 1) & 1, 248);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t dwo = umma_desc_k128(sWO + r * 16384);
#pragma unroll
            for (int st = 0; st < 4; ++st) tc_mma_tf32_ts(tmem + T_OUT, tO + 8 * st, dwo + 2 * st, id_out, (h | st) ? 1u : 0u);   // Out += O_h Wout_h^T  (A = O_h read from TMEM)
            tc_commit(wo_free + r);
          }
          __syncwarp();
          if (md) md[4] = clock64();
          // S(h+2) before the long QKV projection: the softmax warps wait for it twice (before P(h+1) is written over the q
          // columns it reads, and at the start of head h+2)
          if (h + 2 < heads) issue_s(it + 2, false);
          if (md) md[5] = clock64();
          if (elect_one()) {
            if (do_qkv) {                                          // QKV(h+3) (overwrites P(h), O_h: issued after PV(h), out(h))
              const uint64_t dwq = umma_desc_k128(sWQ + (r ^ 1) * WQ_BYTES);
#pragma unroll
              for (int st = 0; st < 8; ++st) {
                const int kb = st >> 2, k = st & 3;
                tc_mma_bf16_ts(tqn, tmem + T_X + 8 * st, dwq + kb * (12288 >> 4) + 2 * k, id_qkv, st ? 1u : 0u);
              }
              tc_commit(qkv_done + (r ^ 1));
              tc_commit(wq_free + (r ^ 1));
            }
          }
          __syncwarp();
          if (md) { md[6] = clock64(); md[7] = md[6]; }
        }
        if (elect_one()) tc_commit(tile_done);
        __syncwarp();
      }
    }
  } else if (warp >= 10) {
    // ============================== staging warps: QKV accumulator -> MMA operands, one head ahead of the softmax ==============================
    const int lg = warp & 3;                                 // TMEM lane group this warp may access
    const int ch = (warp - 10) >> 2;                         // 0: q norm + K" operand, 1: V^T operand
    const int t = lg * 32 + lane;                            // tile row == TMEM lane
    const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16);
    const uint32_t s_base = smem_u32(smem);
    float* qinv = reinterpret_cast<float*>(smem + QINV_OFF);
    float* lnred = reinterpret_cast<float*>(smem + RED_OFF) + 1024;   // LN sum [128][2] | LN sq-sum [128][2], behind the softmax exchange buffers
    // ---------------- tile prologue: gather + LayerNorm + FiLM -> X tile (fp16, TMEM).  Run here, by the warps that have slack:
    // the X tile of the NEXT tile is built while the softmax warps are still on the last heads of this one (every QKV
    // projection of this tile has retired once its last head is staged). ----------------
    auto build_x = [&](long long tile) {
      const TokRow tr = tok_row(p, tile, t);
      const float* src = tr.src;
      // this thread owns channels [ch*64, +64) of the row; the row statistics are exchanged within the pair
      float4 v[16];
      float sm = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        v[c] = src ? *reinterpret_cast<const float4*>(src + ch * 64 + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        sm += v[c].x + v[c].y + v[c].z + v[c].w;
      }
      lnred[t * 2 + ch] = sm;
      pair_sync2(lg);
      const float mean = (lnred[t * 2] + lnred[t * 2 + 1]) * (1.0f / C);
      float ss = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        v[c].x -= mean; v[c].y -= mean; v[c].z -= mean; v[c].w -= mean;
        ss += v[c].x * v[c].x + v[c].y * v[c].y + v[c].z * v[c].z + v[c].w * v[c].w;
      }
      lnred[256 + t * 2 + ch] = ss;
      pair_sync2(lg);
      const float rstd = rsqrtf((lnred[256 + t * 2] + lnred[256 + t * 2 + 1]) * (1.0f / C) + p.ln_eps);
      const float* film = p.film + (long long)tr.n * 2 * C + ch * 64;
      // fp16 operand tile in TMEM: this thread's 64 channels are columns [ch*32, +32) of its lane (two channels per column)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          pk[2 * c] = 0u; pk[2 * c + 1] = 0u;
          if (src) {
            const float4 xv = v[hf * 8 + c];
            const float4 ga = *reinterpret_cast<const float4*>(film + (hf * 8 + c) * 4), be = *reinterpret_cast<const float4*>(film + C + (hf * 8 + c) * 4);
            pk[2 * c] = pack_f16(xv.x * rstd * ga.x + be.x, xv.y * rstd * ga.y + be.y);
            pk[2 * c + 1] = pack_f16(xv.z * rstd * ga.z + be.z, xv.w * rstd * ga.w + be.w);
          }
        }
        tmem_st16(lane_addr + T_X + ch * 32 + hf * 16, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_ready);
    };
    // ---------------- tile epilogue: Out + residual, inverse partition (this thread: channels [ch*64, +64)).  Also run here: the
    // softmax warps go straight on to the next tile's heads. ----------------
    auto epilogue = [&](long long tile, uint32_t tl) {
      const TokRow tr = tok_row(p, tile, t);
      const int i = t & 63;
      mbar_wait_tag(tile_done, tl & 1, 472);
      tc_fence_after();
      float* dst = nullptr;
      if (tr.src) {
        if (i < REG) dst = p.reg_out ? p.reg_out + (tr.wdx * REG + i) * C : nullptr;
        else dst = p.x_out + tr.pix * C;
      }
      float v[32];
#pragma unroll 1
      for (int q = 0; q < 2; ++q) {
        const int c0 = ch * 64 + q * 32;
        tmem_ld32(lane_addr + T_OUT + c0, v); tmem_wait_ld();
        if (DROP && p.drop.thresh) {                                   // nn.Dropout after to_out (maxvit.py:151)
          const uint32_t rid = drop_row(tr.wdx, i);
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const uint32_t hsh = drop_hash(p.drop.seed, rid, drop_group_out(p.drop.salt, (c0 + c) >> 2));
#pragma unroll
            for (int k = 0; k < 4; ++k) v[c + k] *= (int)((hsh >> (8 * k)) & 255u) >= p.drop.thresh ? p.drop.scale : 0.f;
          }
        }
        if (dst) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(tr.src + c0 + c));
            *reinterpret_cast<float4*>(dst + c0 + c) = make_float4(v[c] + rr.x, v[c + 1] + rr.y, v[c + 2] + rr.z, v[c + 3] + rr.w);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(out_free);
    };
    uint32_t itx = 0, tl = 0;
    long long ep_tile = -1;                                  // tile whose epilogue is pending
    if ((long long)blockIdx.x < n_tiles) build_x(blockIdx.x);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      for (int hx = 0; hx < heads; ++hx, ++itx) {
        // the previous tile's epilogue runs once the first three heads of this tile are staged (the MMA warp issued their QKV
        // projections together; the fourth follows a softmax later): the pipeline refill never waits for it
        if (hx == 3 && ep_tile >= 0) { epilogue(ep_tile, tl - 1); ep_tile = -1; }
        const uint32_t r = itx & 1;
        const uint32_t ob = itx % NOPB;                        // operand buffer of this head
        const uint32_t R1 = s_base + R1_OFF + ob * R1_BYTES;
        const uint32_t VT = s_base + VT_OFF + ob * 8192;
        float4 gm[8];                                            // 32 * gamma_q * gamma_k of this head: in flight during the wait
        if (ch == 0) {
          const float4* ksc = reinterpret_cast<const float4*>(p.head_tab + (long long)hx * TAB_FLOATS + TAB_T169 + 8);
#pragma unroll
          for (int c = 0; c < 8; ++c) gm[c] = __ldg(ksc + c);
        }
        mbar_wait_tag(qkv_done + r, (itx >> 1) & 1, 352);
        tc_fence_after();
        const uint32_t tq = lane_addr + (r ? T_QKV1 : T_QKV0);
        if (ch == 0) {
          float v[32], w[32];
          tmem_ld32(tq, v);                                              // q
          tmem_ld32(tq + 32, w);                                         // k
          tmem_wait_ld();
          float nq4[4] = {0.f, 0.f, 0.f, 0.f}, nk4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int d = 0; d < 32; ++d) { nq4[d & 3] = fmaf(v[d], v[d], nq4[d & 3]); nk4[d & 3] = fmaf(w[d], w[d], nk4[d & 3]); }
          const float nq = (nq4[0] + nq4[1]) + (nq4[2] + nq4[3]), nk = (nk4[0] + nk4[1]) + (nk4[2] + nk4[3]);
          const float inv_q = 1.0f / fmaxf(sqrtf(nq), 1e-12f);           // F.normalize(eps=1e-12)  (maxvit.py:30)
          const float inv_k = 1.0f / fmaxf(sqrtf(nk), 1e-12f);
          // K" / 1/|q| of head itx-3 (same buffer) are no longer in use: QKV(itx) was issued behind S(itx-3) and PV(itx-3), so
          // qkv_done(itx) implies both retired, and the softmax of head itx-3 ended before PV(itx-3) was issued
          qinv[ob * 128 + t] = inv_q;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            sts128(R1 + sw128(t, c), w[4 * c] * inv_k * gm[c].x, w[4 * c + 1] * inv_k * gm[c].y, w[4 * c + 2] * inv_k * gm[c].z, w[4 * c + 3] * inv_k * gm[c].w);
          }
        } else {
          float w[32];
          tmem_ld32(tq + 64, w);                                         // v
          tmem_wait_ld();
          // V^T of head itx-3 (same buffer): PV(itx-3) was issued before QKV(itx), so qkv_done(itx) implies it retired
          // V^T (bf16): element (d, key t) of a [32 x 128] K-major tile, 2 k-blocks of 64 keys
          const uint32_t vt = VT + (t >> 6) * 4096 + (t & 7) * 2;
          const int kc = (t & 63) >> 3;
#pragma unroll
          for (int d = 0; d < 32; ++d) sts16(vt + sw128(d, kc), __bfloat16_as_ushort(__float2bfloat16(w[d])));
        }
        tc_fence_before();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(qk_ready + r);
      }
      // the last head of this tile is staged: every QKV projection that reads the X tile has retired
      if (tile + gridDim.x < n_tiles) { tc_fence_after(); build_x(tile + gridDim.x); }
      ep_tile = tile;
    }
    if (ep_tile >= 0) epilogue(ep_tile, tl - 1);
  } else {
    // ============================== softmax warps: two threads per token row ==============================
    const int cw = warp - 2;                                 // 0..7
    const int lg = warp & 3;                                 // TMEM lane group this warp may access
    const int ch = cw >> 2;                                  // column half handled by this thread
    const int ctid = cw * 32 + lane;                         // 0..255
    const int t = lg * 32 + lane;                            // tile row == TMEM lane
    const int half = t >> 6, i = t & 63;                     // window within the pair, token slot
    const uint32_t lane_addr = tmem + ((uint32_t)(lg * 32) << 16);
    const bool is_reg = i < REG;
    // window position of this token (clamped for register / pad slots so that table addresses stay valid)
    const int ti = (i >= REG && i < SEQ) ? i - REG : 0;
    const int ai = ti / WIN, bi = ti - ai * WIN;
    const uint32_t s_base = smem_u32(smem);
    float* red = reinterpret_cast<float*>(smem + RED_OFF);   // [2 head parities][128 rows][2 threads] x (max, sum)
    const float* qinv = reinterpret_cast<const float*>(smem + QINV_OFF);
    // bias rows of this token: window tokens step one 32-byte row back per key row aj; register-token rows (maxvit.py:167:
    // one shared bias for every key) read the constant row, with step 0 -- no per-element select
    const uint32_t b_off = is_reg ? (uint32_t)TAB_T169 * 4u : (uint32_t)(bi * TAB_SB + (ai + 6) * TAB_SR) * 4u;
    const uint32_t b_step = is_reg ? 0u : (uint32_t)TAB_SR * 4u;
    uint32_t it = 0, tl = 0;
    uint32_t my_heads = 0;                                   // heads this CTA processes over all its tiles
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) my_heads += (uint32_t)heads;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      // (the X tile of this tile is built by the staging warps)
      const long long wdx = tile * 2 + half;                 // window of this row (dropout row id)

      for (int h = 0; h < heads; ++h, ++it) {
        const bool dbg = p.dbg && blockIdx.x == 0 && ctid == 0 && tl == 0;
        if (dbg) p.dbg[h * 8 + 0] = clock64();
        const uint32_t r = it & 1;
        const uint32_t tab = s_base + TAB_OFF + r * TAB_FLOATS * 4;
        if (it == 0) mbar_wait_tag(tab_full + 0, 0, 394);                         // per-head bias table (TMA); later heads: waited for below

        // ---------------- S half-row: / |q|, + bias, masked softmax -> normalised P (bf16) ----------------
        if (h == 0) mbar_wait_tag(s_done, it & 1, 398);       // later heads: observed during the previous head (before its P store)
        tc_fence_after();
        if (dbg) p.dbg[h * 8 + 1] = clock64();
        {
          float2 sc2[16];
          float* sc = reinterpret_cast<float*>(sc2);
          tmem_ld32(lane_addr + T_S + half * 64 + ch * 32, sc);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);                            // S(h+1) may now overwrite the accumulator
          if (dbg) p.dbg[h * 8 + 2] = clock64();
          // logits in the exp2 domain: S * (log2e / |q|) + bias * log2e (the table is stored pre-multiplied)
          const float cq = qinv[(it % NOPB) * 128 + t] * LOG2E;                    // written by the staging warps before qk_ready -> S -> s_done
          const uint32_t brow = tab + b_off;                             // register-token rows read the constant row (b_step = 0)
          float m = -INFINITY;
          if (ch == 0) {
            // keys 0..3 are register tokens, keys 4..31 are window rows aj = 0..3
            const float t169 = reinterpret_cast<const float*>(smem + TAB_OFF)[r * TAB_FLOATS + TAB_T169];
#pragma unroll
            for (int j = 0; j < 4; ++j) sc[j] = fmaf(sc[j], cq, t169);
#pragma unroll
            for (int aj = 0; aj < 4; ++aj) {
              const uint32_t a = brow - aj * b_step;
              const float4 b0 = lds128(a), b1 = lds128(a + 16);
              float* q = sc + 4 + aj * 7;
              q[0] = fmaf(q[0], cq, b0.x); q[1] = fmaf(q[1], cq, b0.y); q[2] = fmaf(q[2], cq, b0.z); q[3] = fmaf(q[3], cq, b0.w);
              q[4] = fmaf(q[4], cq, b1.x); q[5] = fmaf(q[5], cq, b1.y); q[6] = fmaf(q[6], cq, b1.z);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, sc[j]);
          } else {
            // keys 32..52 are window rows aj = 4..6; keys 53..63 are padding
#pragma unroll
            for (int aj = 4; aj < 7; ++aj) {
              const uint32_t a = brow - aj * b_step;
              const float4 b0 = lds128(a), b1 = lds128(a + 16);
              float* q = sc + (aj - 4) * 7;
              q[0] = fmaf(q[0], cq, b0.x); q[1] = fmaf(q[1], cq, b0.y); q[2] = fmaf(q[2], cq, b0.z); q[3] = fmaf(q[3], cq, b0.w);
              q[4] = fmaf(q[4], cq, b1.x); q[5] = fmaf(q[5], cq, b1.y); q[6] = fmaf(q[6], cq, b1.z);
            }
#pragma unroll
            for (int j = 0; j < 21; ++j) m = fmaxf(m, sc[j]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(tab_free + r);                      // last read of this head's tables
          // Barriers that completed long ago are polled here, where the thread has independent work in flight, instead of at
          // the hand-over between heads: the next head's table (loaded two heads ahead).
          const bool tab_next = it + 1 < my_heads;
          const bool tab_ok = tab_next ? mbar_try_wait(tab_full + (r ^ 1), ((it + 1) >> 1) & 1) : true;   // result used at the end of the head
          // ONE exchange per head: every thread exponentiates against the maximum of its OWN half row; the pair then swaps
          // (max, sum) and rescales by 2^(own max - row max) together with the normalisation (the same softmax, exactly)
          if (dbg) { p.dbg[h * 8 + 3] = clock64(); p.dbg[h * 8 + 4] = p.dbg[h * 8 + 3]; }
          const float2 nm = make_float2(-m, -m);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          if (ch == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              sc2[k] = fadd2(sc2[k], nm);
              sc2[k].x = ex2(sc2[k].x); sc2[k].y = ex2(sc2[k].y);
              if (k & 1) acc1 = fadd2(acc1, sc2[k]); else acc0 = fadd2(acc0, sc2[k]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 10; ++k) {
              sc2[k] = fadd2(sc2[k], nm);
              sc2[k].x = ex2(sc2[k].x); sc2[k].y = ex2(sc2[k].y);
              if (k & 1) acc1 = fadd2(acc1, sc2[k]); else acc0 = fadd2(acc0, sc2[k]);
            }
            sc[20] = ex2(sc[20] - m); sc[21] = 0.f;
            acc0 = fadd2(acc0, sc2[10]);
#pragma unroll
            for (int k = 11; k < 16; ++k) sc2[k] = make_float2(0.f, 0.f);
          }
          acc0 = fadd2(acc0, acc1);
          const float s_own = acc0.x + acc0.y;
          *reinterpret_cast<float2*>(red + (r * 128 + t) * 4 + ch * 2) = make_float2(m, s_own);
          // P(h) goes over the q | k columns of head h+1's buffer: S(h+1), which reads that q, was issued when this softmax
          // released the S accumulator (s_free) and has long retired
          const bool sd_next = h + 1 < heads;
          const bool sd_ok = sd_next ? mbar_try_wait(s_done, (it + 1) & 1) : true;   // polled here, needed before the P store
          if (dbg) p.dbg[h * 8 + 5] = clock64();
          pair_sync(lg);                                                 // partner's (max, sum) is visible
          const float2 oth = *reinterpret_cast<const float2*>(red + (r * 128 + t) * 4 + (ch ^ 1) * 2);
          const float mrow = fmaxf(m, oth.x);
          const float f_own = ex2(m - mrow), f_oth = ex2(oth.x - mrow);
          const float inv_sum = f_own / fmaf(s_own, f_own, oth.y * f_oth);
          if (DROP && p.drop.thresh) {                                   // nn.Dropout on the probabilities (maxvit.py:146, 209)
            const float ks = inv_sum * p.drop.scale;
            const uint32_t rid = drop_row(wdx, i);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t hsh = drop_hash(p.drop.seed, rid, drop_group_prob(p.drop.salt, h, ch * 8 + c));
#pragma unroll
              for (int k = 0; k < 4; ++k) sc[4 * c + k] *= (int)((hsh >> (8 * k)) & 255u) >= p.drop.thresh ? ks : 0.f;
            }
          } else {
            const float2 is2 = make_float2(inv_sum, inv_sum);
#pragma unroll
            for (int k = 0; k < 16; ++k) sc2[k] = fmul2(sc2[k], is2);
          }
          // P row (bf16 pairs, one 32-bit TMEM column per two keys): own 32 keys -> columns [half*32 + ch*16, +16); the same
          // keys of the other window are zero.
          {
            if (!sd_ok) mbar_wait_tag(s_done, (it + 1) & 1, 399);
            if (!tab_ok) mbar_wait_tag(tab_full + (r ^ 1), ((it + 1) >> 1) & 1, 394);
            const uint32_t tp = lane_addr + (r ? T_QKV0 : T_QKV1) + T_PQ;
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = pack_bf16(sc[2 * c], sc[2 * c + 1]);
            tmem_st16(tp + half * 32 + ch * 16, pk);
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = 0u;
            tmem_st16(tp + (half ^ 1) * 32 + ch * 16, pk);
            if (dbg) p.dbg[h * 8 + 6] = clock64();
            tmem_wait_st();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (dbg) p.dbg[h * 8 + 7] = clock64();
      }

      // (the epilogue of this tile is run by the staging warps)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_w_map(CUtensorMap* m, const void* ptr, long long inner, long long outer, int box_outer, bool f16) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q)
      return set_error("cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(q);
  }
  const int esz = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)inner * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_outer};      // 128-byte rows (one swizzle atom)
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error("attn_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// wqkv_h: fp16 [heads*96][128] (per head: 32 q rows, 32 k rows, 32 v rows); wout_h: fp32 [heads*128][32]
int attn_fused_run(const float* x, float* x_out, const float* reg_in, int reg_per_field, float* reg_out,
                   const float* film, const void* wqkv_h, const float* wout_h, const float* head_tab,
                   const AttnGeom& g, int heads, int dh, float ln_eps, unsigned seed, unsigned salt, int drop_thresh,
                   cudaStream_t st) {
  if (drop_thresh < 0 || drop_thresh > 255) return set_error("attn_fused: dropout threshold %d outside [0, 255]", drop_thresh);
  if (g.C != fa::C || dh != fa::DH) return set_error("attn_fused: needs C=128, dim_head=32 (got C=%d, dh=%d)", g.C, dh);
  if (heads < 4) return set_error("attn_fused: the head pipeline needs at least 4 heads (got %d)", heads);
  if (g.win != fa::WIN || g.R != fa::REG) return set_error("attn_fused: specialised for 7x7 windows + 4 register tokens (got %d, %d)", g.win, g.R);
  CUtensorMap mq, mo;
  int rc = make_w_map(&mq, wqkv_h, 128, (long long)heads * 96, 96, true);
  if (rc) return rc;
  rc = make_w_map(&mo, wout_h, 32, (long long)heads * 128, 128, false);
  if (rc) return rc;
  FusedAttnParams p;
  p.x = x; p.x_out = x_out; p.reg_in = reg_in; p.reg_per_field = reg_per_field; p.reg_out = reg_out; p.film = film;
  p.head_tab = head_tab; p.g = g; p.heads = heads;
  p.ln_eps = ln_eps; p.n_windows = (long long)g.N * g.nwin();
  p.drop.seed = seed; p.drop.salt = salt; p.drop.thresh = drop_thresh; p.drop.scale = 256.0f / (256.0f - (float)drop_thresh);
  p.dbg = nullptr;
  if (const char* e = getenv("VG_ATTN_DBG")) p.dbg = reinterpret_cast<long long*>(strtoull(e, nullptr, 0));
  static PerDeviceFlag attr_pd;
  bool& attr = attr_pd.cur();
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attn_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fa::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("attn_fused smem attr: %s", cudaGetErrorString(e));
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long n_tiles = (p.n_windows + 1) / 2;
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (drop_thresh) attn_fused_kernel<true><<<grid, fa::THREADS, fa::SMEM_BYTES, st>>>(mq, mo, p);
  else attn_fused_kernel<false><<<grid, fa::THREADS, fa::SMEM_BYTES, st>>>(mq, mo, p);
  return check_launch("attn_fused_kernel");
}

}  // namespace vg
