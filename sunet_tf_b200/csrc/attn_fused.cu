// Fused window attention (front half of a Swin block) for the bandwidth-bound stages (C = 96, 192):
//
//   norm1 (:233) -> roll (:238) -> window_partition (:243) -> qkv Linear (:114) -> q*scale (:117) -> q k^T (:118)
//   -> + relative_position_bias (:120-123) -> + shifted-window mask (:125-129) -> softmax -> @ v (:135)
//   -> window_reverse (:251) -> roll back (:255)                                  [SUNet_detail.py]
//
// in ONE kernel: the token stream is read once and the per-head attention output is written once; LayerNorm output,
// q/k/v and the 64x64 score matrices never exist in HBM.
//
// One persistent CTA per SM walks tiles of TWO windows (128 tokens):
//   gather : the 128 token rows are fetched through the roll+partition index map with 16-byte cp.async straight into
//            the 128-byte-swizzled K-major layout tcgen05.mma reads (window_partition / torch.roll are address arithmetic)
//   QKV    : D[128 x 3*GH*hd] = X[128 x C] * Wg^T on tcgen05 (fp16 in, fp32 accumulators in TMEM); the weights of the
//            current head group arrive by TMA.  norm1 runs in place on the gathered tile (fp32 statistics, 4 threads per
//            token); its affine part is folded into the weights at pre-pack: Wg = W * gamma, bias = b + W beta
//   drain  : TMEM -> registers -> fp16 q/k/v operand tiles in shared memory (head_dim padded to 16/32, XOR-swizzled rows)
//   core   : per (window, head): S = q k^T + bias (+ closed-form mask), exp2 softmax, O = P V on mma.sync.m16n8k16 with
//            register-resident S/P (K = 12/24: a 64x64xhd problem per head is below any tcgen05 tile), O parked in the q rows
//   scatter: O rows go back through the same index map (window_reverse + un-roll).
// C = 96 handles all 8 heads per pass (288 accumulator columns); C = 192 walks 4 groups of 2 heads (144 columns); C = 384 walks
// 8 single heads, with the group weights streaming through a k-block ring because they no longer fit beside the 96 KB token tile.
#include "attn_fused.cuh"

#include <stdio.h>
#include <stdlib.h>

#include "device.h"
#include "error.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

namespace {

// Phase cycle counters (SUNET_AF_TIMING=1 at run time): compiled in only with -DSUNET_KERNEL_TIMING=1, they cost ~25 registers
#ifndef SUNET_KERNEL_TIMING
#define SUNET_KERNEL_TIMING 0
#endif
#if SUNET_KERNEL_TIMING
#define AF_T(i) do { if (p.timing) { const long long _t = clock64(); tacc[i] += _t - tq0; tq0 = _t; } } while (0)
#define AF_T_DECL long long tacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define AF_T_START long long tq0 = p.timing ? clock64() : 0
#else
#define AF_T(i) do { } while (0)
#define AF_T_DECL do { } while (0)
#define AF_T_START do { } while (0)
#endif

constexpr float LOG2E = 1.4426950408889634f;
constexpr int NTHREADS = 512;
constexpr int TBL = 232;   // 225 bias entries per head, padded

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// D = A B + C with C in its own (read-only) registers: the first k-step of a score tile takes the relative-position bias slice
// straight from the registers that hold it for the whole kernel - no per-score copy into the accumulator (one MOV / FADD per score,
// ~10% of the instructions of attn_fused<96>, ncu source page r07)
__device__ __forceinline__ void mma_16816_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, float c0, float c1, float c2,
                                            float c3) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c0), "f"(c1), "f"(c2), "f"(c3));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {   // (ex2.approx.f16x2 becomes two MUFU.EX2.F16: same MUFU issue rate, measured)
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

template <int C_, int GH_>
struct FCfg {
  static constexpr int C = C_, HEADS = 8, HD = C / 8, GH = GH_, NG = HEADS / GH;
  static constexpr int HD_PAD = HD <= 16 ? 16 : (HD <= 32 ? 32 : 64);   // operand rows of 32 / 64 / 128 bytes
  static constexpr int BR = GH * HD;             // rows of one of q / k / v in a head group
  static constexpr int NGC = 3 * BR;             // accumulator columns per head group
  static constexpr int NMMA = NGC > 256 ? 3 : 1; // tcgen05.mma instructions per k-step (N <= 256 each)
  static constexpr int NPM = NGC / NMMA;
  static constexpr int KBD = (C + 63) / 64;      // 64-wide k-blocks of token data (one 128-byte swizzle row each)
  static constexpr int KTAIL = (C % 64) ? (C % 64) / 16 : 4;
  // The qkv bias rides the MMA: token columns C, C + 1 hold 1.0 against two extra weight columns with the folded bias as an fp16
  // (hi, lo) pair, in the spare 16-column k-step of the last k-block (BIASK 1: C = 96) or in one more k-block of which a single
  // k-step is multiplied (BIASK 2: C = 192, where shared memory has room for it).  The drain then is TMEM -> fp16 -> smem with no
  // bias loads or adds (90 of its ~160 instructions per thread and 72 columns).
#ifndef SUNET_AF_BIASK
#define SUNET_AF_BIASK 1
#endif
  static constexpr int BIASK = !SUNET_AF_BIASK ? 0
                               : ((C % 64) != 0 && (C % 64) <= 48) ? 1
                               : ((C % 64) == 0 && (KBD + 1) * NGC * 128 <= 80 * 1024) ? 2 : 0;
  static constexpr int KB = KBD + (BIASK == 2 ? 1 : 0);   // k-blocks of the token tile / weight buffer / MMA loop
  static constexpr int WPITCH = BIASK ? KB * 64 : C;   // row pitch (elements) of the packed weights
  static constexpr int RB = HD_PAD * 2;          // bytes per q/k/v operand row
  static constexpr int UNIT_BYTES = 64 * RB;     // one (q|k|v, window, head) operand tile
  static constexpr int NU = 2 * GH;              // (window, head) units per pass
  static constexpr int WPU = NU >= 4 ? 16 / NU : 4;   // warps per unit (a unit has 4 query tiles: with fewer than 4 units half the warps sit the core out)
  static constexpr int MT = 4 / WPU;             // 16-row query tiles per warp
  static constexpr int QC = NGC / 4;             // accumulator columns drained by one column-quarter
  static constexpr int CPR = C / 8;              // 16-byte chunks per token row
  static constexpr int VEC = (HD % 8 == 0) ? 8 : 4;
  static constexpr int VPH = HD / VEC;
  static constexpr int VPT = BR / VEC;           // output vectors per token per pass
  static constexpr int X_BYTES = KB * 16384;
  // weights of a head group: one buffer loaded at once, or (RING: when that does not fit) a ring of 64-wide k-blocks
  static constexpr bool RING = KB * NGC * 128 > 80 * 1024;
  static constexpr int RSTAGES = 4;
  static constexpr int WKB_BYTES = NGC * 128;    // one k-block of a group's weights
  static constexpr int W_BYTES = RING ? RSTAGES * WKB_BYTES : KB * NGC * 128;
  static constexpr int QKV_BYTES = 3 * NU * UNIT_BYTES;
  static constexpr int OFF_X = 0;
  static constexpr int OFF_W = OFF_X + X_BYTES;
  static constexpr int OFF_QKV = OFF_W + W_BYTES;
  static constexpr int OFF_TBL = OFF_QKV + QKV_BYTES;
  static constexpr bool TBL_SMEM = !RING;        // RING configs need the space for a 4th ring stage: the bias slice comes from global
  static constexpr int OFF_HC = OFF_TBL + (TBL_SMEM ? HEADS * TBL * 4 : 0);
  static constexpr int SMEM = OFF_HC + 3 * C * 4 + 1024;
  static constexpr uint32_t TMEM_COLS = NGC <= 256 ? 256 : 512;
  static_assert(C % 32 == 0 && HD % 4 == 0 && QC % 4 == 0, "column slices must be whole 4-column groups");
  static_assert((NPM * 128) % 1024 == 0 && NPM % 16 == 0 && NPM <= 256, "weight sub-tiles must be whole swizzle atoms");
  static_assert(WPU >= 1 && WPU <= 4 && WPU * NU <= 16 && MT * WPU == 4, "warp / unit split");
  static_assert(!RING || NMMA == 1, "the k-block ring carries one MMA-wide sub-tile per stage");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static_assert(!BIASK || !RING, "the bias k-step is only wired into the whole-buffer weight path");
  static_assert(!RING || KB == KBD, "RING mode streams data k-blocks only");
};

// byte offset of 16-byte chunk `ch` of row `row` inside a [64][HD_PAD] operand tile: consecutive rows are RB bytes apart and
// the chunk index is XOR-swizzled with the 128-byte line number so that the 8 rows of an ldmatrix phase hit 8 distinct
// 16-byte bank groups
template <int RB>
__device__ __forceinline__ uint32_t op_off(int row, int ch) {
  if constexpr (RB == 32) return static_cast<uint32_t>(row * 32 + ((ch ^ ((row >> 2) & 1)) << 4));
  else if constexpr (RB == 64) return static_cast<uint32_t>(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
  else return static_cast<uint32_t>(row * 128 + ((ch ^ (row & 7)) << 4));
}

// One 16-row query tile of one (window, head) unit; see attn_core.cu for the register-level scheme.  tb[e][k] holds this
// lane's slice of the relative-position bias, k = 2*MI + row_half - key_row + 7.  With a padded head column HD of V is
// 1.0, so O[:, HD] is the softmax denominator summed by the MMA over the fp16-rounded probabilities that multiply V.
template <int HD, int MT, int MI, bool MASK>
__device__ __forceinline__ void attn_tile(uint32_t q_h, uint32_t k_h, uint32_t v_h, int mt, int lane, const float (&tb)[2][2 * MT + 7],
                                          bool mrow, bool mcol) {
  constexpr int HD_PAD = HD <= 16 ? 16 : (HD <= 32 ? 32 : 64);
  constexpr int RB = HD_PAD * 2;
  constexpr int KS = HD_PAD / 16;
  constexpr int NO = HD_PAD / 8;
  constexpr bool MMA_SUM = HD_PAD != HD;
  const int g = lane >> 2, tq = lane & 3;
  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ldsm_x4(qa[ks], q_h + op_off<RB>(mt * 16 + (lane & 15), ks * 2 + (lane >> 4)));
  float s[8][4];
  if constexpr (MASK) {
    // closed-form SW-MSA mask (SUNet_detail.py:202-221, shift = 4): -100 where the wrapped halves differ, applied once
    constexpr float NEG = -100.f * LOG2E;
    const bool r0hi = (2 * mt) >= 4, r1hi = (2 * mt + 1) >= 4;
    float cm[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) cm[e] = (mcol && ((g >= 4) != ((2 * tq + e) >= 4))) ? NEG : 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float rm0 = (mrow && (r0hi != (nt >= 4))) ? NEG : 0.f;
      const float rm1 = (mrow && (r1hi != (nt >= 4))) ? NEG : 0.f;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = tb[e][2 * MI + 0 - nt + 7] + fminf(rm0, cm[e]);
        s[nt][2 + e] = tb[e][2 * MI + 1 - nt + 7] + fminf(rm1, cm[e]);
      }
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; nt += 2) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t kb[4];   // (nt, k lo), (nt, k hi), (nt+1, k lo), (nt+1, k hi)
      ldsm_x4(kb, k_h + op_off<RB>((nt + (lane >> 4)) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)));
      if (!MASK && ks == 0) {   // accumulators start from the bias: C operand = the bias registers themselves
        mma_16816_c(s[nt], qa[0], kb[0], kb[1], tb[0][2 * MI + 0 - nt + 7], tb[1][2 * MI + 0 - nt + 7], tb[0][2 * MI + 1 - nt + 7],
                    tb[1][2 * MI + 1 - nt + 7]);
        mma_16816_c(s[nt + 1], qa[0], kb[2], kb[3], tb[0][2 * MI + 0 - (nt + 1) + 7], tb[1][2 * MI + 0 - (nt + 1) + 7],
                    tb[0][2 * MI + 1 - (nt + 1) + 7], tb[1][2 * MI + 1 - (nt + 1) + 7]);
      } else {
        mma_16816(s[nt], qa[ks], kb[0], kb[1]);
        mma_16816(s[nt + 1], qa[ks], kb[2], kb[3]);
      }
    }
  }
  // softmax over the 64 keys (a row lives in the 4 lanes of a quad); logits are already in the exp2 domain
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
    m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  // Exponentials and P V interleaved per 16-key slab: the MUFU work of slab kk+1 is issued between the tensor-pipe work of
  // slab kk instead of in one burst (the asm statements are volatile, so the order below is the issue order; with every
  // warp of a scheduler in the same phase, a burst of 32 ex2 per lane leaves the FMA / tensor pipes idle and vice versa)
  float sum0 = 0.f, sum1 = 0.f;
  float o[NO][4];
#pragma unroll
  for (int n = 0; n < NO; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int nt = 2 * kk + h;
      const float e0 = ex2(s[nt][0] - m0), e1 = ex2(s[nt][1] - m0), e2 = ex2(s[nt][2] - m1), e3 = ex2(s[nt][3] - m1);
      pa[2 * h] = pack_half2(e0, e1);       // row i0, keys 2tq, 2tq+1 of key row nt
      pa[2 * h + 1] = pack_half2(e2, e3);   // row i1
      if constexpr (!MMA_SUM) {
        const float2 p0 = __half22float2(*reinterpret_cast<const __half2*>(&pa[2 * h]));
        const float2 p1 = __half22float2(*reinterpret_cast<const __half2*>(&pa[2 * h + 1]));
        sum0 += p0.x + p0.y;
        sum1 += p1.x + p1.y;
      }
    }
#pragma unroll
    for (int n = 0; n < NO; n += 2) {
      uint32_t vb[4];   // transposed 8x8 loads of V[key][d]: (keys lo, n), (keys hi, n), (keys lo, n+1), (keys hi, n+1)
      ldsm_x4_t(vb, v_h + op_off<RB>(kk * 16 + (lane & 15), n + (lane >> 4)));
      mma_16816(o[n], pa, vb[0], vb[1]);
      mma_16816(o[n + 1], pa, vb[2], vb[3]);
    }
  }
  if constexpr (MMA_SUM) {
    constexpr int NS = HD / 8, CS = HD % 8;
    const float c0 = (CS & 1) ? o[NS][1] : o[NS][0];
    const float c1 = (CS & 1) ? o[NS][3] : o[NS][2];
    const int src = (lane & ~3) | (CS >> 1);
    sum0 = __shfl_sync(0xffffffffu, c0, src);
    sum1 = __shfl_sync(0xffffffffu, c1, src);
  } else {
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  }
  // (MUFU.RCP: the IEEE form is eight instructions plus a slow-path call per reciprocal; 1 ulp is far below the fp16 rounding of O)
  float inv0, inv1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv0) : "f"(sum0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(sum1));
  // normalise and park O in this tile's own q rows (already consumed into registers by every lane of this warp)
  __syncwarp();
  const int i0 = mt * 16 + g, i1 = i0 + 8;
#pragma unroll
  for (int n = 0; n < NO; ++n) {
    if (n * 8 + 2 * tq < HD) {
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(q_h + op_off<RB>(i0, n) + 4 * tq), "r"(pack_half2(o[n][0] * inv0, o[n][1] * inv0)) : "memory");
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(q_h + op_off<RB>(i1, n) + 4 * tq), "r"(pack_half2(o[n][2] * inv1, o[n][3] * inv1)) : "memory");
    }
  }
}

template <int HD, int MT, bool MASK>
__device__ __forceinline__ void attn_tiles(uint32_t q_h, uint32_t k_h, uint32_t v_h, int mbase, int lane, const float (&tb)[2][2 * MT + 7],
                                           bool mrow, bool mcol) {
  attn_tile<HD, MT, 0, MASK>(q_h, k_h, v_h, mbase + 0, lane, tb, mrow, mcol);
  if constexpr (MT > 1) attn_tile<HD, MT, 1, MASK>(q_h, k_h, v_h, mbase + 1, lane, tb, mrow, mcol);
  if constexpr (MT > 2) {
    attn_tile<HD, MT, 2, MASK>(q_h, k_h, v_h, mbase + 2, lane, tb, mrow, mcol);
    attn_tile<HD, MT, 3, MASK>(q_h, k_h, v_h, mbase + 3, lane, tb, mrow, mcol);
  }
}

// Drain of one column quarter Q of the accumulator for one token row: qkv = D + bias -> fp16, written as 8-byte groups
// into the (q|k|v, window, head) operand tiles.  Everything that depends on the column is a compile-time constant; TMEM
// loads are issued in batches of 24 / 36 columns per wait.
template <typename K, int Q, int BATCH>
__device__ __forceinline__ void drain_load(uint32_t t_lane, int c0, uint32_t (&v)[BATCH]) {
#pragma unroll
  for (int i = 0; i + 8 <= BATCH; i += 8) tmem_ld8(t_lane + Q * K::QC + c0 + i, *reinterpret_cast<uint32_t(*)[8]>(&v[i]));
  if constexpr (BATCH % 8 != 0) tmem_ld4(t_lane + Q * K::QC + c0 + (BATCH / 8) * 8, *reinterpret_cast<uint32_t(*)[4]>(&v[(BATCH / 8) * 8]));
}
#ifndef SUNET_AF_DRAIN128
#define SUNET_AF_DRAIN128 1   // whole 16-byte chunks of an operand row are written with one conflict-free 128-bit store
#endif
template <typename K, int Q, int BATCH>
__device__ __forceinline__ void drain_store(const uint32_t (&v)[BATCH], int c0, const float* bf, uint32_t d_base, uint32_t d_sx) {
  // a thread owns one token row, rows are 32 / 64 / 128 bytes apart: 64-bit stores of a warp hit every bank pair twice (ncu: 2x the
  // ideal wavefronts), 128-bit stores are conflict-free with the chunk swizzle.  Columns [d, d + 8) of a head that start a chunk are
  // therefore stored as one 16-byte chunk; the 4-column remainder of a 12-wide head stays a 64-bit store.
#pragma unroll
  for (int i = 0; i < BATCH; i += 4) {
    const int n = Q * K::QC + c0 + i;             // compile-time after unrolling
    const int m = n / K::BR, j = n % K::BR;
    const int hl = j / K::HD, d = j % K::HD;
    const bool wide = SUNET_AF_DRAIN128 && (d % 8 == 0) && (d + 8 <= K::HD) && (i + 8 <= BATCH);
    const bool second = SUNET_AF_DRAIN128 && (d % 8 == 4) && (d + 4 <= K::HD) && (i >= 4);   // upper half of a chunk stored by the previous step
    if (second) continue;
    // (BIASK: the bias is already in the accumulator; the add is compiled out rather than fed zeros - `x + 0.f` is not a no-op the
    // compiler may drop (signed zeros) and cost one FADD per drained element)
    float q0 = __uint_as_float(v[i + 0]), q1 = __uint_as_float(v[i + 1]), q2 = __uint_as_float(v[i + 2]), q3 = __uint_as_float(v[i + 3]);
    if constexpr (!K::BIASK) {
      const float4 b4 = *reinterpret_cast<const float4*>(bf + n);
      q0 += b4.x; q1 += b4.y; q2 += b4.z; q3 += b4.w;
    }
    const uint32_t dst = d_base + (m * K::NU + hl) * K::UNIT_BYTES + (d & 7) * 2 + ((static_cast<uint32_t>(d >> 3) << 4) ^ d_sx);
    const uint32_t lo0 = pack_half2(q0, q1);
    const uint32_t lo1 = pack_half2(q2, q3);
    if (wide) {
      const int i4 = (i + 4 < BATCH) ? i + 4 : i;   // (always i + 4 when wide; keeps the index in range for the discarded branch)
      float r0 = __uint_as_float(v[i4 + 0]), r1 = __uint_as_float(v[i4 + 1]), r2 = __uint_as_float(v[i4 + 2]), r3 = __uint_as_float(v[i4 + 3]);
      if constexpr (!K::BIASK) {
        const float4 c4 = *reinterpret_cast<const float4*>(bf + n + 4);
        r0 += c4.x; r1 += c4.y; r2 += c4.z; r3 += c4.w;
      }
      sts128(dst, make_uint4(lo0, lo1, pack_half2(r0, r1), pack_half2(r2, r3)));
    } else {
      sts64(dst, lo0, lo1);
    }
  }
}
template <typename K, int Q>
__device__ __forceinline__ void drain_quarter(uint32_t t_lane, const float* bf, uint32_t d_base, uint32_t d_sx) {
  constexpr int QC = K::QC;
  if constexpr (QC % 24 == 0 && QC > 24) {
    // three batches of 24 columns, software-pipelined: the TMEM loads of batch b + 1 are in flight while batch b is converted and
    // stored (tcgen05.wait::ld covers everything issued so far; the drain is latency-, not bandwidth-bound)
    constexpr int NB = QC / 24;
    uint32_t va[24], vb[24];
    drain_load<K, Q, 24>(t_lane, 0, va);
    tmem_ld_wait();
#pragma unroll
    for (int bi = 0; bi < NB; ++bi) {
      if (bi & 1) {
        if (bi + 1 < NB) drain_load<K, Q, 24>(t_lane, (bi + 1) * 24, va);
        drain_store<K, Q, 24>(vb, bi * 24, bf, d_base, d_sx);
      } else {
        if (bi + 1 < NB) drain_load<K, Q, 24>(t_lane, (bi + 1) * 24, vb);
        drain_store<K, Q, 24>(va, bi * 24, bf, d_base, d_sx);
      }
      if (bi + 1 < NB) tmem_ld_wait();
    }
  } else {
    static_assert(QC % 4 == 0 && QC <= 40, "drain batch");
    uint32_t v[QC];
    drain_load<K, Q, QC>(t_lane, 0, v);
    tmem_ld_wait();
    drain_store<K, Q, QC>(v, 0, bf, d_base, d_sx);
  }
}

struct FParams {
  const __half* x;
  __half* out;
  const float* hconst;  // folded qkv bias [3C], permuted row order
  const float* table;   // [225][heads]
  int B, H, W, shift;
  long long* timing;    // optional [grid][8] phase cycle counters (SUNET_AF_TIMING bring-up aid), else null
};

template <int C, int GH>
__global__ void __launch_bounds__(NTHREADS, 1) attn_fused_kernel(const __grid_constant__ CUtensorMap tmW, const FParams p) {
  using K = FCfg<C, GH>;
  constexpr int RB = K::RB, HD = K::HD, MT = K::MT;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t w_full, mma_done;
  __shared__ __align__(8) uint64_t rk_full[K::RSTAGES], rk_empty[K::RSTAGES];   // RING: k-block ring of the group weights
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sX = smem_u32(smem + K::OFF_X), sW = smem_u32(smem + K::OFF_W), sQKV = smem_u32(smem + K::OFF_QKV);
  float* sTbl = reinterpret_cast<float*>(smem + K::OFF_TBL);
  const float* sBf = reinterpret_cast<const float*>(smem + K::OFF_HC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q4 = warp & 3;             // TMEM lane quadrant this warp may read
  const int quarter = warp >> 2;       // column quarter of the drain / statistics pass
  const int row = q4 * 32 + lane;      // token row of the tile owned in the drain

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    tma_prefetch_desc(&tmW);
    mbar_init(&w_full, 1);
    mbar_init(&mma_done, 1);
    for (int i = 0; i < K::RSTAGES; ++i) { mbar_init(&rk_full[i], 1); mbar_init(&rk_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, K::TMEM_COLS);
    tmem_relinquish();
  }
  if constexpr (K::TBL_SMEM) {
    for (int i = tid; i < K::HEADS * 225; i += NTHREADS) {
      const int e = i / K::HEADS, h = i - e * K::HEADS;
      sTbl[h * TBL + e] = __ldg(p.table + i) * LOG2E;
    }
  }
  if constexpr (!K::BIASK)   // (BIASK: the folded bias is in the weights, the drain does not read this table)
    for (int i = tid; i < 3 * C; i += NTHREADS) reinterpret_cast<float*>(smem + K::OFF_HC)[i] = __ldg(p.hconst + i);
  if constexpr (K::HD_PAD != HD) {
    // pad columns of every operand tile: 0 for q / k, V column HD = 1.0 (softmax denominator through the MMA); the drain and
    // the O store only ever write columns < HD, so this survives the whole kernel
    constexpr int PADW = (K::HD_PAD - HD) / 2;   // 32-bit words per row
    for (int i = tid; i < 3 * K::NU * 64 * PADW; i += NTHREADS) {
      const int w = i % PADW, r = (i / PADW) & 63, unit = i / (PADW * 64);   // unit over [q|k|v][window][head]
      const int d = HD + 2 * w;
      const uint32_t addr = sQKV + unit * K::UNIT_BYTES + op_off<RB>(r, d >> 3) + (d & 7) * 2;
      const uint32_t val = (unit / K::NU == 2 && w == 0) ? 0x00003C00u : 0u;
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(val) : "memory");
    }
  }
  if constexpr (K::BIASK) {
    // bias k-step of the token tile, written once: columns C, C + 1 = 1.0 (fp16 0x3C00), C + 2 .. C + 15 = 0.  The gather and the
    // LayerNorm only ever touch the C / 8 data chunks of a row.
    if (tid < 128) {
      constexpr int gb = C / 8;
      const uint32_t rowb = sX + (gb >> 3) * 16384 + tid * 128;
      sts128(rowb + ((static_cast<uint32_t>(gb & 7) ^ static_cast<uint32_t>(tid & 7)) << 4), make_uint4(0x3C003C00u, 0u, 0u, 0u));
      sts128(rowb + ((static_cast<uint32_t>((gb + 1) & 7) ^ static_cast<uint32_t>(tid & 7)) << 4), make_uint4(0u, 0u, 0u, 0u));
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();   // (the dependency wait itself sits below, after the first weight loads have been issued)

  const int nWc = p.W >> 3, nWr = p.H >> 3, nW = nWr * nWc;
  const long long nwin = static_cast<long long>(p.B) * nW;
  const long long tiles = (nwin + 1) >> 1;
  // Work items are (tile, head group) pairs in tile-major order, split into contiguous, equally long runs over the CTAs: with 3.5
  // tiles per SM (C = 192 at B = 64) whole-tile scheduling loses 13% to the last partial wave.  A CTA that starts or ends inside a
  // tile gathers and normalises that tile itself (the other CTA sharing it does the same; they write disjoint output columns).
  const long long total_items = tiles * K::NG;
  const long long it_begin = total_items * blockIdx.x / gridDim.x;
  const long long it_end = total_items * (blockIdx.x + 1) / gridDim.x;
  const int g_first = static_cast<int>(it_begin % K::NG);

  // Thread <-> token mapping of the gather / statistics / scatter passes: token tk = tid >> 2 of the tile (window
  // wi = warp >> 3, the same window whose (window, head) units this warp runs in the core), 16-byte part tid & 3.
#ifndef SUNET_AF_REMAP
#define SUNET_AF_REMAP 1
#endif
  // The two tokens of a quarter-warp (8 lanes = one 128-byte shared-memory wavefront of 16-byte accesses) must sit 4 rows apart:
  // rows r and r + 1 send their four 16-byte parts to the same four swizzled chunk slots (2-way conflict on every gather /
  // LayerNorm access of the token tile, ncu: 2x the ideal wavefronts), rows r and r + 4 to complementary halves of the 128-byte line.
  const int tk_lin = tid >> 2;
  const int tk = SUNET_AF_REMAP ? ((tk_lin & ~7) | ((tk_lin >> 1) & 3) | ((tk_lin & 1) << 2)) : tk_lin;
  const int part = tid & 3;
  const int wi = warp >> 3, tt = tk & 63;
  // geometry of this thread's window of a tile: global row of token tk (or -1 past the last window) and the mask flags
  struct Geo { int row; bool mrow, mcol; };
  auto tile_geo = [&](long long tile) {
    Geo gg;
    const unsigned win = static_cast<unsigned>(tile) * 2u + wi;
    const unsigned b = win / static_cast<unsigned>(nW);
    const unsigned wimg = win - b * nW;
    const unsigned wr = wimg / static_cast<unsigned>(nWc), wc = wimg - wr * nWc;
    int r = static_cast<int>(wr) * 8 + p.shift + (tt >> 3), c = static_cast<int>(wc) * 8 + p.shift + (tt & 7);
    if (r >= p.H) r -= p.H;
    if (c >= p.W) c -= p.W;
    gg.row = win < static_cast<unsigned>(nwin) ? (static_cast<int>(b) * p.H + r) * p.W + c : -1;
    gg.mrow = p.shift > 0 && static_cast<int>(wr) == nWr - 1;
    gg.mcol = p.shift > 0 && static_cast<int>(wc) == nWc - 1;
    return gg;
  };
  // 16-byte chunk cg = part + 4 j of token tk inside the SW128 token tile
  auto x_chunk = [&](int j) -> uint32_t {
    const int cg = part + 4 * j;
    return sX + (cg >> 3) * 16384 + tk * 128 + ((static_cast<uint32_t>(cg & 7) ^ static_cast<uint32_t>(tk & 7)) << 4);
  };
  auto gather = [&](const Geo& gg) {
    const __half* src = p.x + static_cast<long long>(gg.row) * C + part * 8;
#pragma unroll
    for (int j = 0; j < K::CPR / 4; ++j) {
      if (gg.row >= 0) cp_async16(x_chunk(j), src + j * 32);
      else asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(x_chunk(j)), "r"(0u) : "memory");
    }
    cp_async_commit();
  };
  // LayerNorm of token tk in place (norm1 without its affine part, which is folded into the weights): every thread
  // normalises the chunks it gathered itself; the 4 threads of a token combine their partial sums by shuffle.
  // fp32 statistics with a shifted one-pass variance.
  auto normalize = [&]() {
    constexpr int NJ = K::CPR / 4;
    constexpr bool KEEP = NJ <= 6;     // narrow rows stay in registers between the two passes; wide ones are re-read from smem
    uint4 v[KEEP ? NJ : 1];
    float k0 = 0.f;
    {
      const uint4 f0 = lds128(x_chunk(0));
      k0 = __half2float(__ushort_as_half(static_cast<unsigned short>(f0.x & 0xffffu)));
      k0 = __shfl_sync(0xffffffffu, k0, lane & ~3);   // first element of the row (held by part 0)
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const uint4 u = lds128(x_chunk(j));
      if constexpr (KEEP) v[j] = u;
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(h2[t]);
        const float d0 = f.x - k0, d1 = f.y - k0;
        s1 += d0 + d1;
        s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
      }
    }
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
    const float ms = s1 * (1.0f / C);
    const float var = fmaxf(s2 * (1.0f / C) - ms * ms, 0.f);
    const float a = rsqrtf(var + 1e-5f);
    const float b = -(k0 + ms) * a;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      uint4 u;
      if constexpr (KEEP) u = v[j]; else u = lds128(x_chunk(j));
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float2 f = __half22float2(h2[t]);
        o2[t] = __floats2half2_rn(fmaf(f.x, a, b), fmaf(f.y, a, b));
      }
      sts128(x_chunk(j), o);
    }
    fence_proxy_async_smem();   // the MMA issuer reads the tile through the async proxy after the next block barrier
  };
  auto load_w = [&](int g) {   // thread 0: the folded qkv weights of head group g -> smem, [kb][sub-tile][NPM rows][64] SW128
    mbar_arrive_expect_tx(&w_full, K::W_BYTES);
#pragma unroll
    for (int kb = 0; kb < K::KB; ++kb)
#pragma unroll
      for (int m = 0; m < K::NMMA; ++m)
        tma_load_2d(smem + K::OFF_W + (kb * K::NMMA + m) * K::NPM * 128, &tmW, &w_full, kb * 64, g * K::NGC + m * K::NPM);
  };
  // The issuing WARP runs these convergently and one elected lane issues: inside a single-thread branch every descriptor is
  // thread-divergent for the compiler (R2UR + an ELECT loop per tcgen05.mma, ~180 clk per instruction - see gemm_tcgen05.cu).
  auto issue_mma = [&](uint32_t it) {   // warp 0: D[128 x NGC] = X * Wg^T for work item `it`
    if (K::NG > 1 || it == 0) mbar_wait(&w_full, it & 1);   // a single head group keeps its weights for the whole kernel
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, K::NPM);
    if (elect_one()) {
#pragma unroll
      for (int kb = 0; kb < K::KB; ++kb) {
        const int ksteps = kb < K::KBD - 1 ? 4 : (kb == K::KBD - 1 ? K::KTAIL + (K::BIASK == 1 ? 1 : 0) : 1);   // (kb == KBD: the bias k-block, one k-step)
        const uint64_t adesc = umma_desc_sw128(sX + kb * 16384);
#pragma unroll
        for (int k = 0; k < ksteps; ++k)
#pragma unroll
          for (int m = 0; m < K::NMMA; ++m) {
            const uint64_t bdesc = umma_desc_sw128(sW + (kb * K::NMMA + m) * K::NPM * 128);
            umma_f16_ss(tmem_base + m * K::NPM, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
          }
      }
      tc_commit(&mma_done);
    }
    __syncwarp();
  };

  // RING mode (wide rows): the weights of a head group do not fit next to the token tile, so they stream through a ring of
  // 64-wide k-blocks.  One thread of a warp that sits the core out is both TMA producer and MMA issuer; the ring runs ahead
  // across items (the weight sequence does not depend on the tile), refilling a slot one k-block after its MMAs were issued.
  uint32_t rk_loaded = 0;      // k-blocks whose TMA has been issued (sequence number over items)
  const uint32_t rk_total = K::RING ? static_cast<uint32_t>((it_end - it_begin) * K::KB) : 0u;
  auto ring_load = [&](uint32_t empty_ok) {   // issue the TMA of k-block rk_loaded (its slot must be free; empty_ok: an earlier test said so)
    const uint32_t c = rk_loaded, slot = c % K::RSTAGES;
    if (c >= K::RSTAGES) mbar_wait_hint(&rk_empty[slot], ((c / K::RSTAGES) - 1) & 1, empty_ok);
    const int g = static_cast<int>((g_first + c / K::KB) % K::NG), kb = static_cast<int>(c % K::KB);
    mbar_arrive_expect_tx(&rk_full[slot], K::WKB_BYTES);
    tma_load_2d(smem + K::OFF_W + slot * K::WKB_BYTES, &tmW, &rk_full[slot], kb * 64, g * K::NGC);
    ++rk_loaded;
  };
  // RING mode runs the ring's TMA producer (lane 0 of a second idle warp) and the MMA issuer on different warps: the issuer only
  // waits for full slots, issues and commits - the k-block loop of a single producer + issuer thread (~1k clk per k-block of barrier
  // round trips) did not fit under the core and showed up as a 10% wait for the accumulator at the top of every item (round 1,
  // tools/ab_af_split.sh: 55.9 -> 49.6 us per launch at C = 384).
  auto ring_produce = [&](uint32_t upto) {   // producer thread: TMA of k-blocks [rk_loaded, min(upto, rk_total))
    while (rk_loaded < rk_total && rk_loaded < upto) ring_load(0u);
  };
  auto ring_mma_only = [&](uint32_t it) {
    const uint32_t idesc = umma_idesc_f16(128, K::NPM);
    const uint32_t c0 = it * K::KB;
    uint32_t full_ok = mbar_test(&rk_full[c0 % K::RSTAGES], (c0 / K::RSTAGES) & 1);
#pragma unroll 1
    for (int kb = 0; kb < K::KB; ++kb) {
      const uint32_t c = c0 + kb, slot = c % K::RSTAGES;
      mbar_wait_hint(&rk_full[slot], (c / K::RSTAGES) & 1, full_ok);
      const uint32_t cn = c + 1;
      full_ok = kb + 1 < K::KB ? mbar_test(&rk_full[cn % K::RSTAGES], (cn / K::RSTAGES) & 1) : 0u;
      tc_fence_after();
      const uint64_t adesc = umma_desc_sw128(sX + kb * 16384);
      const uint64_t bdesc = umma_desc_sw128(sW + slot * K::WKB_BYTES);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (kb < K::KB - 1 || k < K::KTAIL)
            umma_f16_ss(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        tc_commit(&rk_empty[slot]);
        if (kb == K::KB - 1) tc_commit(&mma_done);
      }
      __syncwarp();
    }
  };
  // the thread that issues TMA + MMA: lane 0 of warp 0, or in RING mode of a warp without a query tile in the core
  constexpr int ISSUER_WARP = K::RING ? K::WPU * K::GH : 0;
  const bool is_producer = K::RING && tid == (ISSUER_WARP + 1) * 32;
  static_assert(!K::RING || (K::NU * K::WPU < 16 && ISSUER_WARP < 8), "RING mode needs an idle warp for the issuer");
  const bool is_issuer = warp == ISSUER_WARP;   // (the whole warp; one elected lane issues)

  // this warp's (window, head) unit of the core pass
  const int u_hl = (warp & 7) % GH;
  const int unit = wi * GH + u_hl;
  const int mbase = ((warp & 7) / GH) * MT;
  const bool core_warp = mbase < 4;   // with fewer than 4 units per pass some warps have no query tile
  const int lg = lane >> 2, ltq = lane & 3;
  float tb[2][2 * MT + 7];
  int tb_head = -1;
  // drain constants: token row `row` of the tile, column quarter `quarter`
  const uint32_t d_base = sQKV + ((row >> 6) * GH) * K::UNIT_BYTES + (row & 63) * RB;
  const uint32_t d_sx = (RB == 32 ? ((row >> 2) & 1) : (RB == 64 ? ((row >> 1) & 3) : (row & 7))) << 4;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);

  uint32_t item = 0;
  AF_T_DECL;
  Geo geo = tile_geo(it_begin / K::NG);
  // Everything up to here read only parameters (constant across the forward); so do the first weight loads: they are issued
  // before the programmatic-dependency wait and arrive under the previous kernel's tail.  The token stream is touched below.
  if (it_begin < it_end) {
    if (!K::RING && tid == 0) load_w(g_first);
    if (is_producer) ring_produce(K::RSTAGES);   // the free slots only: nothing here may block (no MMA has been issued yet)
  }
  pdl_wait();
  if (it_begin < it_end) {
    gather(geo);
    cp_async_wait_all();
    normalize();
    tc_fence_before();
    __syncthreads();
    if (is_producer) ring_produce(K::KB + K::RSTAGES - 1);
    if (is_issuer) {
      tc_fence_after();
      if (K::RING) ring_mma_only(0); else issue_mma(0);
    }
    __syncwarp();
  }
  uint32_t md_ok = 0;  // early-test result for the current item's mma_done phase
  Geo geo_next = geo;
  {
#pragma unroll 1
    for (long long it = it_begin; it < it_end; ++it, ++item) {
      const long long tile = it / K::NG;
      const int g = static_cast<int>(it - tile * K::NG);
      const bool last_g = g == K::NG - 1;
      const bool has_next = it + 1 < it_end;
      const bool has_next_tile = has_next && last_g;   // the next item opens a new tile
      const long long next_tile = tile + 1;
      AF_T_START;
      mbar_wait_hint(&mma_done, item & 1, md_ok);
      tc_fence_after();
      AF_T(0);
      // the MMAs of this item have read the weight buffer (and, for the last group, the token tile): refill them
      if (!K::RING && K::NG > 1 && tid == 0 && has_next) load_w(last_g ? 0 : g + 1);   // (NG == 1: loaded once, never replaced)
      if (last_g && has_next_tile) {
        geo_next = tile_geo(next_tile);
        gather(geo_next);
      }
      // ---- drain: qkv[row][n] = rstd * D - rstd * mean * s_n + bf_n  -> fp16 operand tiles
      {
        const float* bf = sBf + g * K::NGC;
        switch (quarter) {
          case 0: drain_quarter<K, 0>(t_lane, bf, d_base, d_sx); break;
          case 1: drain_quarter<K, 1>(t_lane, bf, d_base, d_sx); break;
          case 2: drain_quarter<K, 2>(t_lane, bf, d_base, d_sx); break;
          default: drain_quarter<K, 3>(t_lane, bf, d_base, d_sx); break;
        }
      }
      AF_T(1);
      if (last_g && has_next_tile) {   // next tile's LayerNorm, in place in the token tile; barrier (B) below publishes it to the issuer
        cp_async_wait_all();
        normalize();
      }
      AF_T(2);
      tc_fence_before();
      __syncthreads();   // (B) q/k/v operand tiles complete, accumulator drained
      AF_T(3);
      AF_T(8);
      if (is_producer && has_next) ring_produce((item + 2) * K::KB + K::RSTAGES - 1);
      if (is_issuer && has_next) {                     // runs on the tensor pipe while the core below runs on the CUDA cores
        if (K::RING) ring_mma_only(item + 1); else issue_mma(item + 1);
      }
      __syncwarp();
      AF_T(9);
      // ---- core
      {
        const int head = g * GH + u_hl;
        if (head != tb_head && core_warp) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int k = 0; k < 2 * MT + 7; ++k) {
              const int idx = (k + 2 * mbase) * 15 + (lg - 2 * ltq - e + 7);
              tb[e][k] = K::TBL_SMEM ? sTbl[head * TBL + idx] : __ldg(p.table + idx * K::HEADS + head) * LOG2E;
            }
          tb_head = head;
        }
        if (geo.row >= 0 && core_warp) {   // uniform per warp: all its tokens belong to one window
          const uint32_t q_h = sQKV + (0 * K::NU + unit) * K::UNIT_BYTES;
          const uint32_t k_h = sQKV + (1 * K::NU + unit) * K::UNIT_BYTES;
          const uint32_t v_h = sQKV + (2 * K::NU + unit) * K::UNIT_BYTES;
          if (geo.mrow || geo.mcol) attn_tiles<HD, MT, true>(q_h, k_h, v_h, mbase, lane, tb, geo.mrow, geo.mcol);
          else attn_tiles<HD, MT, false>(q_h, k_h, v_h, mbase, lane, tb, false, false);
        }
      }
      AF_T(4);
      const uint32_t md_next = has_next ? mbar_test(&mma_done, (item + 1) & 1) : 0u;   // looked up under the scatter (a test costs ~170 clk)
      // (C) O rows of every unit parked in the q tiles.  The scatter of a window reads only what the 8 core warps of the same window
      // (warps 8 wi .. 8 wi + 7, the very warps that scatter it) have written: a per-window barrier, so a window with cheaper units
      // (no mask) does not wait for the other one.
#ifndef SUNET_AF_CWIN
#define SUNET_AF_CWIN 1
#endif
      if (SUNET_AF_CWIN) named_bar_sync(1 + wi, 256);
      else __syncthreads();
      AF_T(5);
      // ---- scatter (heads are concatenated in order, :135; window_reverse + un-roll through the row map): the 4 threads
      // of a token write consecutive vectors, so every store instruction covers whole 32-byte sectors
      if (geo.row >= 0) {
        __half* dst = p.out + static_cast<long long>(geo.row) * C + g * K::BR;
#pragma unroll
        for (int j = 0; j < (K::VPT + 3) / 4; ++j) {
          const int vv = part + 4 * j;
          if (K::VPT % 4 == 0 || vv < K::VPT) {
            const int hl = vv / K::VPH, d0 = (vv - hl * K::VPH) * K::VEC;
            const uint32_t src = sQKV + (wi * GH + hl) * K::UNIT_BYTES + op_off<RB>(tt, d0 >> 3) + (d0 & 7) * 2;
            if constexpr (K::VEC == 8) {
              *reinterpret_cast<uint4*>(dst + vv * 8) = lds128(src);
            } else {
              uint2 o;
              asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(o.x), "=r"(o.y) : "r"(src) : "memory");
              *reinterpret_cast<uint2*>(dst + vv * 4) = o;
            }
          }
        }
      }
      AF_T(6);
      md_ok = md_next;
      // (D) q tiles free for the next drain.  The drain of a warp overwrites the q / k / v rows of ONE window - the one its TMEM lane
      // quadrant holds (rows 32 q4 .. 32 q4 + 31, window q4 >> 1) - so it has to wait for the scatter of that window only: the 8
      // scatter warps of window w and the 4 warps of the other half that drain w's rows meet on barrier 3 + w (12 warps); a warp that
      // scatters w but drains the other window only arrives.
      // Measured SLOWER than the CTA-wide barrier (in-call A/B, whole forward: 8.374 vs 8.316 ms; C = 96 113.0 vs 111.0 us): off.
#ifndef SUNET_AF_DWIN
#define SUNET_AF_DWIN 0
#endif
      if (SUNET_AF_DWIN) {
        const int wd = q4 >> 1;   // window whose rows this warp drains
        if (wd == wi) {
          named_bar_sync(3 + wi, 384);
        } else {
          asm volatile("bar.arrive %0, %1;" ::"r"(3 + wi), "r"(384) : "memory");
          named_bar_sync(3 + wd, 384);
        }
      } else {
        __syncthreads();
      }
      AF_T(7);
      if (last_g) geo = geo_next;
    }
  }
#if SUNET_KERNEL_TIMING
  if (p.timing && (tid & 31) == 0) {
#pragma unroll
    for (int i = 0; i < 10; ++i) p.timing[(static_cast<long long>(blockIdx.x) * 16 + warp) * 10 + i] = tacc[i];
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, K::TMEM_COLS);
  }
}

// ---- pre-pack: permuted rows, W * gamma (q rows additionally * qscale), row sums of the rounded weights, folded bias
__global__ void attn_fold_kernel(const float* __restrict__ wqkv, const float* __restrict__ bqkv, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, __half* __restrict__ wp, float* __restrict__ hconst, int C, int HD,
                                 int GH, float qscale, int pitch) {
  const int pr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (pr >= 3 * C) return;
  const int BR = GH * HD, NGC = 3 * BR;
  const int g = pr / NGC, within = pr - g * NGC;
  const int m = within / BR, j = within - m * BR;
  const int o = m * C + g * BR + j;   // original row: [q|k|v][head][d]
  const float sc = m == 0 ? qscale : 1.f;
  float bb = 0.f;
  for (int k = lane; k < C; k += 32) {
    const float w = wqkv[static_cast<size_t>(o) * C + k];
    wp[static_cast<size_t>(pr) * pitch + k] = __float2half_rn(w * gamma[k] * sc);
    bb = fmaf(w, beta[k], bb);
  }
  for (int off = 16; off > 0; off >>= 1) bb += __shfl_xor_sync(0xffffffffu, bb, off);
  const float hb = (bb + (bqkv ? bqkv[o] : 0.f)) * sc;
  if (lane == 0) hconst[pr] = hb;
  if (pitch > C) {   // bias k-step: (hi, lo) fp16 pair at columns C, C + 1, zeros up to the pitch
    const __half hi = __float2half_rn(hb);
    const __half lo = __float2half_rn(hb - __half2float(hi));
    for (int k = C + lane; k < pitch; k += 32)
      wp[static_cast<size_t>(pr) * pitch + k] = k == C ? hi : (k == C + 1 ? lo : __float2half_rn(0.f));
  }
}

template <int C, int GH>
int launch_t(const AttnFusedPack& p, const __half* x, __half* out, int B, int H, int W, int shift, cudaStream_t stream) {
  using K = FCfg<C, GH>;
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(attn_fused_kernel<C, GH>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    once.done();
  }
  const int sms = device_sms();
  FParams prm;
  prm.x = x; prm.out = out;
  prm.hconst = p.hconst;
  prm.table = p.table;
  prm.B = B; prm.H = H; prm.W = W; prm.shift = shift;
  prm.timing = nullptr;
  static long long* timing_buf = nullptr;
  const bool timing = SUNET_KERNEL_TIMING && getenv("SUNET_AF_TIMING") != nullptr;
  if (timing) {
    if (!timing_buf) SUNET_CUDA(cudaMalloc(&timing_buf, 148 * 16 * 10 * sizeof(long long)));
    SUNET_CUDA(cudaMemsetAsync(timing_buf, 0, 148 * 16 * 10 * sizeof(long long), stream));
    prm.timing = timing_buf;
  }
  const long long nwin = static_cast<long long>(B) * (H / 8) * (W / 8);
  if (nwin > 0x3fffffffLL || static_cast<long long>(B) * H * W > 0x7fffffffLL) return fail(SUNET_E_SHAPE, "fused attention: too many tokens for 32-bit row indices");
  const long long tiles = (nwin + 1) / 2;
  const long long items = tiles * K::NG;
  const unsigned grid = static_cast<unsigned>(items < sms ? items : sms);
  SUNET_CUDA(launch_pdl(attn_fused_kernel<C, GH>, dim3(grid), dim3(NTHREADS), K::SMEM, stream, p.tmW, prm));
  if (timing) {   // bring-up aid: per-phase cycles averaged over CTAs, for warp 0 (MMA issuer) and the mean of the other warps
    SUNET_CUDA(cudaStreamSynchronize(stream));
    static long long host[148 * 16 * 10];
    SUNET_CUDA(cudaMemcpy(host, timing_buf, sizeof(host), cudaMemcpyDeviceToHost));
    static const char* names[10] = {"wait_mma", "drain", "cpwait", "barB", "core", "barC", "scatter", "barD", "normalize", "issue"};
    double w0[10] = {0}, wr[10] = {0};
    for (unsigned b = 0; b < grid; ++b)
      for (int w = 0; w < 16; ++w)
        for (int i = 0; i < 10; ++i) (w == 0 ? w0[i] : wr[i]) += static_cast<double>(host[(b * 16 + w) * 10 + i]);
    fprintf(stderr, "attn_fused<%d> tiles=%lld grid=%u cycles per CTA:", C, tiles, grid);
    for (int i = 0; i < 10; ++i) fprintf(stderr, " %s %.0f/%.0f", names[i], w0[i] / grid, wr[i] / grid / 15);
    fprintf(stderr, "\n");
  }
  return 0;
}

constexpr int group_heads(int C) { return C == 96 ? 8 : (C == 192 ? 2 : 1); }

}  // namespace

bool attn_fused_supported(int C, int heads) { return heads == 8 && (C == 96 || C == 192 || C == 384); }
int attn_fused_w_pitch(int C) {
  switch (C) {
    case 96: return FCfg<96, group_heads(96)>::WPITCH;
    case 192: return FCfg<192, group_heads(192)>::WPITCH;
    case 384: return FCfg<384, group_heads(384)>::WPITCH;
    default: return C;
  }
}

int attn_fused_prepack(AttnFusedPack* p, int C, int heads, float qscale, const float* gamma, const float* beta, const float* wqkv,
                       const float* bqkv, const float* table, cudaStream_t stream) {
  if (!attn_fused_supported(C, heads)) return fail(SUNET_E_SHAPE, "fused attention: C=%d heads=%d not instantiated", C, heads);
  if (!p->w || !p->hconst) return fail(SUNET_E_ARG, "fused attention: pack buffers not allocated");
  p->C = C; p->heads = heads; p->table = table;
  const int GH = group_heads(C), HD = C / heads;
  const int pitch = attn_fused_w_pitch(C);
  attn_fold_kernel<<<(3 * C + 7) / 8, 256, 0, stream>>>(wqkv, bqkv, gamma, beta, p->w, p->hconst, C, HD, GH, qscale, pitch);
  SUNET_CHECK_LAUNCH();
  const int NGC = 3 * GH * HD;
  const int NPM = NGC > 256 ? NGC / 3 : NGC;
  SUNET_TRY(make_tmap_2d_f16(&p->tmW, p->w, pitch, 3 * C, pitch, NPM));
  return 0;
}

int attn_fused_launch(const AttnFusedPack& p, const __half* x, __half* out, int B, int H, int W, int shift, cudaStream_t stream) {
  if (H % 8 || W % 8) return fail(SUNET_E_SHAPE, "fused attention: token grid %dx%d must be a multiple of the 8x8 window", H, W);
  if (shift != 0 && shift != 4) return fail(SUNET_E_ARG, "fused attention: shift %d unsupported (0 or 4)", shift);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(SUNET_E_ALIGN, "fused attention: x/out must be 16-byte aligned");
  if (B <= 0) return fail(SUNET_E_SHAPE, "fused attention: batch %d", B);
  switch (p.C) {
    case 96: return launch_t<96, group_heads(96)>(p, x, out, B, H, W, shift, stream);
    case 192: return launch_t<192, group_heads(192)>(p, x, out, B, H, W, shift, stream);
    case 384: return launch_t<384, group_heads(384)>(p, x, out, B, H, W, shift, stream);
    default: return fail(SUNET_E_SHAPE, "fused attention: C=%d not instantiated", p.C);
  }
}

}  // namespace sunet
