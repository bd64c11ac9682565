// Window-attention core: per (window, head)  S = Q K^T + rel-pos bias (+ shifted-window mask),
// softmax over the 64 keys, O = P V.   Restates SUNet_detail.py:118-135 for 8x8 windows.
//
// The cyclic shift (torch.roll, :238/:255) and window_partition / window_reverse (:27-56) are never
// materialised: q/k/v rows are gathered from the image-order token tensor with
//     src = ((wr*8 + r + shift) mod H, (wc*8 + c + shift) mod W)
// and O is scattered back through the same map.
//
// q arrives pre-multiplied by qk_scale * log2(e) (folded into the qkv weights at pre-pack), the bias table and the mask
// are multiplied by log2(e) when they are staged, so the softmax is a bare exp2.
//
// Persistent kernel: a CTA walks (window, head-group) units.  The q/k/v rows of the NEXT unit are gathered with
// cp.async (8- or 16-byte, head-wise re-pack into zero/one-padded rows) into the second smem stage while the warps
// run the register-resident flash-style core (mma.sync.m16n8k16, fp16 in, fp32 accumulate) on the current one.
// The linear layers around it (>= 90% of the block FLOPs) run on tcgen05 (gemm_tcgen05.cu, mlp_fused.cu).
#include "attn_core.cuh"
#include "device.h"
#include "error.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

namespace {

constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 16-row query tile of one head.  MI (tile index within this warp) is a template parameter so that the
// relative-position bias - held in registers as tb[e][k], k = 2*MI + row_half - key_row + 7 - is indexed at compile time.
// With a padded head (HD < HD_PAD) column HD of V holds 1.0, so O[:, HD] is the softmax denominator summed by the MMA
// over exactly the fp16-rounded probabilities that multiply V.
template <int HD, int MT, int MI, int MASK>
__device__ __forceinline__ void attn_tile(const __half* q_h, const __half* k_h, const __half* v_h, __half* o_h, int mt, int lane,
                                          const float (&tb)[2][2 * MT + 7], bool mrow, bool mcol, const float* mexp) {
  constexpr int HD_PAD = (HD + 15) / 16 * 16;
  constexpr int LDS = HD_PAD + 8;
  constexpr int KS = HD_PAD / 16;
  constexpr int NO = HD_PAD / 8;
  constexpr bool MMA_SUM = HD_PAD != HD;
  const int g = lane >> 2, tq = lane & 3;
  uint32_t qa[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
    ldsm_x4(qa[ks], smem_u32(q_h + (mt * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8));
  // S accumulators start from the bias (+ mask): rows i0 = mt*16+g (window row 2mt) and i1 = i0+8 (window row 2mt+1),
  // keys j = nt*8 + 2*tq + e (window row nt, column 2*tq+e)
  float s[8][4];
  if (MASK == 1) {
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
  } else {
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s[nt][e] = tb[e][2 * MI + 0 - nt + 7];
        s[nt][2 + e] = tb[e][2 * MI + 1 - nt + 7];
      }
    if (MASK == 2) {  // explicit additive mask tensor (stand-alone WindowAttention.forward(x, mask))
      const int i0 = mt * 16 + g;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = nt * 8 + 2 * tq + e;
          s[nt][e] = fmaf(__ldg(mexp + i0 * 64 + j), LOG2E, s[nt][e]);
          s[nt][2 + e] = fmaf(__ldg(mexp + (i0 + 8) * 64 + j), LOG2E, s[nt][2 + e]);
        }
    }
  }
#pragma unroll
  for (int nt = 0; nt < 8; nt += 2) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      // matrices: (nt, k lo), (nt, k hi), (nt+1, k lo), (nt+1, k hi)
      uint32_t kb[4];
      ldsm_x4(kb, smem_u32(k_h + ((nt + (lane >> 4)) * 8 + (lane & 7)) * LDS + ks * 16 + ((lane >> 3) & 1) * 8));
      mma_16816(s[nt], qa[ks], kb[0], kb[1]);
      mma_16816(s[nt + 1], qa[ks], kb[2], kb[3]);
    }
  }
  // ---- softmax over 64 keys (each row lives in the 4 lanes of a quad); logits are already in the exp2 domain
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
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = ex2(s[nt][0] - m0);
    s[nt][1] = ex2(s[nt][1] - m0);
    s[nt][2] = ex2(s[nt][2] - m1);
    s[nt][3] = ex2(s[nt][3] - m1);
    if constexpr (!MMA_SUM) {
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
  }
  // ---- O = P V
  float o[NO][4];
#pragma unroll
  for (int n = 0; n < NO; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_half2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_half2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_half2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_half2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int n = 0; n < NO; n += 2) {
      // transposed 8x8 loads of V[key][hd]: (keys lo, n), (keys hi, n), (keys lo, n+1), (keys hi, n+1)
      uint32_t vb[4];
      ldsm_x4_t(vb, smem_u32(v_h + (kk * 16 + (lane & 15)) * LDS + (n + (lane >> 4)) * 8));
      mma_16816(o[n], pa, vb[0], vb[1]);
      mma_16816(o[n + 1], pa, vb[2], vb[3]);
    }
  }
  float inv0, inv1;
  if constexpr (MMA_SUM) {
    // column HD sits in n-tile HD/8 at in-tile column HD%8: held by the quad lane tq == (HD%8)/2, element (HD%8)&1
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
  // (MUFU.RCP; the IEEE reciprocal is eight inline instructions plus a slow-path call - see attn_fused.cu)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv0) : "f"(sum0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1) : "f"(sum1));
  // ---- normalise and park O in this tile's own Q rows (already consumed into registers)
  __syncwarp();
  const int i0 = mt * 16 + g, i1 = i0 + 8;
#pragma unroll
  for (int n = 0; n < NO; ++n) {
    const int c = n * 8 + 2 * tq;
    if (c < HD) {
      *reinterpret_cast<uint32_t*>(o_h + i0 * LDS + c) = pack_half2(o[n][0] * inv0, o[n][1] * inv0);
      *reinterpret_cast<uint32_t*>(o_h + i1 * LDS + c) = pack_half2(o[n][2] * inv1, o[n][3] * inv1);
    }
  }
}

template <int HD, int MT, int MASK>
__device__ __forceinline__ void attn_tiles(const __half* q_h, const __half* k_h, const __half* v_h, __half* o_h, int mbase, int lane,
                                           const float (&tb)[2][2 * MT + 7], bool mrow, bool mcol, const float* mexp) {
  attn_tile<HD, MT, 0, MASK>(q_h, k_h, v_h, o_h, mbase + 0, lane, tb, mrow, mcol, mexp);
  if constexpr (MT > 1) attn_tile<HD, MT, 1, MASK>(q_h, k_h, v_h, o_h, mbase + 1, lane, tb, mrow, mcol, mexp);
  if constexpr (MT > 2) {
    attn_tile<HD, MT, 2, MASK>(q_h, k_h, v_h, o_h, mbase + 2, lane, tb, mrow, mcol, mexp);
    attn_tile<HD, MT, 3, MASK>(q_h, k_h, v_h, o_h, mbase + 3, lane, tb, mrow, mcol, mexp);
  }
}

template <int HD, int HPC, int NWARPS>
struct CoreCfg {
  static constexpr int HD_PAD = (HD + 15) / 16 * 16;
  static constexpr int LDS = HD_PAD + 8;            // fp16 elements per smem row (pad breaks ldmatrix bank conflicts)
  static constexpr int NT = NWARPS * 32;
  static constexpr int WPH = NWARPS / HPC;          // warps per head
  static constexpr int MT = 4 / WPH;                // 16-row query tiles per warp
  static constexpr int SEG = HPC * HD;              // contiguous fp16 per token per q/k/v segment handled by one unit
  static constexpr int VEC = (HD % 8 == 0) ? 8 : 4; // elements per copy (16 or 8 bytes); never straddles a head
  static constexpr int VPH = HD / VEC;              // vectors per head
  static constexpr int VPS = SEG / VEC;             // vectors per segment
  static constexpr int NVEC = 64 * 3 * VPS;         // gather copies per unit
  static constexpr int SLOTS = (NVEC + NT - 1) / NT;
  static constexpr int OVEC = 64 * VPS;             // scatter vectors per unit
  static constexpr int OSLOTS = (OVEC + NT - 1) / NT;
  // Heads are 64 * LDS elements plus a skew apart and q / k / v a further 64 bytes, so that the head-wise gather copies
  // and the O scatter of one warp spread over the banks instead of piling onto the ones a multiple of 128 bytes apart.
  static constexpr int HS = 64 * LDS + (HD * 2 + 15) / 16 * 8;   // fp16 elements between consecutive heads
  static constexpr int MAT = HPC * HS + 32;         // fp16 elements of one of q / k / v in a stage
  static constexpr int STAGE_BYTES = 3 * MAT * 2;
  static constexpr int TBL = 232;                   // 225 padded
  static constexpr int SMEM = 2 * STAGE_BYTES + 8 * TBL * 4;   // table for up to... see launch (heads <= 8 staged per group)
  static_assert(WPH >= 1 && WPH <= 4 && WPH * HPC == NWARPS && MT * WPH == 4, "warp / head split");
  static_assert(HD % VEC == 0, "head_dim must be a multiple of 4");
};

template <int HD, int HPC, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) attn_core_kernel(const AttnCoreArgs p, const int units) {
  using K = CoreCfg<HD, HPC, NWARPS>;
  constexpr int LDS = K::LDS, NT = K::NT, MT = K::MT, VEC = K::VEC, VPH = K::VPH, VPS = K::VPS, TBL = K::TBL;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __half* stage0 = reinterpret_cast<__half*>(smem_raw);
  float* sTbl = reinterpret_cast<float*>(smem_raw + 2 * K::STAGE_BYTES);   // [heads <= 8][TBL], already times log2(e)
  __shared__ long long sRow[3][64];   // global row of each window token: previous / current / next unit

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nWc = p.W >> 3, nWr = p.H >> 3;
  const int nW = nWr * nWc;
  const int groups = p.heads / HPC;          // head groups per window

  pdl_wait();
  pdl_launch_dependents();
  // ---- one-time setup: bias table of every head, pad columns of both stages (cp.async never touches them)
  for (int i = tid; i < p.heads * 225; i += NT) {
    const int e = i / p.heads, h = i - e * p.heads;
    sTbl[h * TBL + e] = __ldg(p.bias_table + i) * LOG2E;
  }
  if constexpr (K::HD_PAD != HD) {
    constexpr int PADW = (K::HD_PAD - HD) / 2;  // half2 words per row
    for (int i = tid; i < 2 * 3 * HPC * 64 * PADW; i += NT) {
      const int row = i / PADW, w = i - row * PADW;     // row over [stage][q|k|v][head][token]
      const int t = row & 63, h = (row >> 6) % HPC, mat = (row >> 6) / HPC;   // mat = stage * 3 + (q|k|v)
      // V: column HD := 1.0 (softmax denominator through the MMA); everything else 0
      reinterpret_cast<uint32_t*>(stage0 + static_cast<size_t>(mat) * K::MAT + h * K::HS + t * LDS + HD)[w] =
          (mat % 3 == 2 && w == 0) ? 0x00003C00u : 0u;
    }
  }
  // ---- per-thread copy slots (window independent)
  int g_src[K::SLOTS], g_dt[K::SLOTS];   // g_dt = smem byte offset within a stage | token << 18 | valid << 24
#pragma unroll
  for (int s = 0; s < K::SLOTS; ++s) {
    const int i = tid + s * NT;
    const int t = i / (3 * VPS);
    const int rem = i - t * (3 * VPS);
    const int seg = rem / VPS, vv = rem - seg * VPS;
    const int h = vv / VPH, v = vv - h * VPH;
    g_src[s] = seg * p.C + vv * VEC;
    g_dt[s] = i < K::NVEC ? (((seg * K::MAT + h * K::HS + t * LDS + v * VEC) * 2) | (t << 18) | (1 << 24)) : 0;   // only the last slot can be partial
  }

  auto unit_rows = [&](int u, int st) {   // 64 threads: global row of each window token (st: row-map buffer)
    if (tid < 64) {
      const int win = u / groups;
      long long row;
      if (p.windowed_input) {
        row = static_cast<long long>(win) * 64 + tid;
      } else {
        const int b = win / nW;
        const int wimg = win - b * nW;
        const int wr = wimg / nWc, wc = wimg - wr * nWc;
        const int r = (wr * 8 + (tid >> 3) + p.shift) % p.H;
        const int c = (wc * 8 + (tid & 7) + p.shift) % p.W;
        row = (static_cast<long long>(b) * p.H + r) * p.W + c;
      }
      sRow[st][tid] = row;
    }
  };
  const long long ld = p.ld;
  auto unit_gather = [&](int u, int st, int rb) {
    const int hg = u % groups;
    const __half* base = p.qkv + hg * K::SEG;
    const uint32_t sbase = smem_u32(stage0) + st * K::STAGE_BYTES;
#pragma unroll
    for (int s = 0; s < K::SLOTS; ++s)
      if (s + 1 < K::SLOTS || g_dt[s] != 0)
        cp_async<VEC * 2>(sbase + (g_dt[s] & 0x3ffff), base + sRow[rb][(g_dt[s] >> 18) & 63] * ld + g_src[s]);
    cp_async_commit();
  };

  int u = blockIdx.x;
  if (u < units) unit_rows(u, 0);
  __syncthreads();
  if (u < units) unit_gather(u, 0, 0);
  int st = 0, rb = 0;                      // smem stage / row-map buffer of the current unit
  const int h = warp % HPC;                // this warp's head within the group and its first 16-row query tile
  const int mbase = (warp / HPC) * MT;
  const int g = lane >> 2, tq = lane & 3;
  float tb[2][2 * MT + 7];
  int tb_hg = -1;
  for (; u < units; u += gridDim.x, st ^= 1, rb = (rb == 2 ? 0 : rb + 1)) {
    const int un = u + static_cast<int>(gridDim.x);
    const int rbn = rb == 2 ? 0 : rb + 1;  // last read two units ago (its scatter), before the previous sync (A)
    if (un < units) unit_rows(un, rbn);
    __syncthreads();                       // (A) next unit's row map visible; every warp is done with stage st^1
    if (un < units) unit_gather(un, st ^ 1, rbn);
    else cp_async_commit();
    cp_async_wait<1>();                    // this thread's copies of the current unit have landed
    __syncthreads();                       // (B) ... and everybody else's
    __half* sQ = stage0 + static_cast<size_t>(st) * (K::STAGE_BYTES / 2);
    __half* sK = sQ + K::MAT;
    __half* sV = sK + K::MAT;
    const int win = u / groups;
    const int hg = u - win * groups;
    {
      const __half* q_h = sQ + h * K::HS;
      const __half* k_h = sK + h * K::HS;
      const __half* v_h = sV + h * K::HS;
      __half* o_h = sQ + h * K::HS;
      if (hg != tb_hg) {   // this thread's slice of the relative-position bias (constant per CTA when groups | gridDim.x)
        const float* tbl = sTbl + (hg * HPC + h) * TBL;
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int k = 0; k < 2 * MT + 7; ++k) tb[e][k] = tbl[(k + 2 * mbase) * 15 + (g - 2 * tq - e + 7)];
        tb_hg = hg;
      }
      if (p.mask_mode == 2) {
        const float* mexp = p.mask + static_cast<long long>(win % p.mask_nw) * 4096;
        attn_tiles<HD, MT, 2>(q_h, k_h, v_h, o_h, mbase, lane, tb, false, false, mexp);
      } else {
        const int wimg = win % nW;
        const int wr = wimg / nWc, wc = wimg - wr * nWc;
        const bool mrow = p.mask_mode == 1 && wr == nWr - 1;
        const bool mcol = p.mask_mode == 1 && wc == nWc - 1;
        if (mrow || mcol) attn_tiles<HD, MT, 1>(q_h, k_h, v_h, o_h, mbase, lane, tb, mrow, mcol, nullptr);
        else attn_tiles<HD, MT, 0>(q_h, k_h, v_h, o_h, mbase, lane, tb, false, false, nullptr);
      }
    }
    __syncthreads();                       // (C) O rows of every head parked in sQ
    // ---- scatter O rows back (heads are concatenated in order, :135)
#pragma unroll
    for (int s = 0; s < K::OSLOTS; ++s) {
      const int i = tid + s * NT;
      if (i < K::OVEC) {
        const int t = i / VPS, vv = i - t * VPS;
        const int h = vv / VPH, v = vv - h * VPH;
        const __half* src = sQ + h * K::HS + t * LDS + v * VEC;
        __half* dst = p.out + sRow[rb][t] * p.ldo + hg * K::SEG + vv * VEC;
        if constexpr (VEC == 8) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
        else *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
      }
    }
  }
  cp_async_wait<0>();
}

template <int HD, int HPC, int NWARPS>
int launch_core(const AttnCoreArgs& a, int64_t windows, cudaStream_t stream) {
  using K = CoreCfg<HD, HPC, NWARPS>;
  if (a.heads > 8) return fail(SUNET_E_SHAPE, "attn: at most 8 heads are staged (got %d)", a.heads);
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(attn_core_kernel<HD, HPC, NWARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    once.done();
  }
  const int sms = device_sms();
  const int64_t units = windows * (a.heads / HPC);
  if (units > 0x7fffffff) return fail(SUNET_E_SHAPE, "attn: too many (window, head-group) units");
  const unsigned grid = static_cast<unsigned>(units < sms ? units : sms);
  SUNET_CUDA(launch_pdl(attn_core_kernel<HD, HPC, NWARPS>, dim3(grid), dim3(K::NT), K::SMEM, stream, a, static_cast<int>(units)));
  return 0;
}

}  // namespace

int attn_core_launch(const AttnCoreArgs& a, cudaStream_t stream) {
  if (a.C % a.heads != 0) return fail(SUNET_E_SHAPE, "attn: C=%d not divisible by heads=%d", a.C, a.heads);
  if (a.heads % 4 != 0) return fail(SUNET_E_SHAPE, "attn: heads=%d must be a multiple of 4", a.heads);
  if (a.H % 8 || a.W % 8) return fail(SUNET_E_SHAPE, "attn: token grid %dx%d must be a multiple of the 8x8 window", a.H, a.W);
  if (a.ld % 8 || a.ldo % 8 || (reinterpret_cast<uintptr_t>(a.qkv) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return fail(SUNET_E_ALIGN, "attn: qkv/out need 16-byte aligned rows");
  if (a.mask_mode == 2 && (a.mask == nullptr || a.mask_nw <= 0)) return fail(SUNET_E_ARG, "attn: explicit mask missing");
  if (a.mask_mode == 1 && a.shift != 4) return fail(SUNET_E_ARG, "attn: the closed-form mask is built for shift 4 (window 8)");
  const int hd = a.C / a.heads;
  const int64_t windows = a.windowed_input ? a.num_windows : static_cast<int64_t>(a.B) * (a.H / 8) * (a.W / 8);
  if (windows <= 0 || windows > 0x7fffffff) return fail(SUNET_E_SHAPE, "attn: bad window count");
  const bool h8 = a.heads % 8 == 0;
  switch (hd) {
    case 12: return h8 ? launch_core<12, 8, 16>(a, windows, stream) : launch_core<12, 4, 16>(a, windows, stream);
    case 16: return h8 ? launch_core<16, 8, 16>(a, windows, stream) : launch_core<16, 4, 16>(a, windows, stream);
    case 24: return launch_core<24, 4, 16>(a, windows, stream);
    case 32: return launch_core<32, 4, 16>(a, windows, stream);
    case 48: return launch_core<48, 4, 16>(a, windows, stream);
    case 96: return launch_core<96, 2, 8>(a, windows, stream);
    default: return fail(SUNET_E_SHAPE, "attn: head_dim %d not instantiated (12/16/24/32/48/96)", hd);
  }
}

}  // namespace sunet
