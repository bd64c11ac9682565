// Pixel-shuffle branch of the final x4 Dual up-sample fused with the folded output taps (SUNet_detail.py:354-358 up_p, :742-753):
//
//   P  = PReLU(T W_p0^T)                 [M tokens][16 sub-pixels x 96]      (up_p[0..1]; W_p0 rows pre-ordered (ij, c) at pre-pack)
//   Q  = P.view(M * 16, 96) G_p^T        [M * 16 hi-res pixels][16 tap maps] (up_p[3], conv and the 3x3 output conv folded into G_p)
//
// The two GEMMs used to exchange the [M][1536] fp16 tensor through HBM (805 MB written and read back at B = 64).  Here one
// persistent CTA per SM walks 128-token tiles and the hidden activation never leaves the SM - the dataflow of mlp_fused.cu with
// sub-pixels as "hidden chunks": per sub-pixel ij
//   MMA1 : H   = T_tile * W_p0[ij]^T     (tcgen05, N = 96, fp32 in TMEM, double-buffered, issued two sub-pixels ahead)
//   EPI  : G   = fp16(PReLU(H))          TMEM -> registers -> SW128 smem (A operand of MMA2)
//   MMA2 : Q[:, 16 ij .. 16 ij + 15] = G * G_p^T    (tcgen05, N = 16; G_p stays resident in smem)
// After the 16 sub-pixels the 128 x 256 fp32 tile of tap values does NOT go to HBM (it used to: 268 MB written, 285 MB re-read by the
// 9-tap stencil).  The thread that owns (token, sub-pixel row sy) first adds the bilinear branch - its four sub-pixels' interpolated
// tap maps from the per-token Rb (SUNet_detail.py:360-362, evaluated exactly as the stencil did: rows, then columns) - and then sums
// its 4 x 9 tap values into the 3 x 6 patch of OUTPUT pixels they land on (output rows 4 ty + sy - 1 .. + 1, columns 4 tx - 1 .. 4 tx + 4).
// These "strips" (96 bytes per thread, 100 MB per batch of 64) are all that leaves the SM; tail_finish_kernel adds the 3 - 6 strips
// that overlap each output pixel in a fixed order.
#include "tail_fused.cuh"

#include "elementwise.cuh"

#include <stdio.h>
#include <stdlib.h>

#include "device.h"
#include "error.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

namespace {

constexpr int TILE_M = 128;
constexpr int E = 96;                   // embed dim of the tail (training.yaml EMB_DIM)
constexpr int SUB = 16;                 // sub-pixels of the x4 pixel shuffle
constexpr int NT = 16;                  // folded tap maps per sub-pixel (9 taps x out_chans, padded to 16)
constexpr int KBYTES = TILE_M * 128;    // one [128 rows][64 fp16] SW128 k-block
constexpr int W1KB = E * 128;           // one [96 rows][64] k-block of a sub-pixel's weights
constexpr int EPI_WARPS = 16;
#ifndef SUNET_TAIL_FC1_SPLIT
#define SUNET_TAIL_FC1_SPLIT 1   // measured (tools/ab_tail_fc1.sh): 207 -> 182 us
#endif
// TMA producer, fc1 issuer, 16 epilogue warps, fc2 issuer (+ with SUNET_TAIL_FC1_SPLIT a second fc1 issuer: even / odd sub-pixels)
constexpr int THREADS = 64 + EPI_WARPS * 32 + 32 + (SUNET_TAIL_FC1_SPLIT ? 32 : 0);
constexpr int FC2_WARP = 2 + EPI_WARPS;
constexpr int FC1B_WARP = FC2_WARP + 1;
constexpr int R1 = 4;                   // weight ring stages
constexpr int OFF_X = 0;                // 2 token-tile buffers x 2 k-blocks
constexpr int OFF_HS = OFF_X + 2 * 2 * KBYTES;
constexpr int OFF_R1 = OFF_HS + 2 * 2 * KBYTES;
constexpr int OFF_GP = OFF_R1 + R1 * W1KB;       // [2 k-blocks][16 rows][64]
constexpr int RB_MAX_W = 128;                    // widest token grid the staged bilinear taps are sized for
constexpr int RB_TOKENS = TILE_M + 2 * RB_MAX_W + 2;
constexpr int OFF_RB = OFF_GP + 2 * NT * 128;    // [tokens of the tile +- one image row][12] fp32: the used 9 (+3) of the 16 tap maps of Rb
constexpr int SMEM = OFF_RB + RB_TOKENS * 48 + 1024;
constexpr uint32_t TM_Y = 0;            // Q tile: 256 columns
constexpr uint32_t TM_H = 256;          // H[b] at 256 + 128 b
static_assert(SMEM <= 227 * 1024, "shared memory budget");

#ifndef SUNET_KERNEL_TIMING
#define SUNET_KERNEL_TIMING 0
#endif
#if SUNET_KERNEL_TIMING
#define TF_T(i) do { if (p.timing) { const long long _t = clock64(); tacc[i] += _t - tq0; tq0 = _t; } } while (0)
#else
#define TF_T(i) do { } while (0)
#endif

// align_corners=False taps of the x4 bilinear up-sample along one axis (nn.Upsample, SUNet_detail.py:351 / :362); the same
// expression as bilinear_tap() in elementwise.cu, so the interpolation weights are bit-identical to the stencil path's
__device__ __forceinline__ void tail_bilinear_tap(int d, int n, int& i0, int& i1, float& lam) {
  const float src = fmaxf((d + 0.5f) / 4 - 0.5f, 0.f);
  i0 = static_cast<int>(src);
  i1 = min(i0 + 1, n - 1);
  lam = src - i0;
}

struct Params {
  long long* timing;    // SUNET_KERNEL_TIMING builds: [grid][16 warps][8] phase cycles
  const float* slope;   // PReLU slope (device scalar)
  float* strips;        // [M / 32][4 sub-pixel rows][32 tokens][3][8] fp32 partial output sums (see tail_finish_kernel)
  const float* Rb;      // [M][16] fp32 tap maps of the bilinear branch at token resolution
  int H, W;             // token grid
  int64_t M;
  int64_t tiles;
  int split;            // 1: fc2 MMAs are issued by their own thread (warp FC2_WARP), 0: interleaved with fc1 by warp 1
};

// token m (image-order index, < 2^31) and sub-pixel row sy -> its column tx, the token index of its image's first token, and the
// vertical bilinear taps of hi-res row 4 ty + sy (32-bit arithmetic: 64-bit divisions cost ~150 instructions apiece)
__device__ __forceinline__ unsigned tail_geo(const Params& p, unsigned m, int sy, int& tx, int& y0, int& y1, float& ly) {
  const unsigned hw = static_cast<unsigned>(p.H * p.W);
  const unsigned b = m / hw;
  const unsigned rem = m - b * hw;
  const unsigned ty = rem / static_cast<unsigned>(p.W);
  tx = static_cast<int>(rem - ty * p.W);
  tail_bilinear_tap(4 * static_cast<int>(ty) + sy, p.H, y0, y1, ly);
  return b * hw;
}

__global__ void __launch_bounds__(THREADS, 1)
    tail_up_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                         const __grid_constant__ CUtensorMap tmGp, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full[2], x_empty[2], gp_full;
  __shared__ __align__(8) uint64_t r1_full[R1], r1_empty[R1];
  __shared__ __align__(8) uint64_t h_full[2], h_empty[2], g_done[2], hs_empty[2], y_full, y_empty;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmGp);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], (SUNET_TAIL_FC1_SPLIT && p.split) ? 2 : 1);
      mbar_init(&h_full[i], 1);
      mbar_init(&h_empty[i], EPI_WARPS / 2);   // the epilogue warps work in two groups of 8, group b on the chunks of buffer b
      mbar_init(&g_done[i], EPI_WARPS / 2);
      mbar_init(&hs_empty[i], 1);
    }
    mbar_init(&gp_full, 1);
    mbar_init(&y_full, 1);
    mbar_init(&y_empty, EPI_WARPS);
    for (int i = 0; i < R1; ++i) { mbar_init(&r1_full[i], 1); mbar_init(&r1_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(&gp_full, 2 * NT * 128);
      for (int kb = 0; kb < 2; ++kb) tma_load_2d(smem + OFF_GP + kb * NT * 128, &tmGp, &gp_full, kb * 64, 0);
      uint32_t i1 = 0;
      int lt = 0;
      auto load_x = [&](int64_t tile, int ltile) {
        const int xb = ltile & 1;
        mbar_wait(&x_empty[xb], ((ltile >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&x_full[xb], 2 * KBYTES);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(smem + OFF_X + (xb * 2 + kb) * KBYTES, &tmX, &x_full[xb], kb * 64, static_cast<int>(tile * TILE_M));
      };
      if (static_cast<int64_t>(blockIdx.x) < p.tiles) load_x(blockIdx.x, 0);
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        if (tile + gridDim.x < p.tiles) load_x(tile + gridDim.x, lt + 1);   // one tile ahead
        for (int j = 0; j < SUB; ++j)
          for (int kb = 0; kb < 2; ++kb, ++i1) {
            const int s = i1 % R1;
            mbar_wait(&r1_empty[s], ((i1 / R1) & 1) ^ 1);
            mbar_arrive_expect_tx(&r1_full[s], W1KB);
            tma_load_2d(smem + OFF_R1 + s * W1KB, &tmW1, &r1_full[s], kb * 64, j * E);
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues: a
    // single-lane branch makes every descriptor thread-divergent -> R2UR + ELECT loop per tcgen05.mma, see gemm_tcgen05.cu)
    {
      const uint32_t idesc1 = umma_idesc_f16(TILE_M, E);
      const uint32_t idesc2 = umma_idesc_f16(TILE_M, NT);
      uint32_t i1 = 0, g = 0;
      int lt = 0;
      uint32_t r1_ok = mbar_test(&r1_full[0], 0);
      auto fc1 = [&](int xb, uint32_t gg, bool last) {   // H[gg & 1] = T * W_p0[sub-pixel]^T
        const uint32_t hb = gg & 1, use = gg >> 1;
        if (use > 0) mbar_wait(&h_empty[hb], (use - 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + TM_H + hb * 128;
        for (int kb = 0; kb < 2; ++kb) {
          const int s = i1 % R1;
          mbar_wait_hint(&r1_full[s], (i1 / R1) & 1, r1_ok);
          ++i1;
          r1_ok = (SUNET_TAIL_FC1_SPLIT && kb == 1) ? 0u : mbar_test(&r1_full[i1 % R1], (i1 / R1) & 1);   // (split: the next k-block is the other thread's)
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + OFF_X + (xb * 2 + kb) * KBYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + OFF_R1 + s * W1KB));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (kb == 0 || k < 2)   // K = 96 = 64 + 32
                umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&r1_empty[s]);
            if (kb == 1) {
              tc_commit(&h_full[hb]);
              if (last) tc_commit(&x_empty[xb]);
            }
          }
          __syncwarp();
        }
      };
      auto fc2 = [&](uint32_t gg, int j) {   // Q[:, 16 j .. 16 j + 15] = G_j * G_p^T
        const uint32_t hb = gg & 1;
        mbar_wait(&g_done[hb], (gg >> 1) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + TM_Y + j * NT;
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem + OFF_HS + (hb * 2 + kb) * KBYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + OFF_GP + kb * NT * 128));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (kb == 0 || k < 2)
                umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&hs_empty[hb]);
          if (j == SUB - 1) tc_commit(&y_full);
        }
        __syncwarp();
      };
      if (p.split) {
        // fc1 only: runs ahead as far as the two H buffers allow; fc2 has its own issuing thread, so neither waits behind the
        // other's barrier round trips (an mbarrier wait costs ~200 clk even on a completed phase)
        for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
          const int xb = lt & 1;
          mbar_wait(&x_full[xb], (lt >> 1) & 1);
          tc_fence_after();
#if SUNET_TAIL_FC1_SPLIT
          for (int j = 0; j < SUB; j += 2) { i1 = 2 * (g + j); fc1(xb, g + j, j == SUB - 2); }   // even sub-pixels; warp FC1B_WARP takes the odd ones
#else
          for (int j = 0; j < SUB; ++j) fc1(xb, g + j, j == SUB - 1);
#endif
          g += SUB;
        }
      } else {
      mbar_wait(&gp_full, 0);
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        const int xb = lt & 1;
        mbar_wait(&x_full[xb], (lt >> 1) & 1);
        tc_fence_after();
        fc1(xb, g, false);
        fc1(xb, g + 1, false);
        for (int j = 0; j < SUB; ++j) {
          if (j + 2 < SUB) fc1(xb, g + j + 2, j + 2 == SUB - 1);
          if (j == 0) { mbar_wait(&y_empty, (lt & 1) ^ 1); tc_fence_after(); }   // the previous tile's Q has been drained
          fc2(g + j, j);
        }
        g += SUB;
      }
      }
    }
#if SUNET_TAIL_FC1_SPLIT
  } else if (warp == FC1B_WARP) {
    // ------------------------------------------------------------------ second fc1 issuer: the odd sub-pixels (H buffer 1)
    if (p.split) {   // (whole warp, one elected lane issues)
      const uint32_t idesc1 = umma_idesc_f16(TILE_M, E);
      uint32_t g = 0;
      int lt = 0;
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        const int xb = lt & 1;
        mbar_wait(&x_full[xb], (lt >> 1) & 1);
        tc_fence_after();
        for (int j = 1; j < SUB; j += 2) {
          const uint32_t gg = g + j, hb = 1, use = gg >> 1;
          if (use > 0) mbar_wait(&h_empty[hb], (use - 1) & 1);
          tc_fence_after();
          const uint32_t d = tmem_base + TM_H + hb * 128;
          for (int kb = 0; kb < 2; ++kb) {
            const uint32_t i1 = 2 * gg + kb;
            const int s = i1 % R1;
            mbar_wait(&r1_full[s], (i1 / R1) & 1);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem + OFF_X + (xb * 2 + kb) * KBYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + OFF_R1 + s * W1KB));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (kb == 0 || k < 2)
                  umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
              tc_commit(&r1_empty[s]);
              if (kb == 1) {
                tc_commit(&h_full[hb]);
                if (j == SUB - 1) tc_commit(&x_empty[xb]);
              }
            }
            __syncwarp();
          }
        }
        g += SUB;
      }
    }
#endif
  } else if (warp == FC2_WARP) {
    // ------------------------------------------------------------------ fc2 issuer (split mode)
    if (p.split) {   // (whole warp, one elected lane issues)
      const uint32_t idesc2 = umma_idesc_f16(TILE_M, NT);
      uint32_t g = 0;
      int lt = 0;
      mbar_wait(&gp_full, 0);
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        for (int j = 0; j < SUB; ++j) {
          if (j == 0) { mbar_wait(&y_empty, (lt & 1) ^ 1); tc_fence_after(); }   // the previous tile's Q has been drained
          const uint32_t gg = g + j, hb = gg & 1;
          mbar_wait(&g_done[hb], (gg >> 1) & 1);
          tc_fence_after();
          const uint32_t d = tmem_base + TM_Y + j * NT;
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
              const uint64_t adesc = umma_desc_sw128(smem_u32(smem + OFF_HS + (hb * 2 + kb) * KBYTES));
              const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + OFF_GP + kb * NT * 128));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (kb == 0 || k < 2)
                  umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit(&hs_empty[hb]);
            if (j == SUB - 1) tc_commit(&y_full);
          }
          __syncwarp();
        }
        g += SUB;
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may touch
    const int quarter = e >> 2;        // column quarter
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const float slope = __ldg(p.slope);
    uint32_t g = 0;
    int lt = 0;
    uint32_t h_ok = 0;
#if SUNET_KERNEL_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tq0 = clock64();
#endif
    // Ping-pong: warps of column quarters 0-1 take the even sub-pixels (H / G buffer 0), quarters 2-3 the odd ones (buffer 1), each
    // warp 48 of the 96 columns.  The per-chunk chain (accumulator ready -> TMEM load -> PReLU -> smem -> fence -> arrive) of one
    // group runs under the other group's, instead of all 16 warps waiting through the same round trips together.
    const uint32_t grp = static_cast<uint32_t>(quarter >> 1), half = static_cast<uint32_t>(quarter & 1);
    for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
      // The output pass reads the bilinear branch's tap maps of the 2 x 3 tokens around every (token, sub-pixel row): the tokens of this
      // tile and of one image row either side are staged in shared memory with coalesced 16-byte copies (a direct per-thread read
      // touches 16 lines per instruction: 4.6k L1 wavefronts per tile, measured as +3k clk per tile), 16 sub-pixels ahead of their use
      const int rb_lo = static_cast<int>(max(tile * TILE_M - p.W - 1, static_cast<int64_t>(0)));
      {
        const int rb_hi = static_cast<int>(min(tile * TILE_M + TILE_M + p.W + 1, p.M));
        const int n4 = (rb_hi - rb_lo) * 3;
        const float4* src = reinterpret_cast<const float4*>(p.Rb + static_cast<int64_t>(rb_lo) * NT);
        for (int i = e * 32 + lane; i < n4; i += EPI_WARPS * 32) {
          const int tok = i / 3, part = i - tok * 3;
          const float4 v = __ldg(src + tok * 4 + part);
          sts128(smem_u32(smem + OFF_RB) + i * 16, make_uint4(__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)));
        }
      }
      for (int j = static_cast<int>(grp); j < SUB; j += 2) {
        const uint32_t gg = g + static_cast<uint32_t>(j);
        const uint32_t hb = grp, ph = (gg >> 1) & 1;   // g is a multiple of 16: gg & 1 == grp
        TF_T(7);
        mbar_wait_hint(&h_full[hb], ph, h_ok);
        tc_fence_after();
        TF_T(0);
        uint32_t v[48];
#pragma unroll
        for (int i = 0; i < 6; ++i) tmem_ld8(tmem_base + lane_off + TM_H + hb * 128 + half * 48 + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&v[i * 8]));
        const uint32_t hs_ok = mbar_test(&hs_empty[hb], ph ^ 1);
        tmem_ld_wait();
        TF_T(1);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_empty[hb]);   // the accumulator may be overwritten by the fc1 of sub-pixel gg + 2
        uint4 o[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          __half2* o2 = reinterpret_cast<__half2*>(&o[i]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float a = __uint_as_float(v[i * 8 + 2 * t]), b = __uint_as_float(v[i * 8 + 2 * t + 1]);
            o2[t] = __floats2half2_rn(a >= 0.f ? a : slope * a, b >= 0.f ? b : slope * b);   // PReLU, one shared slope (:356)
          }
        }
        TF_T(2);
        mbar_wait_hint(&hs_empty[hb], ph ^ 1, hs_ok);
        TF_T(3);
        h_ok = mbar_test(&h_full[hb], ph ^ 1);   // this group's next sub-pixel (same buffer, next phase)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int gi = static_cast<int>(half) * 6 + i;   // 16-byte chunk of the 96-wide row: k-block gi >> 3, chunk gi & 7
          sts128(smem_u32(smem + OFF_HS + (hb * 2 + (gi >> 3)) * KBYTES) + row * 128 + ((static_cast<uint32_t>(gi & 7) ^ sw) << 4), o[i]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&g_done[hb]);
        TF_T(4);
      }
      g += SUB;
      // ---- output: Q tile columns [64 quarter, +64) of this row = sub-pixels 4 quarter .. 4 quarter + 3, 16 taps each (9 used)
      const uint32_t stg = smem_u32(smem + OFF_HS) + static_cast<uint32_t>(e) * 4096;
      const int64_t m_warp = tile * TILE_M + q * 32;
      const int64_t m = m_warp + lane;
      const bool live = m_warp < p.M;   // (M is a multiple of 64: whole warps)
      // bilinear branch: vertical taps of this thread's sub-pixel row over the three token columns the four sub-pixels can reach.
      // Nothing here depends on the accumulator: the loads are issued before the wait for the last fc2 and land under it.
      named_bar_sync(2, EPI_WARPS * 32);   // the staged taps of every warp are in place
      float V[3][9] = {};
      int tx = 0;
      if (live) {
        int y0, y1;
        float ly;
        const unsigned tokrow0 = tail_geo(p, static_cast<unsigned>(m), quarter, tx, y0, y1, ly);   // token index of (image, row 0, column 0)
        const float ly1 = 1.f - ly;
        const uint32_t row0 = smem_u32(smem + OFF_RB) + static_cast<uint32_t>(static_cast<int>(tokrow0) + y0 * p.W - rb_lo) * 48u;
        const uint32_t row1 = smem_u32(smem + OFF_RB) + static_cast<uint32_t>(static_cast<int>(tokrow0) + y1 * p.W - rb_lo) * 48u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int xc = min(max(tx - 1 + c, 0), p.W - 1);
          const uint4 ua0 = lds128(row0 + xc * 48), ua1 = lds128(row0 + xc * 48 + 16);
          const uint4 ub0 = lds128(row1 + xc * 48), ub1 = lds128(row1 + xc * 48 + 16);
          float a2, b2;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a2) : "r"(row0 + xc * 48 + 32));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b2) : "r"(row1 + xc * 48 + 32));
          const float4 a0 = make_float4(__uint_as_float(ua0.x), __uint_as_float(ua0.y), __uint_as_float(ua0.z), __uint_as_float(ua0.w));
          const float4 a1 = make_float4(__uint_as_float(ua1.x), __uint_as_float(ua1.y), __uint_as_float(ua1.z), __uint_as_float(ua1.w));
          const float4 b0 = make_float4(__uint_as_float(ub0.x), __uint_as_float(ub0.y), __uint_as_float(ub0.z), __uint_as_float(ub0.w));
          const float4 b1 = make_float4(__uint_as_float(ub1.x), __uint_as_float(ub1.y), __uint_as_float(ub1.z), __uint_as_float(ub1.w));
          const float ra[9] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2};
          const float rb[9] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2};
#pragma unroll
          for (int t = 0; t < 9; ++t) V[c][t] = ra[t] * ly1 + rb[t] * ly;
        }
      }
      mbar_wait(&y_full, lt & 1);
      tc_fence_after();
      TF_T(5);
      uint32_t y[4][9];
#pragma unroll
      for (int sx = 0; sx < 4; ++sx) {
        tmem_ld8(tmem_base + lane_off + TM_Y + quarter * 64 + sx * 16, *reinterpret_cast<uint32_t(*)[8]>(&y[sx][0]));
        tmem_ld1(tmem_base + lane_off + TM_Y + quarter * 64 + sx * 16 + 8, y[sx][8]);
      }
      tmem_ld_wait();
      float strip[3][6];
#pragma unroll
      for (int ry = 0; ry < 3; ++ry)
#pragma unroll
        for (int cx = 0; cx < 6; ++cx) strip[ry][cx] = 0.f;
#pragma unroll
      for (int sx = 0; sx < 4; ++sx) {
        // horizontal taps: sub-pixels 0, 1 interpolate between token columns tx - 1 and tx, sub-pixels 2, 3 between tx and tx + 1
        // (the clamped column loads above make the border cases come out as align_corners=False prescribes)
        int x0, x1;
        float lx;
        tail_bilinear_tap(4 * tx + sx, p.W, x0, x1, lx);
        const float lx1 = 1.f - lx;
        const int c0 = sx < 2 ? 0 : 1;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int t = dy * 3 + dx;
            // tap (dy, dx) of pixel (sy, sx) lands on output pixel (sy - dy + 1, sx - dx + 1)
            float& acc = strip[2 - dy][sx - dx + 2];
            acc += __uint_as_float(y[sx][t]);
            acc = fmaf(V[c0][t], lx1, acc);
            acc = fmaf(V[c0 + 1][t], lx, acc);
          }
      }
      // every TMEM read of this warp is complete: hand the Q accumulator back before the stores
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_empty);
      // staged through this warp's 4 KB slice of the (idle) G buffers (112-byte lane pitch: conflict-free 16-byte stores), then
      // written as one contiguous 3 KB block per warp
#pragma unroll
      for (int ry = 0; ry < 3; ++ry) {
        sts128(stg + lane * 112 + ry * 32, make_uint4(__float_as_uint(strip[ry][0]), __float_as_uint(strip[ry][1]), __float_as_uint(strip[ry][2]), __float_as_uint(strip[ry][3])));
        sts128(stg + lane * 112 + ry * 32 + 16, make_uint4(__float_as_uint(strip[ry][4]), __float_as_uint(strip[ry][5]), 0u, 0u));
      }
      __syncwarp();
      if (live) {
        float* dst = p.strips + ((m_warp >> 5) * 4 + quarter) * (32 * 24);
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const int idx = k * 32 + lane;
          const int tok = idx / 6, part = idx - tok * 6;
          *reinterpret_cast<uint4*>(dst + idx * 4) = lds128(stg + tok * 112 + part * 16);
        }
      }
      named_bar_sync(1, EPI_WARPS * 32);   // every warp is done with its staging slice before the next tile's G stores reuse the buffers
      TF_T(6);
    }
#if SUNET_KERNEL_TIMING
    if (p.timing && lane == 0)
      for (int i = 0; i < 8; ++i) p.timing[(static_cast<long long>(blockIdx.x) * EPI_WARPS + e) * 8 + i] = tacc[i];
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------ strips -> image
// out(Y, X) = sum over the sub-pixel rows y' = Y - 1 .. Y + 1 (in image) of the strip of (token column X / 4, row y') at (ry = Y - y' + 1,
// cx = X % 4 + 1), plus - for the first / last pixel column of a token - the halo column (cx = 5 / 0) of the left / right token's strip;
// this is SUNet_detail.py:753's zero-padded 3x3 conv over conv(cat(up_p, up_b)) with everything linear folded into the tap maps.
// CTA = 8 x 8 tokens = 32 x 32 output pixels; the strips of the 10 x 10 token neighbourhood are staged in shared memory (only the
// parts the tile can reach: the last / first sub-pixel row of the tokens above / below, the halo columns of the tokens left / right),
// token stride 100 words so that the 8 tokens of a pixel row hit 8 different bank groups.  The sum order is fixed (deterministic).
// MODE 0: fp32 NCHW; 1: 8-bit NHWC, rint(clamp(v, 0, 1) * 255); 2: fp32 NCHW + validation epilogue (EvalEpilogue), as tail_stencil.
#ifndef SUNET_TF_STAGE_FLAT
#define SUNET_TF_STAGE_FLAT 0   // 1: the flat-index staging loop of tail_finish_kernel (A/B runs)
#endif
#ifndef SUNET_TF_BATCH
#define SUNET_TF_BATCH 10       // staged loads in flight per thread of tail_finish_kernel: all 10 token rows at once (5: 45.4 us, 10: 37.4 us; the long-scoreboard wait on these loads was the top stall)
#endif
constexpr int TF_BATCH = SUNET_TF_BATCH;
constexpr int TF_T8 = 8;        // tokens per tile side
constexpr int TF_N = TF_T8 + 2; // staged tokens per side
constexpr int TF_STRIDE = 100;  // words per staged token: 4 sub-pixel rows x [3][8] + 4 pad

template <int MODE>
__global__ void __launch_bounds__(256) tail_finish_kernel(const float* __restrict__ S, void* __restrict__ out_v, EvalEpilogue ev, int H, int W) {
  __shared__ __align__(16) float sS[TF_N * TF_N * TF_STRIDE];
  pdl_wait();
  pdl_launch_dependents();
  const int OW = 4 * W, OH = 4 * H;
  const int tiles_x = (W + TF_T8 - 1) / TF_T8, tiles_y = (H + TF_T8 - 1) / TF_T8;
  const int64_t b = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x % (tiles_x * tiles_y);
  const int th0 = (trem / tiles_x) * TF_T8, tw0 = (trem % tiles_x) * TF_T8;
  const int tid = threadIdx.x;
#if SUNET_TF_STAGE_FLAT
  {
    constexpr int TOTAL = TF_N * TF_N * 24, BATCH = 5;   // 16-byte vectors: [token][sub-pixel row][6]
#pragma unroll 1
    for (int i0 = 0; i0 < TOTAL; i0 += 256 * BATCH) {
      float4 q[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {   // all loads of a batch are in flight before the first shared-memory store
        const int i = i0 + k * 256 + tid;
        const int tokl = i / 24, r = i - tokl * 24;
        const int sy = r / 6, v = r - sy * 6;
        const int tr = tokl / TF_N, tc = tokl - tr * TF_N;
        const int ty = th0 - 1 + tr, tx = tw0 - 1 + tc;
        bool need = i < TOTAL && ty >= 0 && ty < H && tx >= 0 && tx < W;
        need = need && (tr != 0 || sy == 3) && (tr != TF_N - 1 || sy == 0) && (tc != 0 || (v & 1) == 1) && (tc != TF_N - 1 || (v & 1) == 0);
        q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need) {
          const int64_t tok = (b * H + ty) * W + tx;
          q[k] = __ldg(reinterpret_cast<const float4*>(S + (((tok >> 5) * 4 + sy) * 32 + (tok & 31)) * 24) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * 256 + tid;
        if (i < TOTAL) {
          const int tokl = i / 24, r = i - tokl * 24;
          *reinterpret_cast<float4*>(sS + tokl * TF_STRIDE + r * 4) = q[k];
        }
      }
    }
  }
#else
  // Staging with the index decomposition hoisted: thread = (staged token column tc, 16-byte vector r of a token's 24), the loop walks
  // the 10 staged token rows - everything but the row test and one address is loop-invariant.  (The flat form above decomposed
  // i -> (token, sub-pixel row, vector) with three divisions per vector, 57% of the kernel's instructions at 78% issue-slot use: ncu r09.)
  if (tid < TF_N * 24) {
    const int tc = tid / 24, r = tid - tc * 24;
    const int sy = r / 6, v = r - sy * 6;
    const int tx = tw0 - 1 + tc;
    const bool col_ok = tx >= 0 && tx < W && (tc != 0 || (v & 1) == 1) && (tc != TF_N - 1 || (v & 1) == 0);
    float* dst = sS + tc * TF_STRIDE + r * 4;
#pragma unroll
    for (int h0 = 0; h0 < TF_N; h0 += TF_BATCH) {
      float4 q[TF_BATCH];
#pragma unroll
      for (int k = 0; k < TF_BATCH; ++k) {   // all loads of a batch are in flight before the first shared-memory store
        const int tr = h0 + k, ty = th0 - 1 + tr;
        const bool need = col_ok && ty >= 0 && ty < H && (tr != 0 || sy == 3) && (tr != TF_N - 1 || sy == 0);
        q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (need) {
          const int64_t tok = (b * H + ty) * W + tx;
          q[k] = __ldg(reinterpret_cast<const float4*>(S + (((tok >> 5) * 4 + sy) * 32 + (tok & 31)) * 24) + v);
        }
      }
#pragma unroll
      for (int k = 0; k < TF_BATCH; ++k) *reinterpret_cast<float4*>(dst + (h0 + k) * TF_N * TF_STRIDE) = q[k];
    }
  }
#endif
  __syncthreads();
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  const int lx_ = tid & 31;
  const int tcl = (lx_ >> 2) + 1, cx = (lx_ & 3) + 1;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int ly_ = (tid >> 5) + 8 * k;
    const int y = 4 * th0 + ly_, x = 4 * tw0 + lx_;
    if (y >= OH || x >= OW) continue;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yl = ly_ + dy + 3;                   // staged sub-pixel row of y' = y + dy - 1 (row 0 = first row of the token above)
      const float* base = sS + ((yl >> 2) * TF_N) * TF_STRIDE + (yl & 3) * 24 + (2 - dy) * 8;
      if ((lx_ & 3) == 0) acc += base[(tcl - 1) * TF_STRIDE + 5];
      acc += base[tcl * TF_STRIDE + cx];
      if ((lx_ & 3) == 3) acc += base[(tcl + 1) * TF_STRIDE];
    }
    if (MODE == 1) {
      static_cast<uint8_t*>(out_v)[(b * OH + y) * OW + x] = static_cast<uint8_t>(__float2int_rn(fminf(fmaxf(acc, 0.f), 1.f) * 255.f));
      continue;
    }
    static_cast<float*>(out_v)[(b * OH + y) * OW + x] = acc;
    if (MODE == 2) {
      // train.py:437-443: luminance target, prob = sigmoid(logits), se = (logits - target)^2, weighted sums, Charbonnier
      const int64_t plane = static_cast<int64_t>(OH) * OW, pix = static_cast<int64_t>(y) * OW + x;
      const float w = ev.weight ? __ldg(ev.weight + b * plane + pix) : 1.f;
      float t;
      if (ev.target_chans == 1) {
        t = __ldg(ev.target + b * plane + pix);
      } else {   // 3 -> 1
        const float* tp = ev.target + b * 3 * plane + pix;
        t = 0.2989f * __ldg(tp) + 0.5870f * __ldg(tp + plane) + 0.1140f * __ldg(tp + 2 * plane);
      }
      if (ev.prob) ev.prob[b * plane + pix] = 1.f / (1.f + expf(-acc));
      const float d = acc - t, se = d * d;
      part[0] += se;
      part[1] += static_cast<double>(se * w);
      part[2] += w;
      part[3] += static_cast<double>(sqrtf(se + ev.eps * ev.eps) * w);
    }
  }
  if (MODE == 2) {
    __shared__ double red[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part[k] += __shfl_xor_sync(0xffffffffu, part[k], o);
      if (lane == 0) red[k][warp] = part[k];
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int wi = 0; wi < 8; ++wi) t += red[threadIdx.x][wi];
      atomicAdd(ev.sums + threadIdx.x, t);
    }
    if (blockIdx.x == 0 && threadIdx.x == 4) atomicAdd(ev.sums + 4, static_cast<double>(gridDim.x / (tiles_x * tiles_y)) * OH * OW);
  }
}

}  // namespace

bool tail_up_fused_supported(int E_, int NT_, int W) { return E_ == E && NT_ == NT && W <= RB_MAX_W; }

int tail_up_fused_launch(const __half* T, const __half* w_p0, const __half* g_p, const float* slope, const float* Rb, float* strips, int B,
                         int H, int W, cudaStream_t stream) {
  const int64_t M = static_cast<int64_t>(B) * H * W;
  if (W > RB_MAX_W) return fail(SUNET_E_SHAPE, "fused tail: token grid wider than %d (got %d)", RB_MAX_W, W);
  if (M <= 0 || M > (int64_t)0x7fffff00 || (static_cast<int64_t>(H) * W) % 64) return fail(SUNET_E_SHAPE, "fused tail: bad token grid %d x %d x %d", B, H, W);
  if ((reinterpret_cast<uintptr_t>(T) & 15) || (reinterpret_cast<uintptr_t>(strips) & 15) || (reinterpret_cast<uintptr_t>(Rb) & 15))
    return fail(SUNET_E_ALIGN, "fused tail: T / Rb / strips must be 16-byte aligned");
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(tail_up_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    once.done();
  }
  const int sms = device_sms();
  alignas(64) CUtensorMap tmX, tmW1, tmGp;
  SUNET_TRY(make_tmap_2d_f16(&tmX, T, E, M, E, TILE_M));
  SUNET_TRY(make_tmap_2d_f16(&tmW1, w_p0, E, SUB * E, E, E));
  SUNET_TRY(make_tmap_2d_f16(&tmGp, g_p, E, NT, E, NT));
  Params prm;
  prm.timing = nullptr;
#if SUNET_KERNEL_TIMING
  static long long* tbuf = nullptr;
  if (getenv("SUNET_TAIL_TIMING")) {
    if (!tbuf) SUNET_CUDA(cudaMalloc(&tbuf, 148 * EPI_WARPS * 8 * sizeof(long long)));
    SUNET_CUDA(cudaMemsetAsync(tbuf, 0, 148 * EPI_WARPS * 8 * sizeof(long long), stream));
    prm.timing = tbuf;
  }
#endif
  prm.slope = slope;
  prm.strips = strips;
  prm.Rb = Rb;
  prm.H = H;
  prm.W = W;
  prm.M = M;
  prm.tiles = (M + TILE_M - 1) / TILE_M;
  static const bool no_split = getenv("SUNET_TAIL_NO_SPLIT") != nullptr;   // read once per process
  prm.split = !no_split;   // measured: 295 -> 224 us (tools/ab_split.sh)
  const unsigned grid = static_cast<unsigned>(prm.tiles < sms ? prm.tiles : sms);
  SUNET_CUDA(launch_pdl(tail_up_fused_kernel, dim3(grid), dim3(THREADS), SMEM, stream, tmX, tmW1, tmGp, prm));
#if SUNET_KERNEL_TIMING
  if (prm.timing) {
    SUNET_CUDA(cudaStreamSynchronize(stream));
    static long long host[148 * EPI_WARPS * 8];
    SUNET_CUDA(cudaMemcpy(host, prm.timing, sizeof(host), cudaMemcpyDeviceToHost));
    static const char* names[8] = {"wait_h", "tmem_ld", "prelu", "wait_hs", "store", "wait_y", "out", "loop"};
    double acc[8] = {0};
    for (unsigned b = 0; b < grid; ++b)
      for (int w = 0; w < EPI_WARPS; ++w)
        for (int i = 0; i < 8; ++i) acc[i] += static_cast<double>(host[(b * EPI_WARPS + w) * 8 + i]);
    fprintf(stderr, "tail_up_fused grid=%u cycles per epilogue warp:", grid);
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %s %.0f", names[i], acc[i] / grid / EPI_WARPS);
    fprintf(stderr, "\n");
  }
#endif
  return 0;
}

int tail_finish(const float* strips, void* out, int out_fmt, const EvalEpilogue* ev, int B, int H, int W, cudaStream_t s) {
  const dim3 grid(static_cast<unsigned>(B) * ((H + TF_T8 - 1) / TF_T8) * ((W + TF_T8 - 1) / TF_T8)), block(256);
  EvalEpilogue e;
  if (ev) {
    if (out_fmt != IMG_F32_NCHW) return fail(SUNET_E_ARG, "tail: the validation epilogue needs fp32 output");
    if (!ev->target || !ev->sums) return fail(SUNET_E_ARG, "tail: validation epilogue without target / sums");
    if (ev->target_chans != 1 && ev->target_chans != 3) return fail(SUNET_E_SHAPE, "tail: target has %d channels, output 1", ev->target_chans);
    e = *ev;
    SUNET_CUDA(launch_pdl(tail_finish_kernel<2>, grid, block, 0, s, strips, out, e, H, W));
  } else if (out_fmt == IMG_U8_NHWC) {
    SUNET_CUDA(launch_pdl(tail_finish_kernel<1>, grid, block, 0, s, strips, out, e, H, W));
  } else {
    SUNET_CUDA(launch_pdl(tail_finish_kernel<0>, grid, block, 0, s, strips, out, e, H, W));
  }
  return 0;
}

}  // namespace sunet
