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
// and after the 16 sub-pixels the 128 x 256 fp32 tile goes out with 16-byte stores (1 KB contiguous per token).
#include "tail_fused.cuh"

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
constexpr int SMEM = OFF_GP + 2 * NT * 128 + 1024;
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

struct Params {
  long long* timing;    // SUNET_KERNEL_TIMING builds: [grid][16 warps][8] phase cycles
  const float* slope;   // PReLU slope (device scalar)
  float* out;           // [M * 16][16] fp32 == [M][256]
  int64_t M;
  int64_t tiles;
  int split;            // 1: fc2 MMAs are issued by their own thread (warp FC2_WARP), 0: interleaved with fc1 by warp 1
};

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
      // ---- output: Q tile columns [64 quarter, +64) of this row = sub-pixels 4 quarter .. 4 quarter + 3, 16 taps each
      mbar_wait(&y_full, lt & 1);
      tc_fence_after();
      TF_T(5);
      // staged through this warp's 4 KB slice of the (idle) G buffers, 32 columns per pass: the warp then writes 4 whole 128-byte
      // row segments per instruction instead of 32 different lines (a thread-per-row store is LSU-bound, measured)
      const uint32_t stg = smem_u32(smem + OFF_HS) + static_cast<uint32_t>(e) * 4096;
      const int64_t m_warp = tile * TILE_M + q * 32;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t y[32];
        tmem_ld32(tmem_base + lane_off + TM_Y + quarter * 64 + c, y);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i)
          sts128(stg + lane * 128 + ((static_cast<uint32_t>(i) ^ static_cast<uint32_t>(lane & 7)) << 4), make_uint4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]));
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int r = kk * 4 + (lane >> 3), ch = lane & 7;
          if (m_warp + r < p.M)
            *reinterpret_cast<uint4*>(p.out + (m_warp + r) * (SUB * NT) + quarter * 64 + c + ch * 4) =
                lds128(stg + r * 128 + ((static_cast<uint32_t>(ch) ^ static_cast<uint32_t>(r & 7)) << 4));
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&y_empty);
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

}  // namespace

bool tail_up_fused_supported(int E_, int NT_) { return E_ == E && NT_ == NT; }

int tail_up_fused_launch(const __half* T, const __half* w_p0, const __half* g_p, const float* slope, float* out, int64_t M, cudaStream_t stream) {
  if (M <= 0 || M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "fused tail: bad row count %lld", (long long)M);
  if ((reinterpret_cast<uintptr_t>(T) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(SUNET_E_ALIGN, "fused tail: T/out must be 16-byte aligned");
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
  prm.out = out;
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

}  // namespace sunet
