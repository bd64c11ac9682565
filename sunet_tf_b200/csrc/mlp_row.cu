// See mlp_row.cuh.  A CTA PAIR (cta_group::2) owns 256 token rows, 128 per CTA; per tile the 4C-wide hidden activation is produced and
// consumed in 128-column chunks and never leaves the SMs:
//
//   TMA  : T tile [128][C] fp16 -> smem (SW128, C/64 k-blocks), once per tile, each CTA its own rows          \  warp 0 (one lane)
//          fc1 weight k-blocks -> ring R1, fc2 weight k-blocks -> ring R2: each CTA loads HALF of the rows      /
//   MMA1 : H    = T * W1h_j^T          (tcgen05 cta_group::2, M = 256, N = 128, K = C; fp32 in both CTAs' TMEM) \  warp 1 of the leader
//   MMA2 : Y   += G_j * W2_j^T         (M = 256, two N = C/2 halves, K = 128)                                   /  (one elected lane)
//   GELU : G_j  = gelu(2 (H + hbias))  TMEM -> regs -> fp16 SW128 smem (A operand of MMA2), own 128 rows        \  warps 2..17 of both
//   OUT  : X    = Y + b2 + R           TMEM -> regs -> staged, whole-row global stores                          /
//
// TMEM: Y takes C = 384 columns, which leaves ONE 128-column H accumulator.  That is enough: the epilogue warps pull H into
// registers (one tcgen05.ld per thread) and hand the accumulator back before they start the GELU math, so fc1 of chunk j+1 runs
// under the GELU pass of chunk j; the tensor pipe sees  fc1(j+1), fc2(j), fc1(j+2), fc2(j+1) ...  back to back (3072 clk of MMA per
// chunk against ~1.8k clk of epilogue).
// Why a pair: with whole rows per CTA every row tile streams the full 2.36 MB of fc1 / fc2 weights, and after the token tile (96 KB)
// and the G buffer (32 KB) only 96 KB of shared memory is left for the weight rings.  A TMA load takes ~3k clk under load, so 96 KB in
// flight sustain 32 B/clk: a single-CTA form of this kernel (M = 128, whole weight blocks per CTA) measured 62 us per launch at
// M = 16384, no faster than the two GEMMs it replaced.  In the pair each CTA stages half of every weight block - the 96 KB of rings
// hold one whole chunk of fc1 AND fc2 weights per CTA, and the L2 -> SM weight traffic halves.
// Barriers: *_full of the TMA rings and the token tile live in the leader (both CTAs' loads complete on them); ring empties, h_full,
// g_empty, y_full exist in both CTAs and are signalled by multicast commits; h_empty, g_full, y_empty live in the leader and collect
// the epilogue warps of both CTAs.
#include "mlp_row.cuh"

#include <stdio.h>
#include <stdlib.h>

#include "act.cuh"
#include "device.h"
#include "error.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

namespace {

constexpr int TILE_M = 128;
constexpr int NC = 128;                 // hidden columns per chunk
constexpr int KBYTES = TILE_M * 128;    // one [128 rows][64 fp16] SW128 k-block
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + EPI_WARPS * 32 + 32;   // TMA producer, fc1 issuer, 16 epilogue warps, fc2 issuer
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int FC2_WARP = 2 + EPI_WARPS;
constexpr int STG_BYTES = 2048;         // per epilogue warp: 32 rows x 32 fp16

template <int C>
struct Cfg {
  static constexpr int HID = 4 * C;
  static constexpr int NCH = HID / NC;          // hidden chunks per tile
  static constexpr int KB1 = C / 64;            // k-blocks of fc1
  static constexpr int NH = C / 2;              // columns per fc2 MMA instruction
  static constexpr int CQ = C / 4;              // output columns per epilogue column-quarter
  static constexpr int W1ROWS = NC / 2;         // fc1 weight rows (hidden units) each CTA of the pair stages per k-block
  static constexpr int W2ROWS = NH / 2;         // fc2 weight rows (output columns) each CTA stages per (k-block, half)
  static constexpr int W1BYTES = W1ROWS * 128;
  static constexpr int W2BYTES = W2ROWS * 128;
  // Ring slots are coarse - half a chunk each - because every slot hand-over costs the issuing thread ~400 clk of barrier round
  // trips whatever its size (measured with the phase counters: 10 slots per chunk of 4 MMAs each paced the whole kernel at 4.1k clk
  // of issue time per chunk against 3.1k clk of MMA work).
  static constexpr int S1KB = KB1 / 2;          // fc1 ring slot: S1KB k-blocks of [NC/2 rows][64]
  static constexpr int S1BYTES = S1KB * W1BYTES;
  static constexpr int S2BYTES = 2 * W2BYTES;   // fc2 ring slot: one k-block of the chunk, both output halves
  static constexpr int R1 = 2;                  // fc1 weight ring: one whole chunk per CTA
  static constexpr int R2 = 2;                  // fc2 weight ring: one whole chunk per CTA
  static constexpr int OFF_T = 0;
  static constexpr int OFF_G = OFF_T + KB1 * KBYTES;
  static constexpr int OFF_R1 = OFF_G + 2 * KBYTES;
  static constexpr int OFF_R2 = OFF_R1 + R1 * S1BYTES;
  static constexpr int OFF_B2 = OFF_R2 + R2 * S2BYTES;   // float [C]
  static constexpr int SMEM = OFF_B2 + C * 4 + 1024;
  static constexpr uint32_t TM_Y = 0;
  static constexpr uint32_t TM_H = C;
  static_assert(C % 128 == 0 && CQ % 32 == 0 && NH % 32 == 0 && NH <= 256, "unsupported width");
  static_assert(C + NC <= 512, "Y and one H accumulator must fit the 512 TMEM columns");
  static_assert(W1BYTES % 1024 == 0 && W2BYTES % 1024 == 0, "weight slots must start on a swizzle atom");
  static_assert(EPI_WARPS * STG_BYTES <= 2 * KBYTES, "output staging aliases the G buffer");
  static_assert(SMEM <= 227 * 1024 - 192, "shared memory budget (static barriers included)");
};

struct Params {
  const float* hbias;
  const float* b2;
  const float* bp;    // PROJ: attn.proj.bias [C] or null
  const __half* R;    // residual of the second add (plain form) / block input = shortcut of the first add (PROJ form)
  __half* X;
  int64_t M;
  int64_t tiles;      // 128-row tiles; a pair takes tiles 2 c and 2 c + 1 of its cluster-strided sequence
  const void* w1;     // weight images, for the L2 prefetch of the prologue
  const void* w2;
  const void* wp;     // PROJ
  uint32_t wbytes;    // bytes of w1 / w2
  long long* timing;  // optional [grid][19 warps][8] phase cycle counters (timing builds, SUNET_MLP_TIMING)
};

// Phase cycle counters: compiled in only with -DSUNET_KERNEL_TIMING=1 (tools/phase_timing.sh)
#ifndef SUNET_KERNEL_TIMING
#define SUNET_KERNEL_TIMING 0
#endif
#if SUNET_KERNEL_TIMING
#define ROW_T_DECL long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tq0 = clock64(); const long long tstart = tq0
#define ROW_T(i) do { const long long _t = clock64(); tacc[i] += _t - tq0; tq0 = _t; } while (0)
#define ROW_T_FLUSH() do { if (p.timing && lane == 0) { tacc[7] = clock64() - tstart; \
    for (int _i = 0; _i < 8; ++_i) p.timing[(static_cast<long long>(blockIdx.x) * 19 + warp) * 8 + _i] = tacc[_i]; } } while (0)
#else
#define ROW_T_DECL do { } while (0)
#define ROW_T(i) do { } while (0)
#define ROW_T_FLUSH() do { } while (0)
#endif

__device__ __forceinline__ uint32_t stg_off(int row, int ch) { return static_cast<uint32_t>(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)); }

// PROJ: the token tile the kernel loads is the attention output; attn.proj runs first (its [C/2][64] weight k-blocks ride the fc2
// ring, its accumulator is the - still idle - fc2 accumulator), the epilogue warps build x1 = P + bp + shortcut, store it, take the
// LayerNorm statistics (two passes over x1 parked in TMEM, as proj_ln.cu) and write the normalised rows over the token tile as the
// fc1 operand (norm2's affine part is folded into the fc1 weights at pre-pack).  SUNet_detail.py:136, :261-262.
// 19 warps = 5 on one scheduler: 96 registers per thread is what a 16K-register SM sub-partition allows (more is unlaunchable)
template <int C, bool PROJ>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
    mlp_row_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmW1,
                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmWp, const Params p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t t_full, t_empty, r1_full[K::R1], r1_empty[K::R1], r2_full[K::R2], r2_empty[K::R2];
  __shared__ __align__(8) uint64_t h_full, h_empty, g_full[2], g_empty[2], y_full, y_empty, p_full, x1_ready;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();   // 0: leader (issues the MMAs)
  const int64_t pair_id = blockIdx.x >> 1;
  const int64_t n_pairs = gridDim.x >> 1;
  const int64_t pair_tiles = (p.tiles + 1) >> 1;   // 256-row tiles
#if SUNET_KERNEL_TIMING
  unsigned long long gt0 = 0;
  if (p.timing && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmT);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    if (PROJ) tma_prefetch_desc(&tmWp);
    mbar_init(&t_full, 1);
    mbar_init(&t_empty, 1);
    for (int i = 0; i < K::R1; ++i) { mbar_init(&r1_full[i], 1); mbar_init(&r1_empty[i], 1); }
    for (int i = 0; i < K::R2; ++i) { mbar_init(&r2_full[i], 1); mbar_init(&r2_empty[i], 1); }
    mbar_init(&h_full, 1);
    mbar_init(&h_empty, 2 * EPI_WARPS);
    for (int i = 0; i < 2; ++i) { mbar_init(&g_full[i], 2 * EPI_WARPS); mbar_init(&g_empty[i], 1); }
    mbar_init(&y_full, 1);
    mbar_init(&y_empty, 2 * EPI_WARPS);
    mbar_init(&p_full, 1);
    mbar_init(&x1_ready, 2 * EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(&tmem_base_smem, 512);
    tmem_relinquish_pair();
  }
  {
    float* b2s = reinterpret_cast<float*>(smem + K::OFF_B2);
    for (int i = threadIdx.x; i < C; i += THREADS) b2s[i] = p.b2 != nullptr ? __ldg(p.b2 + i) : 0.f;
  }
  if (warp == 2 && lane < (PROJ ? 3 : 2)) {
    // The weights of a block are cold (the model's 200 MB of parameters do not stay in the 126 MB L2 between forwards), and the rings
    // hold one chunk: without this every chunk exposes an HBM round trip (51 us per launch in the model against 44 us).  Every CTA
    // asks for 1/grid of the matrices up front - parameters, so ahead of the dependency wait.
    l2_prefetch_slice(lane == 0 ? p.w1 : (lane == 1 ? p.w2 : p.wp), lane == 2 ? p.wbytes / 4 : p.wbytes);   // Wp is C x C, fc1 / fc2 are 4C x C
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / peer TMA completion
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
#if SUNET_KERNEL_TIMING
  if (p.timing && threadIdx.x == 0) {
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    p.timing[148 * 19 * 8 + blockIdx.x * 4 + 0] = static_cast<long long>(gt0);
    p.timing[148 * 19 * 8 + blockIdx.x * 4 + 1] = static_cast<long long>(gt1);
  }
#endif
  pdl_launch_dependents();
  // only the threads that touch the predecessor's output wait for it: the TMA producers before the first T load (the first chunk
  // of fc1 weights is in flight by then), the epilogue warps before the residual reads

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs), in the MMA issuer's consumption
    // order: W1(0), W1(1), W2(0), W1(2), W2(1), ..., W1(NCH-1), W2(NCH-2), W2(NCH-1)
    if (lane == 0) {
      uint32_t i1 = 0, i2 = 0;
      int lt = 0;
      ROW_T_DECL;
      auto load_w1 = [&](int j, int h) {   // half h of the fc1 weights of chunk j: k-blocks h * S1KB ..
        const int s = i1 % K::R1;
        ROW_T(2);
        mbar_wait(&r1_empty[s], ((i1 / K::R1) & 1) ^ 1);
        ROW_T(0);
        if (rank == 0) mbar_arrive_expect_tx(&r1_full[s], 2 * K::S1BYTES);   // both CTAs' bytes land on the leader's barrier
#pragma unroll
        for (int k = 0; k < K::S1KB; ++k)
          tma_load_2d_pair(smem + K::OFF_R1 + s * K::S1BYTES + k * K::W1BYTES, &tmW1, &r1_full[s], (h * K::S1KB + k) * 64,
                           j * NC + static_cast<int>(rank) * K::W1ROWS);
        ++i1;
      };
      auto load_w2 = [&](int j) {
#pragma unroll 1
        for (int kb2 = 0; kb2 < 2; ++kb2, ++i2) {   // k-block of the chunk; both output halves ride one slot
          const int s = i2 % K::R2;
          ROW_T(2);
          mbar_wait(&r2_empty[s], ((i2 / K::R2) & 1) ^ 1);
          ROW_T(1);
          if (rank == 0) mbar_arrive_expect_tx(&r2_full[s], 2 * K::S2BYTES);
#pragma unroll
          for (int half = 0; half < 2; ++half)
            tma_load_2d_pair(smem + K::OFF_R2 + s * K::S2BYTES + half * K::W2BYTES, &tmW2, &r2_full[s], j * NC + kb2 * 64,
                             half * K::NH + static_cast<int>(rank) * K::W2ROWS);
        }
      };
      auto load_wp = [&](int kb) {   // attn.proj weights, k-block kb: same slot shape as an fc2 k-block (both output halves)
        const int s = i2 % K::R2;
        ROW_T(2);
        mbar_wait(&r2_empty[s], ((i2 / K::R2) & 1) ^ 1);
        ROW_T(1);
        if (rank == 0) mbar_arrive_expect_tx(&r2_full[s], 2 * K::S2BYTES);
#pragma unroll
        for (int half = 0; half < 2; ++half)
          tma_load_2d_pair(smem + K::OFF_R2 + s * K::S2BYTES + half * K::W2BYTES, &tmWp, &r2_full[s], kb * 64,
                           half * K::NH + static_cast<int>(rank) * K::W2ROWS);
        ++i2;
      };
      int pre = 0, prep = 0;
      if (pair_id < pair_tiles) {   // parameters: issued ahead of the dependency wait
        if (PROJ) for (; prep < K::R2; ++prep) load_wp(prep);
        for (; pre < 2; ++pre) load_w1(0, pre);
      }
      pdl_wait();
      for (int64_t pt = pair_id; pt < pair_tiles; pt += n_pairs, ++lt) {
        const int64_t tile = 2 * pt + rank;
        if (lt > 0) mbar_wait(&t_empty, (lt - 1) & 1);   // every fc1 MMA of the previous tile has read the token tiles
        if (rank == 0) mbar_arrive_expect_tx(&t_full, 2 * K::KB1 * KBYTES);
        for (int kb = 0; kb < K::KB1; ++kb)
          tma_load_2d_pair(smem + K::OFF_T + kb * KBYTES, &tmT, &t_full, kb * 64, static_cast<int>(tile * TILE_M));   // rows beyond M: zero fill
        if (PROJ)
          for (int kb = lt == 0 ? prep : 0; kb < K::KB1; ++kb) load_wp(kb);
#pragma unroll 1
        for (int s = 0; s <= K::NCH; ++s) {
          if (s < K::NCH)
            for (int h = (s == 0 && lt == 0) ? pre : 0; h < 2; ++h) load_w1(s, h);
          if (s >= 1) load_w2(s - 1);
        }
      }
      ROW_T_FLUSH();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ fc1 issuer (leader only): the whole warp walks the loop and
    // waits on the barriers, one elected lane issues (a single-lane branch makes every descriptor thread-divergent, gemm_tcgen05.cu).
    // fc1 and fc2 are issued by two warps: with one issuer the wait for the second half of a chunk's GELU output (3.4k clk after the
    // chunk's fc1 completes) sat in front of the NEXT chunk's fc1, whose accumulator had been free for 2.5k clk (4.2k clk per chunk
    // against 3.1k clk of MMA work, phase counters).
    if (rank == 0) {
      const uint32_t idesc1 = umma_idesc_f16(2 * TILE_M, NC);
      uint32_t i1 = 0, g = 0;
      int lt = 0;
      uint32_t r1_ok = 0;
      ROW_T_DECL;
      auto r1_acquire = [&]() -> int {
        const int s = i1 % K::R1;
        ROW_T(5);
        mbar_wait_hint(&r1_full[s], (i1 / K::R1) & 1, r1_ok);
        ROW_T(1);
        ++i1;
        r1_ok = mbar_test(&r1_full[i1 % K::R1], (i1 / K::R1) & 1);   // the next slot, looked up under this slot's MMAs
        tc_fence_after();
        return s;
      };
      for (int64_t pt = pair_id; pt < pair_tiles; pt += n_pairs, ++lt) {
        ROW_T(5);
        if (PROJ) mbar_wait(&x1_ready, lt & 1);   // both CTAs' epilogues have written norm2(x1) over the token tiles
        else mbar_wait(&t_full, lt & 1);
        ROW_T(4);
        tc_fence_after();
#pragma unroll 1
        for (int s = 0; s < K::NCH; ++s) {   // fc1 of chunk s: H = T * W1h_s^T
          const uint32_t gg = g + s;
          ROW_T(5);
          if (gg > 0) mbar_wait(&h_empty, (gg - 1) & 1);
          ROW_T(0);   // both CTAs' epilogues have pulled the previous chunk out of the accumulator
          tc_fence_after();
          const uint32_t d = tmem_base + K::TM_H;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int slot = r1_acquire();
            const uint32_t a0 = smem_u32(smem + K::OFF_T + h * K::S1KB * KBYTES);
            const uint32_t b0 = smem_u32(smem + K::OFF_R1 + slot * K::S1BYTES);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < K::S1KB; ++kk) {
                const uint64_t adesc = umma_desc_sw128(a0 + kk * KBYTES);
                const uint64_t bdesc = umma_desc_sw128(b0 + kk * K::W1BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_ss_pair(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc1, (h > 0 || kk > 0 || k > 0) ? 1u : 0u);
              }
              tc_commit_pair(&r1_empty[slot], 3);
              if (h == 1) {
                tc_commit_pair(&h_full, 3);
                if (s == K::NCH - 1) tc_commit_pair(&t_empty, 3);
              }
            }
            __syncwarp();
          }
        }
        g += K::NCH;
      }
      ROW_T_FLUSH();
    }
  } else if (warp == FC2_WARP) {
    // ------------------------------------------------------------------ fc2 issuer (leader only)
    if (rank == 0) {
      const uint32_t idesc2 = umma_idesc_f16(2 * TILE_M, K::NH);
      uint32_t i2 = 0, g = 0;
      int lt = 0;
      uint32_t r2_ok = 0;
      ROW_T_DECL;
      auto r2_acquire = [&]() -> int {
        const int s = i2 % K::R2;
        ROW_T(5);
        mbar_wait_hint(&r2_full[s], (i2 / K::R2) & 1, r2_ok);
        ROW_T(3);
        ++i2;
        r2_ok = mbar_test(&r2_full[i2 % K::R2], (i2 / K::R2) & 1);
        tc_fence_after();
        return s;
      };
      for (int64_t pt = pair_id; pt < pair_tiles; pt += n_pairs, ++lt) {
        if (PROJ) {   // P = attn_out * Wp^T into the (idle) fc2 accumulator
          ROW_T(5);
          mbar_wait(&t_full, lt & 1);
          if (lt > 0) mbar_wait(&y_empty, (lt - 1) & 1);   // the previous tile's output passes have drained Y
          ROW_T(4);
          tc_fence_after();
#pragma unroll 1
          for (int kb = 0; kb < K::KB1; ++kb) {
            const int slot = r2_acquire();
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_T + kb * KBYTES));
            const uint32_t b0 = smem_u32(smem + K::OFF_R2 + slot * K::S2BYTES);
            if (elect_one()) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint64_t bdesc = umma_desc_sw128(b0 + half * K::W2BYTES);
                const uint32_t d = tmem_base + K::TM_Y + half * K::NH;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_ss_pair(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
              }
              tc_commit_pair(&r2_empty[slot], 3);
              if (kb == K::KB1 - 1) tc_commit_pair(&p_full, 3);
            }
            __syncwarp();
          }
        }
#pragma unroll 1
        for (int j = 0; j < K::NCH; ++j) {   // fc2 of chunk j: Y (+)= G * W2_j^T
          const uint32_t gg = g + j;
          ROW_T(5);
          if (!PROJ && j == 0 && lt > 0) mbar_wait(&y_empty, (lt - 1) & 1);
          ROW_T(4);   // the previous tile's output passes have drained Y
#pragma unroll 1
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            ROW_T(5);
            mbar_wait(&g_full[kb2], gg & 1);   // both CTAs' GELU warps have stored (and proxy-fenced) this k-block of G
            ROW_T(2);
            tc_fence_after();
            const int slot = r2_acquire();
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_G + kb2 * KBYTES));
            const uint32_t b0 = smem_u32(smem + K::OFF_R2 + slot * K::S2BYTES);
            if (elect_one()) {
#pragma unroll
              for (int half = 0; half < 2; ++half) {
                const uint64_t bdesc = umma_desc_sw128(b0 + half * K::W2BYTES);
                const uint32_t d = tmem_base + K::TM_Y + half * K::NH;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16_ss_pair(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2, (j > 0 || kb2 > 0 || k > 0) ? 1u : 0u);
              }
              tc_commit_pair(&r2_empty[slot], 3);
              tc_commit_pair(&g_empty[kb2], 3);
              if (kb2 == 1 && j == K::NCH - 1) tc_commit_pair(&y_full, 3);
            }
            __syncwarp();
          }
        }
        g += K::NCH;
      }
      ROW_T_FLUSH();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs, own 128 rows)
    pdl_wait();
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may touch
    const int quarter = e >> 2;        // column quarter
    const int row = q * 32 + lane;     // row of the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t stg = smem_u32(smem + K::OFF_G + e * STG_BYTES);
    const int cr = lane >> 2, cch = lane & 3;   // copy role: row cr + 8 it, 16-byte chunk cch
    const float* b2s = reinterpret_cast<const float*>(smem + K::OFF_B2);
    uint32_t g = 0;
    int lt = 0;
    ROW_T_DECL;
    for (int64_t pt = pair_id; pt < pair_tiles; pt += n_pairs, ++lt) {
      const int64_t tile = 2 * pt + rank;
      float ln_a = 1.f, ln_b = 0.f;   // PROJ: rstd and -mean * rstd of this thread's row
      if constexpr (PROJ) {
        // ---- x1 = P + bp + shortcut -> X (global) and, as fp16, over the attention-output tile: the RAW x1 is the fc1 operand.
        // norm2 is folded into fc1 (exact algebra, as mlp_fused.cu):  0.5 fc1(LN(x1))_n = rstd (D_n - mean s_n) + c_n  with
        // D = x1 W1hg^T, W1hg = fp16(0.5 W1 gamma), s_n = sum_k W1hg[n,k] (the rounded weights), c_n = 0.5 (b1_n + W1 beta), so the
        // row statistics are not needed before the first GELU pass and their exchange runs under the first fc1 MMAs.
        constexpr int NPASS = K::CQ / 32;
        const int64_t m_base = tile * TILE_M + q * 32;
        const int rows_valid = static_cast<int>(min(static_cast<int64_t>(32), p.M - m_base));
        const int col0 = quarter * K::CQ;
        uint4 rpre[NPASS][4];   // shortcut rows of this warp's 32 x CQ block, whole 64-byte row segments per 4 lanes
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
          const __half* rbase = p.R + m_base * C + col0 + ps * 32;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + cr;
            rpre[ps][it] = r < rows_valid ? *(reinterpret_cast<const uint4*>(rbase + static_cast<int64_t>(r) * C) + cch) : make_uint4(0u, 0u, 0u, 0u);   // plain load: X may alias R
          }
        }
        mbar_wait(&p_full, lt & 1);
        tc_fence_after();
        const uint32_t ty = tmem_base + lane_off + K::TM_Y + col0;
        const uint32_t ts = smem_u32(smem + K::OFF_T) + row * 128;
        float s1 = 0.f, s2 = 0.f, k0 = 0.f;   // shifted sums over this thread's CQ columns (shift = its first value)
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
          const int cc = ps * 32;
#pragma unroll
          for (int it = 0; it < 4; ++it) sts128(stg + stg_off(it * 8 + cr, cch), rpre[ps][it]);
          uint32_t v[32];
          tmem_ld32(ty + cc, v);
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 rr = lds128(stg + stg_off(lane, ch));
            const __half2* r2 = reinterpret_cast<const __half2*>(&rr);
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (p.bp != nullptr) {
              b0 = __ldg(reinterpret_cast<const float4*>(p.bp + col0 + cc + ch * 8));
              b1 = __ldg(reinterpret_cast<const float4*>(p.bp + col0 + cc + ch * 8 + 4));
            }
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 o;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 r = __half22float2(r2[t]);
              const __half2 h = __floats2half2_rn(__uint_as_float(v[ch * 8 + 2 * t]) + bb[2 * t] + r.x,
                                                  __uint_as_float(v[ch * 8 + 2 * t + 1]) + bb[2 * t + 1] + r.y);
              const float2 f = __half22float2(h);   // statistics on the rounded stream values, as a stand-alone LayerNorm reads them
              if (ps == 0 && ch == 0 && t == 0) k0 = f.x;
              const float d0 = f.x - k0, d1 = f.y - k0;
              s1 += d0 + d1;
              s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
              ow[t] = *reinterpret_cast<const uint32_t*>(&h);
            }
            sts128(stg + stg_off(lane, ch), o);
            const int gi = (col0 + cc) / 8 + ch;   // 16-byte chunk of the row inside the [128][C] SW128 tile
            sts128(ts + (gi >> 3) * KBYTES + ((static_cast<uint32_t>(gi & 7) ^ sw) << 4), o);
          }
          __syncwarp();
          __half* xbase = p.X + m_base * C + col0 + cc;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + cr;
            if (r < rows_valid) *(reinterpret_cast<uint4*>(xbase + static_cast<int64_t>(r) * C) + cch) = lds128(stg + stg_off(r, cch));
          }
          __syncwarp();
        }
        // P has been read and x1 is in place: fc1 of the first chunk may start
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&x1_ready, 0);
        // per-quarter (mean, M2) around the quarter's own shift, merged over the four column-quarter warps of this lane quarter
        // (parallel variance) through their - now idle - staging tiles: float2 [32] at offset 0 of each
        const float mq = k0 + s1 * (1.0f / K::CQ);
        const float m2q = s2 - s1 * s1 * (1.0f / K::CQ);
        const uint32_t xq = smem_u32(smem + K::OFF_G + ((q + 2) & 3) * STG_BYTES);   // staging tile of the quarter-0 warp of this lane quarter (e = warp - 2)
        named_bar_sync(2 + q, 128);   // all four warps are done with their staging tiles
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stg + lane * 8), "f"(mq), "f"(m2q) : "memory");
        named_bar_sync(2 + q, 128);
        float mean = 0.f, m2 = 0.f, mqs[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float vx, vy;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(vx), "=f"(vy) : "r"(xq + t * 4 * STG_BYTES + lane * 8) : "memory");
          mqs[t] = vx; mean += vx; m2 += vy;
        }
        mean *= 0.25f;
#pragma unroll
        for (int t = 0; t < 4; ++t) m2 = fmaf(static_cast<float>(K::CQ) * (mqs[t] - mean), mqs[t] - mean, m2);
        ln_a = rsqrtf(fmaxf(m2 * (1.0f / C), 0.f) + 1e-5f);
        ln_b = -mean * ln_a;
        // The staging tiles are part of the G buffer, whose next writers are the GELU stores of chunk 0 - of ANY warp, ~3k clk after the
        // last x1_ready arrive, which precedes these reads: every epilogue warp must be past its reads before any of them moves on.
        named_bar_sync(1, EPI_THREADS);
      }
      // ---- GELU passes
#pragma unroll 1
      for (int j = 0; j < K::NCH; ++j, ++g) {
        mbar_wait(&h_full, g & 1);
        ROW_T(0);
        tc_fence_after();
        // each warp takes 16 columns of BOTH 64-column k-blocks of the chunk and publishes them one after the other, so that the
        // fc2 MMAs of the first k-block start half a GELU pass earlier (the GELU -> fc2 chain is as long as a chunk's MMA time)
        uint32_t v[2][16];
        tmem_ld16(tmem_base + lane_off + K::TM_H + quarter * 16, v[0]);
        tmem_ld16(tmem_base + lane_off + K::TM_H + 64 + quarter * 16, v[1]);
        uint32_t g_ok = g > 0 ? mbar_test(&g_empty[0], (g - 1) & 1) : 1u;   // looked up under the TMEM load / GELU math
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&h_empty, 0);   // the accumulator may be overwritten by the fc1 of the next chunk
        ROW_T(1);
        // u = 0.5 * fc1(T) = D + hbias (fc1 weights and bias are pre-scaled by 0.5); GELU(2u) = u + u * tanh(u * P(u^2))
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t w[8];
          if constexpr (PROJ) {
            const float4* hc4 = reinterpret_cast<const float4*>(p.hbias) + (j * NC + hh * 64 + quarter * 16) / 2;   // (s, c) pairs
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 c2 = __ldg(hc4 + i);   // (s, c) of two consecutive hidden columns; warp-uniform address: broadcast
              const float g0 = gelu_half_arg(fmaf(ln_a, __uint_as_float(v[hh][2 * i + 0]), fmaf(ln_b, c2.x, c2.y)));
              const float g1 = gelu_half_arg(fmaf(ln_a, __uint_as_float(v[hh][2 * i + 1]), fmaf(ln_b, c2.z, c2.w)));
              const __half2 p0 = __floats2half2_rn(g0, g1);
              w[i] = *reinterpret_cast<const uint32_t*>(&p0);
            }
          } else {
            const float4* hb4 = reinterpret_cast<const float4*>(p.hbias + j * NC + hh * 64 + quarter * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 bq = __ldg(hb4 + i);   // warp-uniform address: broadcast
              const float g0 = gelu_half_arg(__uint_as_float(v[hh][4 * i + 0]) + bq.x), g1 = gelu_half_arg(__uint_as_float(v[hh][4 * i + 1]) + bq.y);
              const float g2 = gelu_half_arg(__uint_as_float(v[hh][4 * i + 2]) + bq.z), g3 = gelu_half_arg(__uint_as_float(v[hh][4 * i + 3]) + bq.w);
              const __half2 p0 = __floats2half2_rn(g0, g1), p1 = __floats2half2_rn(g2, g3);
              w[2 * i] = *reinterpret_cast<const uint32_t*>(&p0);
              w[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&p1);
            }
          }
          ROW_T(2);
          if (g > 0) mbar_wait_hint(&g_empty[hh], (g - 1) & 1, g_ok);   // fc2 of the previous chunk has consumed this k-block of G
          ROW_T(3);
          if (hh == 0) g_ok = g > 0 ? mbar_test(&g_empty[1], (g - 1) & 1) : 1u;
          const uint32_t gs = smem_u32(smem + K::OFF_G + hh * KBYTES) + row * 128;
#pragma unroll
          for (int i = 0; i < 2; ++i)
            sts128(gs + (((static_cast<uint32_t>(quarter * 2 + i)) ^ sw) << 4), make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
          // The G stores of this warp are in this SM's shared memory and visible to the async proxy once fence.proxy.async retires; the
          // arrive below is sent after it, and the leader issues the MMAs that read them only after it has received every arrive.  A
          // release.cluster arrive (MEMBAR.ALL.GPU + CCTL.IVALL per call, 1.2k clk on the GELU -> fc2 chain, measured) buys nothing here.
          fence_proxy_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&g_full[hh], 0);
          ROW_T(4);
        }
      }
      // ---- output: Y + b2 + R -> X, staged through the (now idle) G buffer so that global accesses cover whole 64-byte row segments
      {
        const int64_t m_base = tile * TILE_M + q * 32;
        const int rows_valid = static_cast<int>(min(static_cast<int64_t>(32), p.M - m_base));   // <= 0 for a peer tile beyond M
        const int col0 = quarter * K::CQ;
        // the residual rows of this warp's 32 x CQ block, fetched before the wait for the last fc2 (whole 64-byte row segments per 4 lanes)
        constexpr int NPASS = K::CQ / 32;
        uint4 rpre[NPASS][4];
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
          const __half* rbase = (PROJ ? p.X : p.R) + m_base * C + col0 + ps * 32;   // PROJ: x1, stored by this very thread above
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + cr;
            rpre[ps][it] = r < rows_valid ? *(reinterpret_cast<const uint4*>(rbase + static_cast<int64_t>(r) * C) + cch) : make_uint4(0u, 0u, 0u, 0u);   // plain load: X may alias R
          }
        }
        mbar_wait(&y_full, lt & 1);   // every MMA of the tile is complete: G is no longer read
        ROW_T(5);
        tc_fence_after();
#pragma unroll
        for (int ps = 0; ps < NPASS; ++ps) {
          const int cc = ps * 32;
#pragma unroll
          for (int it = 0; it < 4; ++it) sts128(stg + stg_off(it * 8 + cr, cch), rpre[ps][it]);
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_off + K::TM_Y + col0 + cc, v);
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const uint4 rr = lds128(stg + stg_off(lane, ch));
            const __half2* r2 = reinterpret_cast<const __half2*>(&rr);
            const float4 b0 = *reinterpret_cast<const float4*>(b2s + col0 + cc + ch * 8);       // warp-uniform: broadcast
            const float4 b1 = *reinterpret_cast<const float4*>(b2s + col0 + cc + ch * 8 + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            uint4 o;
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 r = __half22float2(r2[t]);
              const __half2 h = __floats2half2_rn(__uint_as_float(v[ch * 8 + 2 * t]) + bb[2 * t] + r.x,
                                                  __uint_as_float(v[ch * 8 + 2 * t + 1]) + bb[2 * t + 1] + r.y);
              ow[t] = *reinterpret_cast<const uint32_t*>(&h);
            }
            sts128(stg + stg_off(lane, ch), o);
          }
          __syncwarp();
          __half* xbase = p.X + m_base * C + col0 + cc;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + cr;
            if (r < rows_valid) *(reinterpret_cast<uint4*>(xbase + static_cast<int64_t>(r) * C) + cch) = lds128(stg + stg_off(r, cch));
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&y_empty, 0);
        // the staging tiles live inside the G buffer: no warp may start the next tile's GELU stores before every warp is done here
        named_bar_sync(1, EPI_THREADS);
        ROW_T(6);
      }
    }
    ROW_T_FLUSH();
  }
  tc_fence_before();
#if SUNET_KERNEL_TIMING
  if (p.timing && threadIdx.x == 0) {
    unsigned long long gt2;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt2));
    p.timing[148 * 19 * 8 + blockIdx.x * 4 + 2] = static_cast<long long>(gt2);
  }
#endif
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other can still touch its smem / TMEM / barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

static long long* row_timing_buf(cudaStream_t stream) {
  static long long* buf = nullptr;
  if (!SUNET_KERNEL_TIMING || getenv("SUNET_MLP_TIMING") == nullptr) return nullptr;
  if (!buf && cudaMalloc(&buf, (148 * 19 * 8 + 148 * 4) * sizeof(long long)) != cudaSuccess) return nullptr;
  cudaMemsetAsync(buf, 0, (148 * 19 * 8 + 148 * 4) * sizeof(long long), stream);
  return buf;
}
static void row_timing_report(int C, unsigned grid, const long long* buf, cudaStream_t stream) {
  if (!buf) return;
  cudaStreamSynchronize(stream);
  static long long host[148 * 19 * 8 + 148 * 4];
  cudaMemcpy(host, buf, sizeof(host), cudaMemcpyDeviceToHost);
  static const char* en[8] = {"wait_h", "ld", "gelu", "wait_g_empty", "store", "wait_y", "out", "total"};
  static const char* in[8] = {"wait_h_empty", "wait_r1", "wait_g_full", "wait_r2", "wait_t/y", "issue", "-", "total"};   // fc1 issuer: 0 1 4 5 7, fc2 issuer: 2 3 4 5 7
  static const char* pn[8] = {"wait_r1_empty", "wait_r2_empty", "issue", "-", "-", "-", "-", "total"};
  {   // wall-clock span of the grid (globaltimer, ns): first CTA start -> last prologue end -> last CTA at the final cluster barrier
    long long s0 = 0, s1 = 0, p1 = 0, e0 = 0, e1 = 0;
    for (unsigned b = 0; b < grid; ++b) {
      const long long* t = host + 148 * 19 * 8 + b * 4;
      if (b == 0 || t[0] < s0) s0 = t[0];
      if (b == 0 || t[0] > s1) s1 = t[0];
      if (b == 0 || t[1] > p1) p1 = t[1];
      if (b == 0 || t[2] < e0) e0 = t[2];
      if (b == 0 || t[2] > e1) e1 = t[2];
    }
    fprintf(stderr, "mlp_row<%d> grid %u span (ns): last CTA start +%lld, last prologue end +%lld, first CTA done +%lld, last CTA done +%lld\n", C, grid,
            s1 - s0, p1 - s0, e0 - s0, e1 - s0);
  }
  for (int rank = 0; rank < 2; ++rank) {
    double e[8] = {0}, is[8] = {0}, is2[8] = {0}, pr[8] = {0};
    int n = 0;
    for (unsigned b = rank; b < grid; b += 2, ++n)
      for (int i = 0; i < 8; ++i) {
        for (int w = 2; w < 18; ++w) e[i] += static_cast<double>(host[(b * 19 + w) * 8 + i]) / 16;
        is[i] += static_cast<double>(host[(b * 19 + 1) * 8 + i]);
        is2[i] += static_cast<double>(host[(b * 19 + 18) * 8 + i]);
        pr[i] += static_cast<double>(host[(b * 19 + 0) * 8 + i]);
      }
    fprintf(stderr, "mlp_row<%d> rank %d (avg cycles over %d CTAs) epilogue warp:", C, rank, n);
    for (int i = 0; i < 8; ++i) fprintf(stderr, " %s %.0f", en[i], e[i] / n);
    fprintf(stderr, " | fc1 issuer:");
    for (int i = 0; i < 8; ++i) if (in[i][0] != '-' && i != 2 && i != 3) fprintf(stderr, " %s %.0f", in[i], is[i] / n);
    fprintf(stderr, " | fc2 issuer:");
    for (int i = 0; i < 8; ++i) if (in[i][0] != '-' && i != 0 && i != 1) fprintf(stderr, " %s %.0f", in[i], is2[i] / n);
    fprintf(stderr, " | producer:");
    for (int i = 0; i < 8; ++i) if (pn[i][0] != '-') fprintf(stderr, " %s %.0f", pn[i], pr[i] / n);
    fprintf(stderr, "\n");
  }
}

template <int C, bool PROJ>
int launch_t(const MlpRowPack& pk, const __half* T, const __half* R, __half* X, int64_t M, cudaStream_t stream) {
  using K = Cfg<C>;
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute((mlp_row_kernel<C, PROJ>), cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    once.done();
  }
  alignas(64) CUtensorMap tmT;
  SUNET_TRY(make_tmap_2d_f16(&tmT, T, C, M, C, TILE_M));
  Params p;
  p.hbias = PROJ ? pk.hbiasg : pk.hbias; p.b2 = pk.b2; p.bp = pk.bp; p.R = R; p.X = X; p.M = M;
  p.tiles = (M + TILE_M - 1) / TILE_M;
  const int64_t pair_tiles = (p.tiles + 1) / 2;
  const int64_t clusters = device_sms() / 2;
  const unsigned grid = static_cast<unsigned>(2 * (pair_tiles < clusters ? pair_tiles : clusters));
  p.w1 = PROJ ? pk.w1hg : pk.w1h; p.w2 = pk.w2; p.wp = pk.wp; p.wbytes = static_cast<uint32_t>(4u * C * C * sizeof(__half));
  p.timing = row_timing_buf(stream);
  SUNET_CUDA(launch_pdl((mlp_row_kernel<C, PROJ>), dim3(grid), dim3(THREADS), K::SMEM, stream, tmT, PROJ ? pk.tmW1g : pk.tmW1, pk.tmW2,
                        PROJ ? pk.tmWp : pk.tmW2, p));
  SUNET_CHECK_LAUNCH();
  row_timing_report(C, grid, p.timing, stream);
  return 0;
}

// pre-pack of the PROJ form: w1hg = fp16(0.5 * W1 * gamma), hconst[n] = (s_n, c_n) with s_n = sum_k w1hg[n,k] (the ROUNDED weights the
// MMA multiplies) and c_n = 0.5 * (b1 + W1 beta); one warp per hidden unit
__global__ void row_fold_ln_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, __half* __restrict__ w1hg, float2* __restrict__ hconst, int HID, int C) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= HID) return;
  float bb = 0.f, ss = 0.f;
  for (int k = lane; k < C; k += 32) {
    const float w = w1[static_cast<size_t>(n) * C + k];
    const __half h = __float2half_rn(0.5f * w * gamma[k]);
    w1hg[static_cast<size_t>(n) * C + k] = h;
    ss += __half2float(h);
    bb = fmaf(w, beta[k], bb);
  }
  for (int o = 16; o > 0; o >>= 1) {
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if (lane == 0) hconst[n] = make_float2(ss, 0.5f * (bb + (b1 ? b1[n] : 0.f)));
}

}  // namespace

bool mlp_row_supported(int C) { return C == 384; }

int mlp_row_prepare(MlpRowPack* p) {
  if (!mlp_row_supported(p->C)) return fail(SUNET_E_SHAPE, "mlp_row: C=%d not instantiated (384)", p->C);
  if (!p->w1h || !p->hbias || !p->w2) return fail(SUNET_E_ARG, "mlp_row: weights not set");
  const int C = p->C, HID = 4 * C;
  SUNET_TRY(make_tmap_2d_f16(&p->tmW1, p->w1h, C, HID, C, NC / 2));    // each CTA of a pair stages half of the rows of a block
  SUNET_TRY(make_tmap_2d_f16(&p->tmW2, p->w2, HID, C, HID, C / 4));
  return 0;
}

int mlp_row_set_proj(MlpRowPack* p, const __half* wp, const float* bp, const float* gamma, const float* beta, const float* w1, const float* b1,
                     __half* w1hg, float* hbiasg, cudaStream_t stream) {
  if (!mlp_row_supported(p->C)) return fail(SUNET_E_SHAPE, "mlp_row: C=%d not instantiated (384)", p->C);
  if (!wp || !gamma || !beta || !w1 || !w1hg || !hbiasg) return fail(SUNET_E_ARG, "mlp_row: proj pack buffers not set");
  const int C = p->C, HID = 4 * C;
  row_fold_ln_kernel<<<(HID + 7) / 8, 256, 0, stream>>>(w1, b1, gamma, beta, w1hg, reinterpret_cast<float2*>(hbiasg), HID, C);
  SUNET_CHECK_LAUNCH();
  p->wp = wp; p->bp = bp; p->w1hg = w1hg; p->hbiasg = hbiasg;
  SUNET_TRY(make_tmap_2d_f16(&p->tmW1g, w1hg, C, HID, C, NC / 2));
  SUNET_TRY(make_tmap_2d_f16(&p->tmWp, wp, C, C, C, C / 4));
  p->has_proj = 1;
  return 0;
}

int mlp_row_proj_launch(const MlpRowPack& p, const __half* attn_out, const __half* shortcut, __half* X, int64_t M, cudaStream_t stream) {
  if (!p.has_proj) return fail(SUNET_E_STATE, "mlp_row: proj weights not packed");
  if (M <= 0) return 0;
  if (M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "mlp_row: M too large for 32-bit TMA coordinates");
  if ((reinterpret_cast<uintptr_t>(attn_out) | reinterpret_cast<uintptr_t>(shortcut) | reinterpret_cast<uintptr_t>(X)) & 15)
    return fail(SUNET_E_ALIGN, "mlp_row: operands must be 16-byte aligned");
  if (attn_out == X) return fail(SUNET_E_ARG, "mlp_row: X must not alias attn_out (tiles of other CTAs are still being read)");
  if (p.C == 384) return launch_t<384, true>(p, attn_out, shortcut, X, M, stream);
  return fail(SUNET_E_SHAPE, "mlp_row: C=%d not instantiated", p.C);
}

int mlp_row_launch(const MlpRowPack& p, const __half* T, const __half* R, __half* X, int64_t M, cudaStream_t stream) {
  if (M <= 0) return 0;
  if (M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "mlp_row: M too large for 32-bit TMA coordinates");
  if ((reinterpret_cast<uintptr_t>(T) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(X)) & 15)
    return fail(SUNET_E_ALIGN, "mlp_row: operands must be 16-byte aligned");
  if (T == X) return fail(SUNET_E_ARG, "mlp_row: X must not alias T (tiles of other CTAs are still being read)");
  if (p.C == 384) return launch_t<384, false>(p, T, R, X, M, stream);
  return fail(SUNET_E_SHAPE, "mlp_row: C=%d not instantiated", p.C);
}

}  // namespace sunet
