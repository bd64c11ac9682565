// Fused LayerNorm -> fc1 -> GELU -> fc2 -> +residual for one Swin block (SUNet_detail.py:262 with Mlp :18-24).
//
// One persistent CTA per SM walks 128-token tiles.  Per tile the 4C-wide hidden activation is produced and consumed in
// 128-column chunks and never leaves the SM:
//
//   TMA  : x tile [128][C] fp16 (raw residual stream)  -> smem (SW128)         \  warp 0 (one lane)
//          fc1 / fc2 weight k-blocks                   -> two smem rings       /
//   MMA1 : H_j  = x * W1g_j^T          (tcgen05, fp32 in TMEM, N = 128)        \  warp 1 (one lane)
//   MMA2 : Y   += G_j * W2_j^T         (tcgen05, fp32 in TMEM, N = C)          /
//   GELU : G_j  = gelu(rstd * H_j - rstd * mu * s + b1f)  TMEM -> regs -> fp16 SW128 smem (A operand of MMA2)   \ warps 2..17
//   OUT  : y    = Y + b2 + x           TMEM -> regs -> global                                                   /
//
// LayerNorm fold (exact algebra):  fc1(LN(x))_n = rstd * (sum_k (W1[n,k] g[k]) x_k  -  mu * s_n) + (b1_n + sum_k W1[n,k] beta[k])
// with s_n = sum_k fp16(W1[n,k] g[k]) summed over the SAME rounded weights the MMA multiplies, so the only difference to
// "normalise, round to fp16, multiply" is that the activations are not rounded a second time.  mu / rstd per token are
// computed by the epilogue warps from the smem tile (fp32, shifted one-pass variance), which also keeps the residual in
// registers so the tile buffer can be refilled as soon as the last fc1 MMA of the tile has read it.
#include "mlp_fused.cuh"

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
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int EPI_THREADS = EPI_WARPS * 32;
// mlp_proj_fused with two MMA-issuing threads: warp 1 issues fc1, a 19th warp proj + fc2, so neither sits behind the other's
// barrier round trips.  Measured against a single issuing thread (round 1, tools/ab_mlp_split.sh): 3.30 -> 3.21 ms over the
// 32 launches of a forward; an idle 19th warp alone costs +2.6%, the split itself gains 6%.
constexpr int PTHREADS = THREADS + 32;
// A/B switch (tools/build_variant.py): 0 = CTA-wide barriers on the tile path, 1 = per-row-quadrant statistics barrier,
// 2 = + output staging inside the quadrant's own pieces of the hidden buffers (no CTA-wide barrier left on the tile path)
// Measured in one call on one box (tools/ab_variants.py, whole forward): 8.388 / 8.362 / 8.373 ms for 0 / 1 / 2 - boxes differ by ~1%.
#ifndef SUNET_MLP_QSYNC
#define SUNET_MLP_QSYNC 1
#endif
constexpr int FC2_WARP = 2 + EPI_WARPS;

template <int C>
struct Cfg {
  static constexpr int HID = 4 * C;
  static constexpr int NCH = HID / NC;                     // hidden chunks per tile
  static constexpr int KB1 = (C + 63) / 64;                // k-blocks of fc1 (K = C)
  static constexpr int KTAIL = (C % 64) ? (C % 64) / 16 : 4;  // k-steps in the last fc1 k-block
  // proj kernel: the fc1 bias rides the MMA when the last k-block has a spare 16-column k-step (C = 96): x1[:, C] = x1[:, C + 1] = 1
  // against two extra weight columns holding the bias as an fp16 (hi, lo) pair - the GELU pass loses its bias loads and adds
  // (40 of ~400 instructions per chunk in an issue-bound loop) for one more k-step of an idle tensor pipe.
#ifndef SUNET_MLP_BIASK
#define SUNET_MLP_BIASK 1
#endif
  static constexpr bool BIASK = SUNET_MLP_BIASK && (C % 64) != 0 && (C % 64) <= 48;
  static constexpr int W1PITCH = BIASK ? KB1 * 64 : C;       // row pitch (elements) of the packed proj-variant fc1 weights
  static constexpr int NXBUF = C <= 96 ? 2 : 1;            // token-tile buffers
  static constexpr int NYBUF = C <= 128 ? 2 : 1;           // fc2 accumulators in TMEM
  // C = 96 ring depths: 2 fc1 slots + 4 fc2 / proj slots (80 KB) instead of 3 + 3 (84 KB): the fc2 ring is the one whose refill
  // latency sits on the GELU -> fc2 -> next proj chain (measured, tools/ab_variants.py: 122.7 -> 119.5 us per launch)
#ifndef SUNET_MLP_R1_96
#define SUNET_MLP_R1_96 2
#endif
#ifndef SUNET_MLP_R2_96
#define SUNET_MLP_R2_96 4
#endif
  static constexpr int R1 = C <= 96 ? SUNET_MLP_R1_96 : 3;  // fc1 weight ring: [128 rows][64] k-blocks
  static constexpr int R2 = C <= 96 ? SUNET_MLP_R2_96 : 2;  // fc2 weight ring: [C rows][64] k-blocks
  static constexpr int R2BYTES = C * 128;
  static constexpr int QC = C / 4;                         // output columns per epilogue column-quarter
  static constexpr int QCH = C / 32;                       // 16-byte chunks (8 fp16) per quarter
  // shared memory map (offsets from the 1024-aligned base)
  static constexpr int OFF_X = 0;
  static constexpr int OFF_HS = OFF_X + NXBUF * KB1 * KBYTES;
  static constexpr int OFF_R1 = OFF_HS + 2 * 2 * KBYTES;
  static constexpr int OFF_R2 = OFF_R1 + R1 * KBYTES;
  static constexpr int OFF_HC = OFF_R2 + R2 * R2BYTES;     // float2 [HID]
  static constexpr int OFF_B2 = OFF_HC + HID * 8;          // float [C]
  static constexpr int OFF_ST = OFF_B2 + C * 4;            // float2 [2][4][128]
  static constexpr int SMEM = OFF_ST + 2 * 4 * 128 * 8 + 1024;
  static constexpr uint32_t TM_Y = 0;                      // Y[b] at column b * 128 (NYBUF == 2) or 0
  static constexpr uint32_t TM_H = 256;                    // H[b] at column 256 + 128 b
  static_assert(C % 32 == 0 && C <= 256, "C must be a multiple of 32, at most 256");
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

struct Params {
  const float2* hconst;
  const float* b2;
  __half* out;
  int64_t M;
  int64_t tiles;
  long long* timing;   // optional [grid][16 epilogue warps][8] phase cycle counters (SUNET_MLP_TIMING bring-up aid)
  long long* trace;    // optional event trace of CTA 0 (timing builds): [0] = count, then (event, clock) pairs
};

// Phase cycle counters of the epilogue warps: compiled in only with -DSUNET_KERNEL_TIMING=1 (they cost ~20 registers)
#ifndef SUNET_KERNEL_TIMING
#define SUNET_KERNEL_TIMING 0
#endif
#define TR_SLICE 400   // event-trace slots per traced thread (timing builds)
#if SUNET_KERNEL_TIMING
// event trace of CTA 0, lane 0 of the first and last epilogue warp and the issuing threads: (warp << 8 | event, clock) pairs
// (no atomics, no loads: every traced thread appends to its own slice - warp w owns events [400 w, 400 w + 400) - with plain stores,
// so that logging costs the issuing threads a few cycles, not a global round trip)
__device__ __forceinline__ void tr_log(long long* tr, unsigned& n, int ev, long long t) {
  if (n < TR_SLICE) {
    const unsigned i = (threadIdx.x >> 5) * TR_SLICE + n;
    tr[1 + 2 * i] = ev; tr[2 + 2 * i] = t;
    ++n;
  }
}
#define TR(ev) do { if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) tr_log(p.trace, trn, (static_cast<int>(threadIdx.x >> 5) << 8) | (ev), clock64()); } while (0)
#define MLP_T(i) do { if (p.timing) { const long long _t = clock64(); tacc[i] += _t - tq0; tq0 = _t; \
    if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) == 2 || (threadIdx.x >> 5) == 17)) tr_log(p.trace, trn, (static_cast<int>(threadIdx.x >> 5) << 8) | (i), _t); } } while (0)
#define MLP_T_DECL long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define MLP_T_START long long tq0 = p.timing ? clock64() : 0
#define MLP_T_RESTART do { if (p.timing) tq0 = clock64(); } while (0)
#else
#define MLP_T(i) do { } while (0)
#define TR(ev) do { } while (0)
#define MLP_T_DECL do { } while (0)
#define MLP_T_START do { } while (0)
#define MLP_T_RESTART do { } while (0)
#endif

template <int C>
__global__ void __launch_bounds__(THREADS, 1)
    mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const Params p) {
  using K = Cfg<C>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full[2], x_empty[2];
  __shared__ __align__(8) uint64_t r1_full[K::R1], r1_empty[K::R1], r2_full[K::R2], r2_empty[K::R2];
  __shared__ __align__(8) uint64_t h_full[2], gelu_done[2], hs_empty[2], y_full[2], y_empty[2];
  __shared__ uint32_t tmem_base_smem;

  // 1024-byte alignment for the SW128 tiles; pointer arithmetic on smem_raw keeps the shared address space visible to
  // the compiler (LDS / STS instead of generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if SUNET_KERNEL_TIMING
  unsigned trn = 0;   // events logged by this thread (event trace of CTA 0)
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1 + EPI_WARPS);
      mbar_init(&h_full[i], 1);
      mbar_init(&gelu_done[i], EPI_WARPS);
      mbar_init(&hs_empty[i], 1);
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], EPI_WARPS);
    }
    for (int i = 0; i < K::R1; ++i) { mbar_init(&r1_full[i], 1); mbar_init(&r1_empty[i], 1); }
    for (int i = 0; i < K::R2; ++i) { mbar_init(&r2_full[i], 1); mbar_init(&r2_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  {  // per-column constants of the two epilogues
    float2* hc = reinterpret_cast<float2*>(smem + K::OFF_HC);
    for (int i = threadIdx.x; i < K::HID; i += THREADS) hc[i] = __ldg(p.hconst + i);
    float* b2s = reinterpret_cast<float*>(smem + K::OFF_B2);
    for (int i = threadIdx.x; i < C; i += THREADS) b2s[i] = __ldg(p.b2 + i);
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
      uint32_t i1 = 0, i2 = 0, g = 0;
      int lt = 0;
      auto load_x = [&](int64_t tile, int ltile) {
        const int xb = K::NXBUF == 2 ? (ltile & 1) : 0;
        const uint32_t use = K::NXBUF == 2 ? (ltile >> 1) : ltile;
        mbar_wait(&x_empty[xb], (use & 1) ^ 1);
        mbar_arrive_expect_tx(&x_full[xb], K::KB1 * KBYTES);
        for (int kb = 0; kb < K::KB1; ++kb)
          tma_load_2d(smem + K::OFF_X + (xb * K::KB1 + kb) * KBYTES, &tmX, &x_full[xb], kb * 64, static_cast<int>(tile * TILE_M));
      };
      auto load_w2 = [&](uint32_t chunk) {  // chunk index within its tile
        for (int kb = 0; kb < 2; ++kb, ++i2) {
          const int s = i2 % K::R2;
          mbar_wait(&r2_empty[s], ((i2 / K::R2) & 1) ^ 1);
          mbar_arrive_expect_tx(&r2_full[s], K::R2BYTES);
          tma_load_2d(smem + K::OFF_R2 + s * K::R2BYTES, &tmW2, &r2_full[s], static_cast<int>(chunk) * NC + kb * 64, 0);
        }
      };
      if (K::NXBUF == 2 && static_cast<int64_t>(blockIdx.x) < p.tiles) load_x(blockIdx.x, 0);
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        if (K::NXBUF == 2) {
          if (tile + gridDim.x < p.tiles) load_x(tile + gridDim.x, lt + 1);   // prefetch one tile ahead
        } else {
          load_x(tile, lt);
        }
        for (int j = 0; j < K::NCH; ++j, ++g) {
          for (int kb = 0; kb < K::KB1; ++kb, ++i1) {
            const int s = i1 % K::R1;
            mbar_wait(&r1_empty[s], ((i1 / K::R1) & 1) ^ 1);
            mbar_arrive_expect_tx(&r1_full[s], KBYTES);
            tma_load_2d(smem + K::OFF_R1 + s * KBYTES, &tmW1, &r1_full[s], kb * 64, j * NC);
          }
          if (g > 0) load_w2(j == 0 ? K::NCH - 1 : j - 1);   // fc2 weights of the previous chunk (MMA order)
        }
      }
      if (g > 0) load_w2(K::NCH - 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_f16(TILE_M, NC);
      const uint32_t idesc2 = umma_idesc_f16(TILE_M, C);
      uint32_t i1 = 0, i2 = 0, g = 0;
      int lt = 0;
      bool pending = false;
      uint32_t pg = 0;    // global chunk index of the pending fc2
      int pj = 0, plt = 0;
      auto mma2 = [&]() {
        const uint32_t hb = pg & 1;
        const int yb = K::NYBUF == 2 ? (plt & 1) : 0;
        const uint32_t yuse = K::NYBUF == 2 ? (plt >> 1) : plt;
        mbar_wait(&gelu_done[hb], (pg >> 1) & 1);
        if (pj == 0) mbar_wait(&y_empty[yb], (yuse & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + K::TM_Y + (K::NYBUF == 2 ? yb * 128 : 0);
        for (int kb = 0; kb < 2; ++kb, ++i2) {
          const int s = i2 % K::R2;
          mbar_wait(&r2_full[s], (i2 / K::R2) & 1);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_HS + (hb * 2 + kb) * KBYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + K::OFF_R2 + s * K::R2BYTES));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2,
                        (pj > 0 || kb > 0 || k > 0) ? 1u : 0u);
          tc_commit(&r2_empty[s]);
        }
        tc_commit(&hs_empty[hb]);
        if (pj == K::NCH - 1) tc_commit(&y_full[yb]);
      };
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        const int xb = K::NXBUF == 2 ? (lt & 1) : 0;
        const uint32_t xuse = K::NXBUF == 2 ? (lt >> 1) : lt;
        mbar_wait(&x_full[xb], xuse & 1);
        tc_fence_after();
        for (int j = 0; j < K::NCH; ++j, ++g) {
          const uint32_t hb = g & 1;
          // H[hb] was drained by the GELU pass of chunk g-2: mma2(g-2), issued earlier, already waited for it
          const uint32_t d = tmem_base + K::TM_H + hb * 128;
          for (int kb = 0; kb < K::KB1; ++kb, ++i1) {
            const int s = i1 % K::R1;
            mbar_wait(&r1_full[s], (i1 / K::R1) & 1);
            tc_fence_after();
            const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_X + (xb * K::KB1 + kb) * KBYTES));
            const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + K::OFF_R1 + s * KBYTES));
            const int ksteps = kb == K::KB1 - 1 ? K::KTAIL : 4;
            for (int k = 0; k < ksteps; ++k)
              umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc1,
                          (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&r1_empty[s]);
          }
          tc_commit(&h_full[hb]);
          if (j == K::NCH - 1) tc_commit(&x_empty[xb]);   // every fc1 MMA of this tile has read the token tile
          if (pending) mma2();
          pending = true; pg = g; pj = j; plt = lt;
        }
      }
      if (pending) mma2();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quadrant this warp may touch
    const int quarter = e >> 2;        // column quarter
    const int row = q * 32 + lane;     // row of the tile
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float2* hc = reinterpret_cast<const float2*>(smem + K::OFF_HC);
    const float* b2s = reinterpret_cast<const float*>(smem + K::OFF_B2);
    float2* stats = reinterpret_cast<float2*>(smem + K::OFF_ST);
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint32_t g = 0;
    int lt = 0;
    MLP_T_DECL;
    for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
      const int xb = K::NXBUF == 2 ? (lt & 1) : 0;
      const uint32_t xuse = K::NXBUF == 2 ? (lt >> 1) : lt;
      // ---- row statistics + residual capture
      MLP_T_START;
      mbar_wait(&x_full[xb], xuse & 1);
      MLP_T(0);
      const uint32_t xs = smem_u32(smem + K::OFF_X + xb * K::KB1 * KBYTES);
      uint4 res[K::QCH];
      float s1 = 0.f, s2 = 0.f;
      {
        const uint4 first = lds128(xs + row * 128 + (sw << 4));   // chunk 0 of this row (columns 0..7)
        const float k0 = __half2float(__ushort_as_half(static_cast<unsigned short>(first.x & 0xffffu)));
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) {
          const int gi = quarter * K::QCH + i;
          const int kb = gi >> 3, ch = gi & 7;
          res[i] = lds128(xs + kb * KBYTES + row * 128 + ((static_cast<uint32_t>(ch) ^ sw) << 4));
          const __half2* h2 = reinterpret_cast<const __half2*>(&res[i]);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(h2[t]);
            const float d0 = f.x - k0, d1 = f.y - k0;
            s1 += d0 + d1;
            s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
          }
        }
        float2* st = stats + (lt & 1) * 4 * 128;
        st[quarter * 128 + row] = make_float2(s1, s2);
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[xb]);        // this warp no longer reads the token tile
        named_bar_sync(1, EPI_THREADS);
        s1 = 0.f; s2 = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 v = st[t * 128 + row];
          s1 += v.x; s2 += v.y;
        }
        const float ms = s1 * (1.0f / C);
        const float var = fmaxf(s2 * (1.0f / C) - ms * ms, 0.f);
        const float rstd = rsqrtf(var + 1e-5f);
        s1 = rstd;                  // a
        s2 = -(k0 + ms) * rstd;     // b
      }
      const float a = s1, b = s2;
      MLP_T(1);
      // ---- GELU passes
      for (int j = 0; j < K::NCH; ++j, ++g) {
        const uint32_t hb = g & 1, ph = (g >> 1) & 1;
        mbar_wait(&h_full[hb], ph);
        tc_fence_after();
        MLP_T(2);
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + K::TM_H + hb * 128 + quarter * 32, v);
        tmem_ld_wait();
        const float4* hc4 = reinterpret_cast<const float4*>(hc + j * NC + quarter * 32);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 c2 = hc4[i];   // (s, b1f) of two consecutive hidden columns (warp-uniform address: broadcast)
          const float h0 = fmaf(a, __uint_as_float(v[2 * i]), fmaf(b, c2.x, c2.y));
          const float h1 = fmaf(a, __uint_as_float(v[2 * i + 1]), fmaf(b, c2.z, c2.w));
          const __half2 x2 = __floats2half2_rn(h0, h1);
          w[i] = gelu_fast_h2(*reinterpret_cast<const uint32_t*>(&x2));
        }
        MLP_T(3);
        mbar_wait(&hs_empty[hb], ph ^ 1);
        MLP_T(4);   // fc2 of chunk g-2 has consumed this buffer
        const uint32_t hs = smem_u32(smem + K::OFF_HS + (hb * 2 + (quarter >> 1)) * KBYTES) + row * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sts128(hs + (((static_cast<uint32_t>((quarter & 1) * 4 + i)) ^ sw) << 4), make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gelu_done[hb]);
        MLP_T(5);
      }
      // ---- output: Y + b2 + residual -> global
      {
        const int yb = K::NYBUF == 2 ? (lt & 1) : 0;
        const uint32_t yuse = K::NYBUF == 2 ? (lt >> 1) : lt;
        mbar_wait(&y_full[yb], yuse & 1);
        tc_fence_after();
        MLP_T(6);
        const uint32_t ty = tmem_base + lane_off + K::TM_Y + (K::NYBUF == 2 ? yb * 128 : 0) + quarter * K::QC;
        uint32_t y[K::QC];
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) tmem_ld8(ty + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&y[i * 8]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[yb]);
        const int64_t m = tile * TILE_M + row;
        if (m < p.M) {
          __half* orow = p.out + m * C + quarter * K::QC;
#pragma unroll
          for (int i = 0; i < K::QCH; ++i) {
            const __half2* r2 = reinterpret_cast<const __half2*>(&res[i]);
            const float* bb = b2s + quarter * K::QC + i * 8;
            uint4 o;
            __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 r = __half22float2(r2[t]);
              o2[t] = __floats2half2_rn(__uint_as_float(y[i * 8 + 2 * t]) + bb[2 * t] + r.x,
                                        __uint_as_float(y[i * 8 + 2 * t + 1]) + bb[2 * t + 1] + r.y);
            }
            *reinterpret_cast<uint4*>(orow + i * 8) = o;
          }
        }
        MLP_T(7);
      }
    }
#if SUNET_KERNEL_TIMING
    if (p.timing && lane == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) p.timing[(static_cast<long long>(blockIdx.x) * EPI_WARPS + e) * 16 + i] = tacc[i];
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================ proj + shortcut + MLP
// The back half of a Swin block in one kernel (SUNet_detail.py:136 proj, :261 first residual, :262 norm2 + Mlp + second
// residual):   x1 = shortcut + attn_out Wp^T + bp ;   y = x1 + fc2(GELU(fc1(LN(x1))))
// Same dataflow as mlp_fused_kernel with one more MMA in front:
//   TMA  : attn_out tile [128][C] -> smem X ; Wp k-blocks ride the fc2 weight ring (same [C rows][64] shape)
//   MMA0 : P = attn_out * Wp^T  (fp32, in the TMEM columns of this tile's fc2 accumulator, which is idle until fc2 starts)
//   EPI0 : x1 = P + bp + shortcut (global, issued before the wait) -> fp16 -> written IN PLACE over the attn_out tile as
//          the A operand of fc1 and kept in registers as the residual; LayerNorm statistics of x1 (per-quarter shifted
//          sums merged with the parallel-variance formula, no cross-thread shift needed)
//   then fc1 / GELU / fc2 / output as in mlp_fused_kernel, with fc1 issued two hidden chunks ahead of fc2.
struct ProjParams {
  const float* hbias;     // 0.5 * (fc1.bias + fc1.weight beta) [4C]
  const float* b2;
  const float* bp;        // proj.bias [C]
  const __half* shortcut; // block input x [M][C] (may alias out: each thread reads and later writes the same elements)
  __half* out;
  int64_t M;
  int64_t tiles;
  long long* timing;   // see Params::timing
  long long* trace;    // see Params::trace
};

// 19 warps put 5 on one scheduler: 96 registers per thread is the cap of a 16K-register SM sub-partition (104 is unlaunchable)
template <int C>
__global__ void __launch_bounds__(PTHREADS, 1)
    mlp_proj_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWp,
                          const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const ProjParams p) {
  using K = Cfg<C>;
  constexpr int OFF_BP = K::OFF_ST + 2 * 4 * 128 * 8;   // float [C] after the statistics exchange
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full[2], x_empty[2], x1_ready[2], p_full[2];
  __shared__ __align__(8) uint64_t r1_full[K::R1], r1_empty[K::R1], r2_full[K::R2], r2_empty[K::R2];
  __shared__ __align__(8) uint64_t h_full[2], h_empty[2], gelu_done[2], hs_empty[2], y_full[2], y_empty[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#if SUNET_KERNEL_TIMING
  unsigned trn = 0;   // events logged by this thread (event trace of CTA 0)
#endif

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWp);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&x1_ready[i], EPI_WARPS);
      mbar_init(&p_full[i], 1);
      mbar_init(&h_full[i], 1);
      mbar_init(&h_empty[i], EPI_WARPS);
      mbar_init(&gelu_done[i], EPI_WARPS);
      mbar_init(&hs_empty[i], 1);
      mbar_init(&y_full[i], 1);
      mbar_init(&y_empty[i], EPI_WARPS);
    }
    for (int i = 0; i < K::R1; ++i) { mbar_init(&r1_full[i], 1); mbar_init(&r1_empty[i], 1); }
    for (int i = 0; i < K::R2; ++i) { mbar_init(&r2_full[i], 1); mbar_init(&r2_empty[i], 1); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  {
    float* hb = reinterpret_cast<float*>(smem + K::OFF_HC);
    if constexpr (!K::BIASK)   // (BIASK: the bias is in the weights, nothing reads this table)
      for (int i = threadIdx.x; i < K::HID; i += PTHREADS) hb[i] = __ldg(p.hbias + i);
    float* b2s = reinterpret_cast<float*>(smem + K::OFF_B2);
    float* bps = reinterpret_cast<float*>(smem + OFF_BP);
    for (int i = threadIdx.x; i < C; i += PTHREADS) { b2s[i] = __ldg(p.b2 + i); bps[i] = __ldg(p.bp + i); }
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
      // Three independent streams - token tiles, the fc1 weight ring, the proj / fc2 weight ring - each in the order its consumer
      // reads it, served by ONE thread that never blocks on any of them: it polls the next slot of every stream and issues whatever
      // is free.  (The in-order producer it replaces sat in the wait for a free fc2-ring slot - freed at the pace of the GELU
      // passes - with the next tile's fc1 weights and proj weights queued behind it: 1.5k clk per tile at the first fc1 accumulator
      // and 1.4k clk at the proj accumulator of every tile, tools/phase_timing.sh.)
      constexpr bool EARLY = K::NXBUF == 2 && K::NYBUF == 2;
      const int64_t first = blockIdx.x, stride = gridDim.x;
      const int64_t my_tiles = first < p.tiles ? (p.tiles - first + stride - 1) / stride : 0;
      // X stream: local tile xl
      int64_t xl = 0;
      // r1 stream: i1 counts k-blocks over (tile, chunk, kb)
      uint32_t i1 = 0;
      const uint64_t n1 = static_cast<uint64_t>(my_tiles) * K::NCH * K::KB1;
      // r2 stream: i2 counts k-blocks; (r2_lt, r2_k) = position inside the per-tile sequence; r2_pro = proj k-blocks of the first tile (EARLY)
      uint32_t i2 = 0;
      int64_t r2_lt = 0;
      int r2_k = 0, r2_pro = (EARLY && my_tiles > 0) ? K::KB1 : 0;
      // per tile, in the consumer's order: [Wp: KB1 k-blocks] then the two W2 k-blocks of every hidden chunk.  EARLY: the Wp of a
      // tile's sequence is the NEXT tile's (absent for the last tile), the first tile's own Wp is the prologue r2_pro.
      auto r2_has_wp = [&](int64_t ltile) -> bool { return EARLY ? (ltile + 1 < my_tiles) : true; };
      auto r2_len = [&](int64_t ltile) -> int { return (r2_has_wp(ltile) ? K::KB1 : 0) + 2 * K::NCH; };
      auto r2_item = [&](int64_t ltile, int k, const CUtensorMap*& tm, int& col) {
        if (r2_has_wp(ltile)) {
          if (k < K::KB1) { tm = &tmWp; col = k * 64; return; }
          k -= K::KB1;
        }
        tm = &tmW2; col = (k >> 1) * NC + (k & 1) * 64;
      };
      while (xl < my_tiles || i1 < n1 || r2_lt < my_tiles) {
        bool progress = false;
        if (xl < my_tiles) {
          const int xb = K::NXBUF == 2 ? static_cast<int>(xl & 1) : 0;
          const uint32_t use = static_cast<uint32_t>(K::NXBUF == 2 ? (xl >> 1) : xl);
          if (mbar_test(&x_empty[xb], (use & 1) ^ 1)) {
            mbar_arrive_expect_tx(&x_full[xb], K::KB1 * KBYTES);
            for (int kb = 0; kb < K::KB1; ++kb)
              tma_load_2d(smem + K::OFF_X + (xb * K::KB1 + kb) * KBYTES, &tmX, &x_full[xb], kb * 64, static_cast<int>((first + xl * stride) * TILE_M));
            ++xl;
            progress = true;
          }
        }
        if (i1 < n1) {
          const int s = i1 % K::R1;
          if (mbar_test(&r1_empty[s], ((i1 / K::R1) & 1) ^ 1)) {
            const int kb = static_cast<int>(i1 % K::KB1), j = static_cast<int>((i1 / K::KB1) % K::NCH);
            mbar_arrive_expect_tx(&r1_full[s], KBYTES);
            tma_load_2d(smem + K::OFF_R1 + s * KBYTES, &tmW1, &r1_full[s], kb * 64, j * NC);
            ++i1;
            progress = true;
          }
        }
        if (r2_lt < my_tiles) {
          const int s = i2 % K::R2;
          if (mbar_test(&r2_empty[s], ((i2 / K::R2) & 1) ^ 1)) {
            const CUtensorMap* tm = &tmWp;
            int col = 0;
            if (r2_pro > 0) { col = (K::KB1 - r2_pro) * 64; --r2_pro; }
            else {
              r2_item(r2_lt, r2_k, tm, col);
              if (++r2_k == r2_len(r2_lt)) { r2_k = 0; ++r2_lt; }
            }
            mbar_arrive_expect_tx(&r2_full[s], K::R2BYTES);
            tma_load_2d(smem + K::OFF_R2 + s * K::R2BYTES, tm, &r2_full[s], col, 0);
            ++i2;
            progress = true;
          }
        }
        if (!progress) __nanosleep(64);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ fc1 issuer (the r1 ring is consumed only here)
    // The whole warp walks the loop and waits on the barriers; one elected lane issues.  (Inside an `if (lane == 0)` every descriptor is
    // thread-divergent for the compiler: R2UR + an ELECT loop around each tcgen05.mma / commit, ~180 clk per instruction - see
    // gemm_tcgen05.cu.)
    {
      const uint32_t idesc1 = umma_idesc_f16(TILE_M, NC);
      uint32_t i1 = 0, g = 0;
      int lt = 0;
      uint32_t r1_ok = mbar_test(&r1_full[0], 0);
#if SUNET_KERNEL_TIMING
      long long it_acc[4] = {0, 0, 0, 0};
      const long long it_start = p.timing ? clock64() : 0;
#define ISS_T0 const long long _i0 = p.timing ? clock64() : 0
#define ISS_T1(k) do { if (p.timing) { const long long _i1 = clock64(); it_acc[k] += _i1 - _i0; if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0) tr_log(p.trace, trn, (static_cast<int>(threadIdx.x >> 5) << 8) | 0x40 | (k), _i1); } } while (0)
#else
#define ISS_T0 do { } while (0)
#define ISS_T1(k) do { } while (0)
#endif
      auto r1_acquire = [&]() -> int {
        const int s = i1 % K::R1;
        { ISS_T0; mbar_wait_hint(&r1_full[s], (i1 / K::R1) & 1, r1_ok); ISS_T1(2); }
        ++i1;
        r1_ok = mbar_test(&r1_full[i1 % K::R1], (i1 / K::R1) & 1);
        tc_fence_after();
        return s;
      };
      // One barrier round trip costs this thread 350-550 clk even on a completed phase (event trace, tools/decode_mlp_trace.py), and
      // the chain x1_ready -> h_empty -> KB1 ring slots -> issue sat on the tile-boundary critical path (2.2k clk from x1_ready to
      // the first fc1 issue against the 2.1k clk the epilogue spends in `out`).  Everything the first chunk of a tile needs EXCEPT
      // x1 - its accumulator and its ring slots - is therefore acquired while the epilogue is still building x1.
      // (Only the FIRST chunk of a tile is acquired whole: for the others a ring slot is handed back as soon as its k-block has been
      // issued, so that the refill of slot kb runs under the MMAs of kb + 1 - acquiring all of them first cost C = 192 9%.)
      constexpr bool PRE = K::KB1 <= K::R1;   // all k-blocks of a chunk fit in the ring at once
      auto fc1_acquire = [&](uint32_t gg, bool whole, int (&slots)[K::KB1]) {
        const uint32_t hb = gg & 1, use = gg >> 1;
        if (use > 0) { ISS_T0; mbar_wait(&h_empty[hb], (use - 1) & 1); ISS_T1(1); }   // the epilogue has loaded the previous contents of this accumulator
        if (PRE && whole) {
#pragma unroll
          for (int kb = 0; kb < K::KB1; ++kb) slots[kb] = r1_acquire();
        }
      };
      auto fc1_issue = [&](int xb, uint32_t gg, int j, bool last, bool whole, int (&slots)[K::KB1]) {   // H[gg & 1] = x1 * W1h_j^T   (gg: global chunk index)
        const uint32_t hb = gg & 1;
        tc_fence_after();
        const uint32_t d = tmem_base + K::TM_H + hb * 128;
#pragma unroll
        for (int kb = 0; kb < K::KB1; ++kb) {
          const int s = (PRE && whole) ? slots[kb] : r1_acquire();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_X + (xb * K::KB1 + kb) * KBYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + K::OFF_R1 + s * KBYTES));
          constexpr int KS_LAST = K::KTAIL + (K::BIASK ? 1 : 0);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (kb < K::KB1 - 1 || k < KS_LAST)
                umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc1, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&r1_empty[s]);
            if (kb == K::KB1 - 1) {
              tc_commit(&h_full[hb]);
              if (last) tc_commit(&x_empty[xb]);   // every fc1 MMA of this tile has read x1: the buffer may be refilled
            }
          }
          __syncwarp();
        }
        TR(0x50 + j);
#if SUNET_KERNEL_TIMING && defined(SUNET_MLP_DIAG_HFULL)
        if (j == 0 && p.trace && blockIdx.x == 0) { mbar_wait(&h_full[hb], (gg >> 1) & 1); TR(0x6f); }   // diagnosis: when does this chunk's accumulator complete?
#endif
        (void)j;
      };
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        const int xb = K::NXBUF == 2 ? (lt & 1) : 0;
        const uint32_t xuse = K::NXBUF == 2 ? (lt >> 1) : lt;
        int slots[K::KB1];
        fc1_acquire(g, true, slots);
        { ISS_T0; mbar_wait(&x1_ready[xb], xuse & 1); ISS_T1(0); }
        fc1_issue(xb, g, 0, K::NCH == 1, true, slots);
        for (int j = 1; j < K::NCH; ++j) {   // runs ahead as far as the two H accumulators allow
          fc1_acquire(g + j, false, slots);
          fc1_issue(xb, g + j, j, j == K::NCH - 1, false, slots);
        }
        g += K::NCH;
      }
#if SUNET_KERNEL_TIMING
      if (p.timing) {
        it_acc[3] = clock64() - it_start;
        for (int i = 0; i < 4; ++i) p.timing[148 * EPI_WARPS * 16 + static_cast<long long>(blockIdx.x) * 16 + i] = it_acc[i];
      }
#endif
    }
  } else if (warp == FC2_WARP) {
    // ------------------------------------------------------------------ proj + fc2 issuer (the r2 ring is consumed only here, in the
    // producer's order: [Wp] then per chunk the two W2 k-blocks, the next tile's Wp before the last chunk when EARLY)
    {   // (whole warp, one elected lane issues: see the fc1 issuer)
      const uint32_t idesc2 = umma_idesc_f16(TILE_M, C);
      uint32_t i2 = 0, g = 0;
      int lt = 0;
      uint32_t r2_ok = mbar_test(&r2_full[0], 0);
#if SUNET_KERNEL_TIMING
      long long it_acc[5] = {0, 0, 0, 0, 0};
      const long long it_start = p.timing ? clock64() : 0;
#endif
      auto r2_acquire = [&]() -> int {
        const int s = i2 % K::R2;
        { ISS_T0; mbar_wait_hint(&r2_full[s], (i2 / K::R2) & 1, r2_ok); ISS_T1(1); }
        ++i2;
        r2_ok = mbar_test(&r2_full[i2 % K::R2], (i2 / K::R2) & 1);
        tc_fence_after();
        return s;
      };
      auto fc2 = [&](int yb, uint32_t gg, int j) {   // Y[yb] (+)= G_j * W2_j^T
        const uint32_t hb = gg & 1;
        { ISS_T0; mbar_wait(&gelu_done[hb], (gg >> 1) & 1); ISS_T1(0); }
        tc_fence_after();
        const uint32_t d = tmem_base + K::TM_Y + (K::NYBUF == 2 ? yb * 128 : 0);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const int s = r2_acquire();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_HS + (hb * 2 + kb) * KBYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + K::OFF_R2 + s * K::R2BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2,
                          (j > 0 || kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&r2_empty[s]);
            if (kb == 1) {
              tc_commit(&hs_empty[hb]);
              if (j == K::NCH - 1) tc_commit(&y_full[yb]);
            }
          }
          __syncwarp();
        }
        TR(0x70 + j);
      };
      auto mma0 = [&](int lt2) {   // P = attn_out * Wp^T into the (idle) fc2 accumulator of local tile lt2
        const int xb = K::NXBUF == 2 ? (lt2 & 1) : 0;
        const uint32_t xuse = K::NXBUF == 2 ? (lt2 >> 1) : lt2;
        const int yb = K::NYBUF == 2 ? (lt2 & 1) : 0;
        const uint32_t yuse = K::NYBUF == 2 ? (lt2 >> 1) : lt2;
        TR(0x5f);
        { ISS_T0; mbar_wait(&x_full[xb], xuse & 1); ISS_T1(2); }
        { ISS_T0; mbar_wait(&y_empty[yb], (yuse & 1) ^ 1); ISS_T1(3); }
        tc_fence_after();
        const uint32_t d = tmem_base + K::TM_Y + (K::NYBUF == 2 ? yb * 128 : 0);
#pragma unroll
        for (int kb = 0; kb < K::KB1; ++kb) {
          const int s = r2_acquire();
          const uint64_t adesc = umma_desc_sw128(smem_u32(smem + K::OFF_X + (xb * K::KB1 + kb) * KBYTES));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(smem + K::OFF_R2 + s * K::R2BYTES));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (kb < K::KB1 - 1 || k < K::KTAIL)
                umma_f16_ss(d, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc2, (kb > 0 || k > 0) ? 1u : 0u);
            tc_commit(&r2_empty[s]);
            if (kb == K::KB1 - 1) tc_commit(&p_full[yb]);
          }
          __syncwarp();
        }
        TR(0x60);
      };
      constexpr bool EARLY = K::NXBUF == 2 && K::NYBUF == 2;
      if (EARLY && static_cast<int64_t>(blockIdx.x) < p.tiles) mma0(0);
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        const int yb = K::NYBUF == 2 ? (lt & 1) : 0;
        // EARLY: the NEXT tile's proj goes first.  Its inputs (attn_out tile, Wp, the other fc2 accumulator - free once `out` of the
        // previous tile has loaded it) are all there long before this tile's first GELU pass ends, so P is waiting when the epilogue
        // gets to it (issued before the last fc2 of this tile it arrived 0.5-1.1k clk late, behind 1.7k clk of barrier round trips).
        if (EARLY ? (tile + gridDim.x < p.tiles) : true) mma0(EARLY ? lt + 1 : lt);
        for (int j = 0; j < K::NCH; ++j) fc2(yb, g + j, j);
        g += K::NCH;
      }
#if SUNET_KERNEL_TIMING
      if (p.timing) {
        it_acc[4] = clock64() - it_start;
        for (int i = 0; i < 5; ++i) p.timing[148 * EPI_WARPS * 16 + static_cast<long long>(blockIdx.x) * 16 + 8 + i] = it_acc[i];
      }
#endif
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue warps
    const int e = warp - 2;
    const int q = warp & 3;
    const int quarter = e >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float* hbv = reinterpret_cast<const float*>(smem + K::OFF_HC);
    const float* b2s = reinterpret_cast<const float*>(smem + K::OFF_B2);
    const float* bps = reinterpret_cast<const float*>(smem + OFF_BP);
    float2* stats = reinterpret_cast<float2*>(smem + K::OFF_ST);
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint32_t g = 0;
    int lt = 0;
    // the shortcut slice of this thread's row is fetched one tile ahead (global latency ~1-2 us against a ~5 us tile)
    constexpr bool EARLY = K::NXBUF == 2 && K::NYBUF == 2;
    constexpr bool PREFETCH = C <= 96;   // wider rows do not have the registers for it (96-register cap at 576 threads)
    uint4 nxt[K::QCH];
    auto fetch_shortcut = [&](int64_t tile) {
      const int64_t mm = tile * TILE_M + row;
      const bool ok = tile < p.tiles && mm < p.M;
      const __half* srow = p.shortcut + (ok ? mm : 0) * C + quarter * K::QC;
#pragma unroll
      for (int i = 0; i < K::QCH; ++i) nxt[i] = ok ? __ldg(reinterpret_cast<const uint4*>(srow + i * 8)) : make_uint4(0u, 0u, 0u, 0u);
    };
    if (PREFETCH) fetch_shortcut(blockIdx.x);
    // Early test results for the three waits at a tile boundary (p_full, y_full, the first h_full): a wait on an already completed
    // phase still costs 350-450 clk - 1.4k clk right after the global stores of `out` (event trace) - a test issued a phase earlier
    // costs nothing on the critical path.
    uint32_t h_ok = 0, p_ok = 0, y_ok = 0;
    // GELU store addresses of this thread inside a hidden buffer: computed once (left to itself the compiler rematerialises the
    // swizzle arithmetic in every chunk - the kernel is issue-bound, ~45 integer instructions per chunk)
    uint32_t hs_row = smem_u32(smem + K::OFF_HS + (quarter >> 1) * KBYTES) + row * 128;
    uint32_t hs_x0 = ((static_cast<uint32_t>((quarter & 1) * 4 + 0)) ^ sw) << 4, hs_x1 = ((static_cast<uint32_t>((quarter & 1) * 4 + 1)) ^ sw) << 4;
    uint32_t hs_x2 = ((static_cast<uint32_t>((quarter & 1) * 4 + 2)) ^ sw) << 4, hs_x3 = ((static_cast<uint32_t>((quarter & 1) * 4 + 3)) ^ sw) << 4;
    if constexpr (C <= 96) asm volatile("" : "+r"(hs_row), "+r"(hs_x0), "+r"(hs_x1), "+r"(hs_x2), "+r"(hs_x3));
    MLP_T_DECL;
    MLP_T_START;
    // EARLY (two token-tile buffers and two fc2 accumulators, i.e. C <= 96): x1 of the NEXT tile is built before the output of the
    // current one, so the last fc2 of this tile, the proj of the next and its first fc1 chunks run on the tensor pipe under epilogue
    // work instead of being waited for (those three waits were ~20% of the epilogue warps' time).
    auto epi0 = [&](int64_t tile, int lt, uint4 (&res)[K::QCH]) {
      const int xb = K::NXBUF == 2 ? (lt & 1) : 0;
      const int yb = K::NYBUF == 2 ? (lt & 1) : 0;
      const uint32_t yuse = K::NYBUF == 2 ? (lt >> 1) : lt;
      (void)xb; (void)yb; (void)yuse;
      // ---- EPI0: x1 = P + bp + shortcut
      MLP_T_RESTART;
      if (!PREFETCH) fetch_shortcut(tile);
#pragma unroll
      for (int i = 0; i < K::QCH; ++i) res[i] = nxt[i];
      mbar_wait_hint(&p_full[yb], yuse & 1, p_ok);
      p_ok = 0;
      tc_fence_after();
      MLP_T(0);
      {
        const uint32_t tp = tmem_base + lane_off + K::TM_Y + (K::NYBUF == 2 ? yb * 128 : 0) + quarter * K::QC;
        uint32_t pv[K::QC];
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) tmem_ld8(tp + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&pv[i * 8]));
        tmem_ld_wait();
        float s1 = 0.f, s2 = 0.f, k0 = 0.f;
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) {
          const __half2* r2 = reinterpret_cast<const __half2*>(&res[i]);
          const float* bb = bps + quarter * K::QC + i * 8;
          uint4 o;
          __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 r = __half22float2(r2[t]);
            o2[t] = __floats2half2_rn(__uint_as_float(pv[i * 8 + 2 * t]) + bb[2 * t] + r.x, __uint_as_float(pv[i * 8 + 2 * t + 1]) + bb[2 * t + 1] + r.y);
            const float2 f = __half22float2(o2[t]);   // statistics on the rounded stream values, as a separate LayerNorm pass would see them
            if (i == 0 && t == 0) k0 = f.x;
            const float d0 = f.x - k0, d1 = f.y - k0;
            s1 += d0 + d1;
            s2 = fmaf(d0, d0, fmaf(d1, d1, s2));
          }
          res[i] = o;   // x1 (fp16 stream value): residual of the second add
        }
        // per-quarter (mean, M2) around the quarter's own shift, merged over the 4 quarters (parallel variance)
        const float mq = k0 + s1 * (1.0f / K::QC);
        const float m2q = s2 - s1 * s1 * (1.0f / K::QC);
        float2* st = stats + (lt & 1) * 4 * 128;
        st[quarter * 128 + row] = make_float2(mq, m2q);
        // only the 4 column-quarter warps of a row quadrant exchange statistics (a CTA-wide barrier made all 16 wait for the slowest).
        // Without EARLY this barrier is also the one that separates the staging reads of the previous tile's `output` from this
        // tile's GELU stores into the same buffers: CTA-wide there.
        if constexpr (EARLY && SUNET_MLP_QSYNC >= 1) named_bar_sync(2 + q, 128);
        else named_bar_sync(1, EPI_THREADS);
        float mean = 0.f, m2 = 0.f;
        float mqs[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 v = st[t * 128 + row];
          mqs[t] = v.x; mean += v.x; m2 += v.y;
        }
        mean *= 0.25f;
#pragma unroll
        for (int t = 0; t < 4; ++t) m2 = fmaf(static_cast<float>(K::QC) * (mqs[t] - mean), mqs[t] - mean, m2);
        const float rstd = rsqrtf(fmaxf(m2 * (1.0f / C), 0.f) + 1e-5f);
        const float nb = -mean * rstd;
        // norm2 without its affine part (folded into fc1 at pre-pack), written over the attn_out tile as the fc1 operand
        const uint32_t xs = smem_u32(smem + K::OFF_X + xb * K::KB1 * KBYTES);
        if constexpr (EARLY) {   // `output` of the PREVIOUS local tile follows this epi0: look its accumulator up now
          if (lt > 0) y_ok = mbar_test(&y_full[(lt - 1) & 1], ((lt - 1) >> 1) & 1);
        }
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) {
          const __half2* r2 = reinterpret_cast<const __half2*>(&res[i]);
          uint4 o;
          __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 f = __half22float2(r2[t]);
            o2[t] = __floats2half2_rn(fmaf(f.x, rstd, nb), fmaf(f.y, rstd, nb));
          }
          const int gi = quarter * K::QCH + i;
          sts128(xs + (gi >> 3) * KBYTES + row * 128 + ((static_cast<uint32_t>(gi & 7) ^ sw) << 4), o);
        }
        if constexpr (K::BIASK) {   // columns C, C + 1 = 1.0 (fp16 0x3C00), C + 2 .. C + 7 = 0: the bias k-step (its other 8 columns are the TMA zero fill)
          if (quarter == 3) {
            constexpr int gb = C / 8;
            sts128(xs + (gb >> 3) * KBYTES + row * 128 + ((static_cast<uint32_t>(gb & 7) ^ sw) << 4), make_uint4(0x3C003C00u, 0u, 0u, 0u));
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&x1_ready[xb]);
      }
      MLP_T(1);
    };
    auto gelu_passes = [&](int64_t tile, int lt) {
      const int xb = K::NXBUF == 2 ? (lt & 1) : 0;
      const int yb = K::NYBUF == 2 ? (lt & 1) : 0;
      const uint32_t yuse = K::NYBUF == 2 ? (lt >> 1) : lt;
      (void)xb; (void)yb; (void)yuse;
      // ---- GELU passes
      // (EARLY: the next tile's shortcut is requested here, once per tile and outside the chunk loop - inside it the compiler
      // predicates the ~90 address / load instructions instead of branching, and every chunk pays their issue slots)
      if (PREFETCH && EARLY) fetch_shortcut(tile + gridDim.x);
      for (int j = 0; j < K::NCH; ++j, ++g) {
        const uint32_t hb = g & 1, ph = (g >> 1) & 1;
        mbar_wait_hint(&h_full[hb], ph, h_ok);
        tc_fence_after();
#if SUNET_KERNEL_TIMING
        if (p.timing) { const long long _t = clock64(); tacc[2] += _t - tq0; tacc[8 + (j < 3 ? j : 3)] += _t - tq0; tq0 = _t;
          if (p.trace && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) == 2 || (threadIdx.x >> 5) == 17)) tr_log(p.trace, trn, (static_cast<int>(threadIdx.x >> 5) << 8) | 2, _t); }
#endif
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_off + K::TM_H + hb * 128 + quarter * 32, v);
        const uint32_t hs_ok = mbar_test(&hs_empty[hb], ph ^ 1);   // looked up under the TMEM load / GELU math
        if (PREFETCH && !EARLY && j == K::NCH - 2) fetch_shortcut(tile + gridDim.x);   // next tile's shortcut, well ahead of its use
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_empty[hb]);   // the accumulator may be overwritten by the fc1 of chunk g + 2
        MLP_T(4);   // (timing builds: slot 4 = TMEM load, slot 3 = GELU math; the hs_empty wait is folded into slot 5)
        // u = 0.5 * fc1(LN(x1)) = D + hbias (fc1 weights and bias are pre-scaled by 0.5); GELU(2u) = u + u * tanh(u * P(u^2)), fp32 on
        // the FMA pipe (7 FMA-pipe clocks per element against 8 MUFU clocks: HFMA2 issues at half rate, so packed-half math is no cheaper)
        const float4* hb4 = reinterpret_cast<const float4*>(hbv + j * NC + quarter * 32);
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // (BIASK: the bias is already in the accumulator - and `x + 0.f` is not a no-op the compiler may drop (signed zeros): it cost
          // one FADD per element, 4.6% of the kernel's instructions, until the add itself became conditional)
          float a0 = __uint_as_float(v[4 * i + 0]), a1 = __uint_as_float(v[4 * i + 1]), a2 = __uint_as_float(v[4 * i + 2]), a3 = __uint_as_float(v[4 * i + 3]);
          if constexpr (!K::BIASK) {
            const float4 bq = hb4[i];   // warp-uniform address: broadcast
            a0 += bq.x; a1 += bq.y; a2 += bq.z; a3 += bq.w;
          }
          const float g0 = gelu_half_arg(a0), g1 = gelu_half_arg(a1);
          const float g2 = gelu_half_arg(a2), g3 = gelu_half_arg(a3);
          const __half2 p0 = __floats2half2_rn(g0, g1), p1 = __floats2half2_rn(g2, g3);
          w[2 * i] = *reinterpret_cast<const uint32_t*>(&p0);
          w[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&p1);
        }
        MLP_T(3);
#if SUNET_KERNEL_TIMING
        const long long _ths0 = p.timing ? clock64() : 0;
#endif
        mbar_wait_hint(&hs_empty[hb], ph ^ 1, hs_ok);
#if SUNET_KERNEL_TIMING
        if (p.timing) tacc[12 + (j < 3 ? j : 3)] += clock64() - _ths0;
#endif
        h_ok = j + 1 < K::NCH ? mbar_test(&h_full[hb ^ 1], ((g + 1) >> 1) & 1) : 0u;   // next chunk's accumulator, looked up under the stores
        if constexpr (EARLY) {   // last chunk: the next local tile's proj accumulator (epi0 of that tile follows)
          if (j == K::NCH - 1 && tile + gridDim.x < p.tiles) p_ok = mbar_test(&p_full[(lt + 1) & 1], ((lt + 1) >> 1) & 1);
        }
        if constexpr (PREFETCH) {   // C = 96: the precomputed addresses fit the register budget (at C = 192 they spill)
          const uint32_t hs = hs_row + hb * (2 * KBYTES);
          sts128(hs + hs_x0, make_uint4(w[0], w[1], w[2], w[3]));
          sts128(hs + hs_x1, make_uint4(w[4], w[5], w[6], w[7]));
          sts128(hs + hs_x2, make_uint4(w[8], w[9], w[10], w[11]));
          sts128(hs + hs_x3, make_uint4(w[12], w[13], w[14], w[15]));
        } else {
          const uint32_t hs = smem_u32(smem + K::OFF_HS + (hb * 2 + (quarter >> 1)) * KBYTES) + row * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts128(hs + (((static_cast<uint32_t>((quarter & 1) * 4 + i)) ^ sw) << 4), make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gelu_done[hb]);
        MLP_T(5);
      }
    };
    auto output = [&](int64_t tile, int lt, const uint4 (&res)[K::QCH]) {
      const int xb = K::NXBUF == 2 ? (lt & 1) : 0;
      const int yb = K::NYBUF == 2 ? (lt & 1) : 0;
      const uint32_t yuse = K::NYBUF == 2 ? (lt >> 1) : lt;
      (void)xb; (void)yb; (void)yuse;
      // ---- output: Y + b2 + x1 -> global
      {
        mbar_wait_hint(&y_full[yb], yuse & 1, y_ok);
        y_ok = 0;
        tc_fence_after();
        MLP_T(6);
        const uint32_t ty = tmem_base + lane_off + K::TM_Y + (K::NYBUF == 2 ? yb * 128 : 0) + quarter * K::QC;
        uint32_t y[K::QC];
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) tmem_ld8(ty + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&y[i * 8]));
        if constexpr (EARLY) {   // the first fc1 accumulator of the next local tile (its x1 was published before this `output`)
          if (tile + gridDim.x < p.tiles) h_ok = mbar_test(&h_full[g & 1], (g >> 1) & 1);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&y_empty[yb]);
        // Stage the 32 x C fp16 block of this row quadrant through the (now idle) GELU buffers so that global stores cover whole
        // rows: a thread-per-row store touches 32 different 128-byte lines per instruction and is LSU-bound (measured).
        // The staging rows of quadrant q live ONLY in bytes that the GELU stores of quadrant q's own rows write (rows 32 q .. 32 q + 31
        // of each of the four [128][64] hidden k-block buffers, 4 KB apiece), so the hazard between these staging accesses and the next
        // tile's GELU stores is confined to the 4 warps of the quadrant: no CTA-wide barrier on the tile path.
        constexpr int PITCH = C * 2 + 16;   // bytes; the 16-byte pad makes 8 consecutive rows hit 8 distinct bank groups
        // (SUNET_MLP_QSYNC = 2, EARLY configs only; measured no better than the CTA-wide barrier it removes, so off by default)
        constexpr bool QSTG = EARLY && SUNET_MLP_QSYNC >= 2;
        constexpr int RP = QSTG ? 16 : 32;   // staging rows per piece
        static_assert(!QSTG || 16 * PITCH <= 4096, "output staging must fit the quadrant's own pieces of the hidden buffers");
        const uint32_t stg = smem_u32(smem + K::OFF_HS) + static_cast<uint32_t>(q) * (QSTG ? 4096 : 32 * PITCH);
        auto stg_row = [&](int r) -> uint32_t { return stg + static_cast<uint32_t>(r / RP) * KBYTES + static_cast<uint32_t>(r % RP) * PITCH; };
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) {
          const __half2* r2 = reinterpret_cast<const __half2*>(&res[i]);
          const float* bb = b2s + quarter * K::QC + i * 8;
          uint4 o;
          __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 r = __half22float2(r2[t]);
            o2[t] = __floats2half2_rn(__uint_as_float(y[i * 8 + 2 * t]) + bb[2 * t] + r.x,
                                      __uint_as_float(y[i * 8 + 2 * t + 1]) + bb[2 * t + 1] + r.y);
          }
          sts128(stg_row(lane) + (quarter * K::QCH + i) * 16, o);
        }
        named_bar_sync(2 + q, 128);   // the 4 column-quarter warps of this row quadrant
        {
          constexpr int CPR = C / 8;   // 16-byte chunks per row
          const int64_t m0 = tile * TILE_M + q * 32 + quarter * 8;   // this warp writes 8 of the quadrant's 32 rows
#pragma unroll
          for (int kk = 0; kk < (8 * CPR + 31) / 32; ++kk) {
            const int idx = kk * 32 + lane;
            const int r = idx / CPR, ch = idx - r * CPR;
            if (idx < 8 * CPR && m0 + r < p.M)
              *reinterpret_cast<uint4*>(p.out + (m0 + r) * C + ch * 8) = lds128(stg_row(quarter * 8 + r) + ch * 16);
          }
        }
        if constexpr (QSTG) named_bar_sync(2 + q, 128);   // the staging reads above against the quadrant's next GELU stores into the same bytes
        else if constexpr (EARLY) named_bar_sync(1, EPI_THREADS);
        MLP_T(7);
      }
        };
    if constexpr (EARLY) {
      uint4 res_cur[K::QCH], res_nxt[K::QCH];
      if (static_cast<int64_t>(blockIdx.x) < p.tiles) epi0(blockIdx.x, 0, res_cur);
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        gelu_passes(tile, lt);
        const bool more = tile + gridDim.x < p.tiles;
        if (more) epi0(tile + gridDim.x, lt + 1, res_nxt);
        output(tile, lt, res_cur);
#pragma unroll
        for (int i = 0; i < K::QCH; ++i) res_cur[i] = res_nxt[i];
      }
    } else {
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++lt) {
        uint4 res[K::QCH];
        epi0(tile, lt, res);
        gelu_passes(tile, lt);
        output(tile, lt, res);
      }
    }
#if SUNET_KERNEL_TIMING
    if (p.timing && lane == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) p.timing[(static_cast<long long>(blockIdx.x) * EPI_WARPS + e) * 16 + i] = tacc[i];
    }
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- pre-pack: W1g = fp16(W1 * gamma), s = rowsum(W1g as rounded), b1f = b1 + W1 beta; one warp per hidden unit
__global__ void mlp_fold_ln_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, __half* __restrict__ w1g, float2* __restrict__ hconst, int HID, int C) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= HID) return;
  float s = 0.f, bb = 0.f;
  for (int k = lane; k < C; k += 32) {
    const float w = w1[static_cast<size_t>(n) * C + k];
    const __half h = __float2half_rn(w * gamma[k]);
    w1g[static_cast<size_t>(n) * C + k] = h;
    s += __half2float(h);
    bb = fmaf(w, beta[k], bb);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
  }
  if (lane == 0) hconst[n] = make_float2(s, bb + (b1 ? b1[n] : 0.f));
}

// proj variant: W1h = fp16(0.5 * W1 * gamma), hbias = 0.5 * (b1 + W1 beta)  (LayerNorm runs in the kernel, GELU takes u = x / 2).
// pitch > C (BIASK): rows are padded to whole 64-column k-blocks, columns C / C + 1 hold hbias as an fp16 (hi, lo) pair, the rest 0.
__global__ void mlp_fold_half_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, __half* __restrict__ w1h, float* __restrict__ hbias, int HID, int C, int pitch) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= HID) return;
  float bb = 0.f;
  for (int k = lane; k < C; k += 32) {
    const float w = w1[static_cast<size_t>(n) * C + k];
    w1h[static_cast<size_t>(n) * pitch + k] = __float2half_rn(0.5f * w * gamma[k]);
    bb = fmaf(w, beta[k], bb);
  }
  for (int o = 16; o > 0; o >>= 1) bb += __shfl_xor_sync(0xffffffffu, bb, o);
  const float hb = 0.5f * (bb + (b1 ? b1[n] : 0.f));
  if (lane == 0) hbias[n] = hb;
  if (pitch > C) {
    const __half hi = __float2half_rn(hb);
    const __half lo = __float2half_rn(hb - __half2float(hi));
    for (int k = C + lane; k < pitch; k += 32)
      w1h[static_cast<size_t>(n) * pitch + k] = k == C ? hi : (k == C + 1 ? lo : __float2half_rn(0.f));
  }
}

__global__ void cast_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}

static long long* mlp_trace_buf(cudaStream_t stream) {
  static long long* buf = nullptr;
  if (!SUNET_KERNEL_TIMING || getenv("SUNET_MLP_TRACE") == nullptr) return nullptr;
  if (!buf && cudaMalloc(&buf, (1 + 2 * 20 * TR_SLICE) * sizeof(long long)) != cudaSuccess) return nullptr;
  cudaMemsetAsync(buf, 0, (1 + 2 * 20 * TR_SLICE) * sizeof(long long), stream);
  return buf;
}
static void mlp_trace_report(int C, const long long* buf, cudaStream_t stream) {
  static int printed = 0;
  if (!buf) return;
  cudaStreamSynchronize(stream);
  if (printed >= 2 && C == 96) return;
  if (C == 96) ++printed;
  static long long host[1 + 2 * 20 * TR_SLICE];
  cudaMemcpy(host, buf, sizeof(host), cudaMemcpyDeviceToHost);
  const int n = 20 * TR_SLICE;
  long long t0 = 0;
  for (int i = 0; i < n; ++i) if (host[2 + 2 * i] != 0 && (t0 == 0 || host[2 + 2 * i] < t0)) t0 = host[2 + 2 * i];
  fprintf(stderr, "TRACE mlp_proj_fused<%d> (t, warp, event)\n", C);
  for (int i = 0; i < n; ++i)
    if (host[2 + 2 * i] != 0) fprintf(stderr, "TR %lld %lld 0x%02llx\n", host[2 + 2 * i] - t0, host[1 + 2 * i] >> 8, host[1 + 2 * i] & 0xff);
}
static long long* mlp_timing_buf(cudaStream_t stream) {
  static long long* buf = nullptr;
  if (!SUNET_KERNEL_TIMING || getenv("SUNET_MLP_TIMING") == nullptr) return nullptr;
  if (!buf && cudaMalloc(&buf, (148 * EPI_WARPS * 16 + 148 * 16) * sizeof(long long)) != cudaSuccess) return nullptr;
  cudaMemsetAsync(buf, 0, (148 * EPI_WARPS * 16 + 148 * 16) * sizeof(long long), stream);
  return buf;
}
static void mlp_timing_report(const char* what, int C, unsigned grid, const long long* buf, cudaStream_t stream) {
  if (!buf) return;
  cudaStreamSynchronize(stream);
  static long long host[148 * EPI_WARPS * 16 + 148 * 16];
  cudaMemcpy(host, buf, sizeof(host), cudaMemcpyDeviceToHost);
  static const char* names[16] = {"wait_x", "stats", "wait_h", "gelu", "wait_hs", "store", "wait_y", "out",
                                  "wait_h[j=0]", "wait_h[j=1]", "wait_h[j=2]", "wait_h[j>2]", "wait_hs[j=0]", "wait_hs[j=1]", "wait_hs[j=2]", "wait_hs[j>2]"};
  double acc[16] = {0};
  for (unsigned b = 0; b < grid; ++b)
    for (int w = 0; w < EPI_WARPS; ++w)
      for (int i = 0; i < 16; ++i) acc[i] += static_cast<double>(host[(b * EPI_WARPS + w) * 16 + i]);
  fprintf(stderr, "%s<%d> grid=%u cycles per epilogue warp:", what, C, grid);
  for (int i = 0; i < 16; ++i) fprintf(stderr, " %s %.0f", names[i], acc[i] / grid / EPI_WARPS);
  // per-CTA minimum / maximum over the 16 epilogue warps (averaged over CTAs): a wait that even the LAST warp to arrive pays is a
  // real bubble, a wait only the first warps pay is skew between warps that share an issue port
  fprintf(stderr, " | min/max over warps:");
  for (int i = 0; i < 8; ++i) {
    double mn = 0, mx = 0;
    for (unsigned b = 0; b < grid; ++b) {
      long long lo = host[(b * EPI_WARPS) * 16 + i], hi = lo;
      for (int w = 1; w < EPI_WARPS; ++w) { const long long v = host[(b * EPI_WARPS + w) * 16 + i]; lo = v < lo ? v : lo; hi = v > hi ? v : hi; }
      mn += lo; mx += hi;
    }
    fprintf(stderr, " %s %.0f/%.0f", names[i], mn / grid, mx / grid);
  }
  {   // total cycles per warp, and per-warp (CTA 0) wait_h to see the skew pattern
    double tot = 0;
    for (int i = 0; i < 8; ++i) tot += acc[i] / grid / EPI_WARPS;
    fprintf(stderr, " | total %.0f | cta0 wait_h by warp:", tot);
    for (int w = 0; w < EPI_WARPS; ++w) fprintf(stderr, " %lld", host[w * 16 + 2]);
    fprintf(stderr, " | cta0 gelu by warp:");
    for (int w = 0; w < EPI_WARPS; ++w) fprintf(stderr, " %lld", host[w * 16 + 3]);
  }
  // issuer threads (proj kernel): cycles spent waiting, per CTA
  static const char* inames[16] = {"fc1:x1_ready", "fc1:h_empty", "fc1:r1_full", "fc1:total", "-", "-", "-", "-",
                                   "fc2:gelu_done", "fc2:r2_full", "fc2:x_full", "fc2:y_empty", "fc2:total", "-", "-", "-"};
  double iacc[16] = {0};
  for (unsigned b = 0; b < grid; ++b)
    for (int i = 0; i < 16; ++i) iacc[i] += static_cast<double>(host[148 * EPI_WARPS * 16 + b * 16 + i]);
  fprintf(stderr, " | issuers:");
  for (int i = 0; i < 16; ++i) if (inames[i][0] != '-') fprintf(stderr, " %s %.0f", inames[i], iacc[i] / grid);
  fprintf(stderr, "\n");
}

template <int C>
int launch_t(const MlpFusedPack& p, const __half* x, __half* out, int64_t M, cudaStream_t stream) {
  using K = Cfg<C>;
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(mlp_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    once.done();
  }
  alignas(64) CUtensorMap tmX;
  SUNET_TRY(make_tmap_2d_f16(&tmX, x, C, M, C, TILE_M));
  Params prm;
  prm.hconst = reinterpret_cast<const float2*>(p.hconst);
  prm.b2 = p.b2;
  prm.out = out;
  prm.M = M;
  prm.tiles = (M + TILE_M - 1) / TILE_M;
  const int sms = device_sms();
  const unsigned grid = static_cast<unsigned>(prm.tiles < sms ? prm.tiles : sms);
  prm.timing = mlp_timing_buf(stream);
  prm.trace = nullptr;
  SUNET_CUDA(launch_pdl(mlp_fused_kernel<C>, dim3(grid), dim3(THREADS), K::SMEM, stream, tmX, p.tmW1, p.tmW2, prm));
  mlp_timing_report("mlp_fused", C, grid, prm.timing, stream);
  return 0;
}

template <int C>
int launch_proj_t(const MlpFusedPack& p, const __half* attn_out, const __half* shortcut, __half* out, int64_t M, cudaStream_t stream) {
  using K = Cfg<C>;
  constexpr int SMEM = K::SMEM + C * 4;
  static_assert(SMEM <= 227 * 1024, "shared memory budget (proj variant)");
  static DeviceOnce once;
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(mlp_proj_fused_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    once.done();
  }
  alignas(64) CUtensorMap tmX;
  SUNET_TRY(make_tmap_2d_f16(&tmX, attn_out, C, M, C, TILE_M));
  ProjParams prm;
  prm.hbias = p.hbias;
  prm.b2 = p.b2;
  prm.bp = p.bp;
  prm.shortcut = shortcut;
  prm.out = out;
  prm.M = M;
  prm.tiles = (M + TILE_M - 1) / TILE_M;
  const int sms = device_sms();
  const unsigned grid = static_cast<unsigned>(prm.tiles < sms ? prm.tiles : sms);
  prm.timing = mlp_timing_buf(stream);
  prm.trace = prm.timing ? mlp_trace_buf(stream) : nullptr;
  SUNET_CUDA(launch_pdl(mlp_proj_fused_kernel<C>, dim3(grid), dim3(PTHREADS), SMEM, stream, tmX, p.tmWp, p.tmW1h, p.tmW2, prm));
  mlp_timing_report("mlp_proj_fused (wait_p, epi0, wait_h, gelu, wait_hs, store, wait_y, out)", C, grid, prm.timing, stream);
  mlp_trace_report(C, prm.trace, stream);
  return 0;
}

}  // namespace

bool mlp_fused_supported(int C) { return C == 96 || C == 192; }
int mlp_fused_w1h_pitch(int C) {
  switch (C) {
    case 96: return Cfg<96>::W1PITCH;
    case 192: return Cfg<192>::W1PITCH;
    default: return C;
  }
}

int mlp_fused_prepack(MlpFusedPack* p, int C, const float* gamma, const float* beta, const float* w1, const float* b1,
                      const float* w2, const float* b2, cudaStream_t stream) {
  if (!mlp_fused_supported(C)) return fail(SUNET_E_SHAPE, "fused mlp: C=%d not instantiated (96, 192)", C);
  if (!p->w1g || !p->w2 || !p->hconst || !p->b2) return fail(SUNET_E_ARG, "fused mlp: pack buffers not allocated");
  p->C = C;
  const int HID = 4 * C;
  mlp_fold_ln_kernel<<<(HID + 7) / 8, 256, 0, stream>>>(w1, b1, gamma, beta, p->w1g, reinterpret_cast<float2*>(p->hconst), HID, C);
  SUNET_CHECK_LAUNCH();
  cast_f16_kernel<<<148, 256, 0, stream>>>(w2, p->w2, static_cast<size_t>(C) * HID);
  SUNET_CHECK_LAUNCH();
  if (b2) SUNET_CUDA(cudaMemcpyAsync(p->b2, b2, C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  else SUNET_CUDA(cudaMemsetAsync(p->b2, 0, C * sizeof(float), stream));
  SUNET_TRY(make_tmap_2d_f16(&p->tmW1, p->w1g, C, HID, C, NC));
  SUNET_TRY(make_tmap_2d_f16(&p->tmW2, p->w2, HID, C, HID, C));
  return 0;
}

int mlp_fused_set_proj(MlpFusedPack* p, const float* wp, const float* bp, const float* gamma, const float* beta, const float* w1,
                       const float* b1, cudaStream_t stream) {
  if (!p->wp || !p->bp || !p->w1h || !p->hbias) return fail(SUNET_E_ARG, "fused mlp: proj pack buffers not allocated");
  const int C = p->C;
  const int pitch = mlp_fused_w1h_pitch(C);
  mlp_fold_half_kernel<<<(4 * C + 7) / 8, 256, 0, stream>>>(w1, b1, gamma, beta, p->w1h, p->hbias, 4 * C, C, pitch);
  SUNET_CHECK_LAUNCH();
  SUNET_TRY(make_tmap_2d_f16(&p->tmW1h, p->w1h, pitch, 4 * C, pitch, NC));
  cast_f16_kernel<<<148, 256, 0, stream>>>(wp, p->wp, static_cast<size_t>(C) * C);
  SUNET_CHECK_LAUNCH();
  if (bp) SUNET_CUDA(cudaMemcpyAsync(p->bp, bp, C * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  else SUNET_CUDA(cudaMemsetAsync(p->bp, 0, C * sizeof(float), stream));
  SUNET_TRY(make_tmap_2d_f16(&p->tmWp, p->wp, C, C, C, C));
  p->has_proj = 1;
  return 0;
}

int mlp_proj_fused_launch(const MlpFusedPack& p, const __half* attn_out, const __half* shortcut, __half* out, int64_t M, cudaStream_t stream) {
  if (!p.has_proj) return fail(SUNET_E_STATE, "fused mlp: proj weights not packed");
  if (M <= 0 || M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "fused mlp: bad row count %lld", (long long)M);
  if ((reinterpret_cast<uintptr_t>(attn_out) & 15) || (reinterpret_cast<uintptr_t>(shortcut) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return fail(SUNET_E_ALIGN, "fused mlp: attn_out/shortcut/out must be 16-byte aligned");
  if (attn_out == out) return fail(SUNET_E_ARG, "fused mlp: out must not alias attn_out (tiles of other CTAs are still being read)");
  switch (p.C) {
    case 96: return launch_proj_t<96>(p, attn_out, shortcut, out, M, stream);
    case 192: return launch_proj_t<192>(p, attn_out, shortcut, out, M, stream);
    default: return fail(SUNET_E_SHAPE, "fused mlp: C=%d not instantiated", p.C);
  }
}

int mlp_fused_launch(const MlpFusedPack& p, const __half* x, __half* out, int64_t M, cudaStream_t stream) {
  if (M <= 0 || M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "fused mlp: bad row count %lld", (long long)M);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(SUNET_E_ALIGN, "fused mlp: x/out must be 16-byte aligned");
  switch (p.C) {
    case 96: return launch_t<96>(p, x, out, M, stream);
    case 192: return launch_t<192>(p, x, out, M, stream);
    default: return fail(SUNET_E_SHAPE, "fused mlp: C=%d not instantiated", p.C);
  }
}

}  // namespace sunet
