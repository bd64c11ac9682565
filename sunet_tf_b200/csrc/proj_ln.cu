// See proj_ln.cuh.  Pipeline: warp 0 = TMA producer (A k-block [128][64] + the whole [C][64] weight k-block per ring stage),
// warp 1 = tcgen05 issuer (two N = C/2 instructions per k-step into one [128][C] fp32 accumulator), warps 2..17 = epilogue
// (thread <-> token row, warp <-> (TMEM lane quarter, column quarter)): bias + shortcut, fp16 rounding, X1 store, two-pass row
// statistics exchanged between the four column-quarter warps of a lane quarter through shared memory, normalise, T store.
// Global loads / stores of the epilogue go through a 2 KB per-warp staging tile so that every instruction covers whole 64-byte
// row segments (thread-per-row accesses are LSU-bound).
#include "proj_ln.cuh"

#include "device.h"
#include "error.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

namespace {

constexpr int TILE_M = 128;
constexpr int BK = 64;
constexpr int A_BYTES = TILE_M * BK * 2;
constexpr int EPI_WARPS = 16;
constexpr int THREADS = 64 + EPI_WARPS * 32;
constexpr int STG_BYTES = 2048;   // per warp: 32 rows x 32 fp16

template <int C, int KTOT>
struct Cfg {
  static constexpr int NH = C / 2;                 // columns per MMA instruction
  static constexpr int CQ = C / 4;                 // columns per epilogue column-quarter
  static constexpr int KB = KTOT / BK;             // k-blocks
  static constexpr int W_BYTES = C * BK * 2;       // one weight k-block
  static constexpr int STAGE = A_BYTES + W_BYTES;
  // The epilogue's staging tiles alias ring stage 0: the epilogue of a tile starts after its last MMA has read the ring, and the
  // producer does not refill the ring for the next tile before the epilogue has handed the accumulator back (acc_empty).  With
  // one tile per CTA (M <= 128 * SMs, every whole-model call) nothing is lost; more tiles per CTA serialise load and epilogue.
  static constexpr int STAGES = (226 * 1024 - 1024 - 2 * 4 * TILE_M * 4) / STAGE;
  static constexpr int OFF_STG = 0;
  static constexpr int OFF_RED = STAGES * STAGE;   // float [2][4][128]
  static constexpr int SMEM = OFF_RED + 2 * 4 * TILE_M * 4 + 1024;
  static constexpr uint32_t TMEM_COLS = C <= 128 ? 128 : (C <= 256 ? 256 : 512);
  static_assert(KTOT % BK == 0 && CQ % 32 == 0 && NH % 16 == 0 && NH <= 256 && C <= 512, "unsupported width");
  static_assert((NH * 128) % 1024 == 0, "second weight half must start on a swizzle atom");
  static_assert(STAGES >= 2 && EPI_WARPS * STG_BYTES <= STAGE, "ring too shallow / staging must fit ring stage 0");
};

struct Params {
  const __half* R;
  __half* X1;
  __half* T;
  const float* bias;
  const float* gamma;
  const float* beta;
  int64_t M;
  int64_t tiles;
};

__device__ __forceinline__ uint32_t stg_off(int row, int ch) { return static_cast<uint32_t>(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4)); }

template <int C, int KTOT, bool LN>
__global__ void __launch_bounds__(THREADS, 1)
    proj_ln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Params p) {
  using K = Cfg<C, KTOT>;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[K::STAGES], empty_bar[K::STAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < K::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, EPI_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, K::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  // only the threads that touch the predecessor's output wait for it (the TMA producer before its first activation load, the
  // epilogue warps before the shortcut reads); the weight k-blocks of the first ring stages are in flight before that wait
  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0, local = 0;
      constexpr int PRE = K::STAGES < K::KB ? K::STAGES : K::KB;
      if (static_cast<int64_t>(blockIdx.x) < p.tiles) {
        for (int kb = 0; kb < PRE; ++kb) {
          uint8_t* sb = smem + kb * K::STAGE + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[kb], K::STAGE);
          tma_load_2d(sb, &tmW, &full_bar[kb], kb * BK, 0);
          tma_load_2d(sb + K::NH * 128, &tmW, &full_bar[kb], kb * BK, K::NH);
        }
      }
      pdl_wait();
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
        const int m0 = static_cast<int>(tile * TILE_M);
        mbar_wait(&acc_empty, (local & 1) ^ 1);   // the previous tile's epilogue no longer uses the staging tiles inside stage 0
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * K::STAGE;
          uint8_t* sb = sa + A_BYTES;
          const bool early = local == 0 && kb < PRE;   // expect_tx and the weight k-block were issued before the dependency wait
          if (!early) mbar_arrive_expect_tx(&full_bar[s], K::STAGE);
          tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m0);
          if (!early) {
            tma_load_2d(sb, &tmW, &full_bar[s], kb * BK, 0);
            tma_load_2d(sb + K::NH * 128, &tmW, &full_bar[s], kb * BK, K::NH);
          }
          if (++s == K::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp walks the loop, one elected lane issues (a single-lane branch makes every descriptor thread-divergent
    // for the compiler: R2UR + an ELECT loop per tcgen05.mma, ~180 clk per instruction - see gemm_tcgen05.cu)
    {
      const uint32_t idesc = umma_idesc_f16(TILE_M, K::NH);
      int s = 0;
      uint32_t ph = 0, local = 0;
      for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
        mbar_wait(&acc_empty, (local & 1) ^ 1);   // the epilogue has drained the accumulator of the previous tile
        tc_fence_after();
        for (int kb = 0; kb < K::KB; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * K::STAGE);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc0 = umma_desc_sw128(sa + A_BYTES);
          const uint64_t bdesc1 = umma_desc_sw128(sa + A_BYTES + K::NH * 128);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint32_t accum = (kb > 0 || k > 0) ? 1u : 0u;
              umma_f16_ss(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc0 + static_cast<uint64_t>(2 * k), idesc, accum);
              umma_f16_ss(tmem_base + K::NH, adesc + static_cast<uint64_t>(2 * k), bdesc1 + static_cast<uint64_t>(2 * k), idesc, accum);
            }
            tc_commit(&empty_bar[s]);
            if (kb == K::KB - 1) tc_commit(&acc_full);
          }
          __syncwarp();
          if (++s == K::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    pdl_wait();
    const int e = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int quarter = e >> 2;             // column quarter
    const int col0 = quarter * K::CQ;
    const uint32_t stg = smem_u32(smem + K::OFF_STG + e * STG_BYTES);
    float* red = reinterpret_cast<float*>(smem + K::OFF_RED);   // [2][4][128]
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + col0;
    const int cr = lane >> 2, cch = lane & 3;   // copy role: row cr + 8 it, 16-byte chunk cch
    uint32_t local = 0;
    for (int64_t tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
      const int64_t m_base = tile * TILE_M + q * 32;
      const int rows_valid = static_cast<int>(min(static_cast<int64_t>(32), p.M - m_base));
      mbar_wait(&acc_full, local & 1);
      tc_fence_after();
      float sum = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < K::CQ; cc += 32) {
        // shortcut rows into the staging tile, whole 64-byte segments per 4 lanes
        const __half* rbase = p.R + m_base * C + col0 + cc;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 8 + cr;
          if (r < rows_valid) sts128(stg + stg_off(r, cch), *(reinterpret_cast<const uint4*>(rbase + static_cast<int64_t>(r) * C) + cch));   // plain load: X1 may alias R
        }
        uint32_t v[32];
        tmem_ld32(taddr + cc, v);
        tmem_ld_wait();
        __syncwarp();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const uint4 rr = lds128(stg + stg_off(lane, ch));
          const __half2* r2 = reinterpret_cast<const __half2*>(&rr);
          float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
          if (p.bias != nullptr) {
            b0 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cc + ch * 8));
            b1 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + cc + ch * 8 + 4));
          }
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(r2[j]);
            const __half2 h = __floats2half2_rn(__uint_as_float(v[ch * 8 + 2 * j]) + bb[2 * j] + t.x,
                                                __uint_as_float(v[ch * 8 + 2 * j + 1]) + bb[2 * j + 1] + t.y);
            const float2 back = __half22float2(h);   // statistics on the rounded values, as the stand-alone LayerNorm reads them
            sum += back.x + back.y;
            ow[j] = *reinterpret_cast<const uint32_t*>(&h);
            v[ch * 8 + 2 * j] = __float_as_uint(back.x);
            v[ch * 8 + 2 * j + 1] = __float_as_uint(back.y);
          }
          sts128(stg + stg_off(lane, ch), o);
        }
        if (LN) tmem_st32(taddr + cc, v);   // X1 (rounded) parked over the accumulator for the two statistics passes
        __syncwarp();
        __half* xbase = p.X1 + m_base * C + col0 + cc;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 8 + cr;
          if (r < rows_valid) *(reinterpret_cast<uint4*>(xbase + static_cast<int64_t>(r) * C) + cch) = lds128(stg + stg_off(r, cch));
        }
        __syncwarp();
      }
      if (LN) {
      tmem_st_wait();
      // ---- row statistics across the four column-quarter warps of this lane quarter (two-pass: mean, then centred squares)
      red[quarter * TILE_M + row] = sum;
      named_bar_sync(1 + q, 128);
      const float mean = (red[row] + red[TILE_M + row] + red[2 * TILE_M + row] + red[3 * TILE_M + row]) * (1.f / C);
      float sq = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < K::CQ; cc += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = __uint_as_float(v[i]) - mean; sq += d * d; }
      }
      float* red2 = red + 4 * TILE_M;
      red2[quarter * TILE_M + row] = sq;
      named_bar_sync(1 + q, 128);
      const float rstd = rsqrtf((red2[row] + red2[TILE_M + row] + red2[2 * TILE_M + row] + red2[3 * TILE_M + row]) * (1.f / C) + 1e-5f);
#pragma unroll 1
      for (int cc = 0; cc < K::CQ; cc += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + cc + ch * 8));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + col0 + cc + ch * 8 + 4));
          const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + cc + ch * 8));
          const float4 e1 = __ldg(reinterpret_cast<const float4*>(p.beta + col0 + cc + ch * 8 + 4));
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float x0 = __uint_as_float(v[ch * 8 + 2 * j]), x1 = __uint_as_float(v[ch * 8 + 2 * j + 1]);
            const __half2 h = __floats2half2_rn((x0 - mean) * rstd * gg[2 * j] + ee[2 * j], (x1 - mean) * rstd * gg[2 * j + 1] + ee[2 * j + 1]);
            ow[j] = *reinterpret_cast<const uint32_t*>(&h);
          }
          sts128(stg + stg_off(lane, ch), o);
        }
        __syncwarp();
        __half* tbase = p.T + m_base * C + col0 + cc;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 8 + cr;
          if (r < rows_valid) *(reinterpret_cast<uint4*>(tbase + static_cast<int64_t>(r) * C) + cch) = lds128(stg + stg_off(r, cch));
        }
        __syncwarp();
      }
      }   // LN
      // all TMEM reads of this warp are complete: hand the accumulator (and the staging tiles inside ring stage 0, which the
      // next tile's TMA loads overwrite through the async proxy) back
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty);
      // red / red2 are rewritten by the next tile only after both named barriers of that tile: every reader of this tile's
      // values is past its reads by then (the second barrier orders red, the first barrier of the next tile orders red2)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, K::TMEM_COLS);
  }
}

template <int C, int KTOT, bool LN>
int launch_t(const ProjLnPack& pk, const __half* A, const __half* R, __half* X1, __half* T, int64_t M, cudaStream_t stream) {
  using K = Cfg<C, KTOT>;
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute((proj_ln_kernel<C, KTOT, LN>), cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM));
    once.done();
  }
  alignas(64) CUtensorMap tmA, tmW;
  SUNET_TRY(make_tmap_2d_f16(&tmA, A, KTOT, M, KTOT, TILE_M));
  SUNET_TRY(make_tmap_2d_f16(&tmW, pk.w, KTOT, C, KTOT, K::NH));
  Params p;
  p.R = R; p.X1 = X1; p.T = T; p.bias = pk.bias; p.gamma = pk.gamma; p.beta = pk.beta; p.M = M;
  p.tiles = (M + TILE_M - 1) / TILE_M;
  const int sms = device_sms();
  const unsigned grid = static_cast<unsigned>(p.tiles < sms ? p.tiles : sms);
  SUNET_CUDA(launch_pdl(proj_ln_kernel<C, KTOT, LN>, dim3(grid), dim3(THREADS), K::SMEM, stream, tmA, tmW, p));
  SUNET_CHECK_LAUNCH();
  return 0;
}

}  // namespace

bool proj_ln_supported(int C) { return C == 384; }
bool row_gemm_supported(int C, int Kdim) { return C == 384 && Kdim == 1536; }

int proj_ln_launch(const ProjLnPack& p, const __half* A, const __half* R, __half* X1, __half* T, int64_t M, cudaStream_t stream) {
  if (M <= 0) return 0;
  if (M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "proj_ln: M too large for 32-bit TMA coordinates");
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(X1) | reinterpret_cast<uintptr_t>(T)) & 15)
    return fail(SUNET_E_ALIGN, "proj_ln: operands must be 16-byte aligned");
  if (p.C == 384 && p.gamma != nullptr && p.beta != nullptr && T != nullptr) return launch_t<384, 384, true>(p, A, R, X1, T, M, stream);
  return fail(SUNET_E_SHAPE, "proj_ln: C=%d not instantiated", p.C);
}

int row_gemm_residual_launch(const __half* W, const float* bias, int C, int Kdim, const __half* A, const __half* R, __half* X,
                             int64_t M, cudaStream_t stream) {
  if (M <= 0) return 0;
  if (M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "row_gemm: M too large for 32-bit TMA coordinates");
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(R) | reinterpret_cast<uintptr_t>(X)) & 15)
    return fail(SUNET_E_ALIGN, "row_gemm: operands must be 16-byte aligned");
  ProjLnPack p;
  p.w = W; p.bias = bias; p.C = C;
  if (C == 384 && Kdim == 1536) return launch_t<384, 1536, false>(p, A, R, X, nullptr, M, stream);
  return fail(SUNET_E_SHAPE, "row_gemm: C=%d K=%d not instantiated", C, Kdim);
}

}  // namespace sunet
