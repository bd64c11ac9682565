// Persistent warp-specialised tcgen05 GEMM (sm_100a):
//   TMA (128B swizzle) -> smem ring -> tcgen05.mma (fp32 accumulators, double-buffered in TMEM)
//   -> tcgen05.ld epilogue (bias / GELU / PReLU / residual) -> global.
//   warp 0    : TMA producer (one elected lane); runs ahead across tiles, bounded only by the smem ring
//   warp 1    : TMEM allocator + MMA issuer (one elected lane)
//   warps 2-17: epilogue; warp w owns TMEM lanes [32*(w&3), +32) == tile rows and one quarter of the tile columns
// One CTA per SM loops over 128 x block_n output tiles (tile = blockIdx.x + i * gridDim.x, n fastest so that
// co-running CTAs share A rows in L2).  While the epilogue drains accumulator stage s, the MMA warp fills stage s^1
// and the producer prefetches the operands of the tiles after that.
#include "act.cuh"
#include "device.h"
#include "error.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 fp16 = one 128-byte swizzle row
static constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
static constexpr int MAX_STAGES = 12;
static constexpr int EPI_WARPS = 16;
static constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
static constexpr int GEMM_MAX_DYN_SMEM = 226 * 1024;  // 227 KB per CTA minus the static barriers

// Per-warp staging tile: 32 rows x 128 bytes, 16-byte chunk j of row r lives at chunk slot (j ^ (r & 7)).
// Row-wise accesses (thread = row) and coalesced accesses (consecutive lanes = consecutive chunks of a row) are both
// bank-conflict free for 8 chunks per row (2-way at most for narrower groups).
__device__ __forceinline__ uint32_t stage_off(int r, int ch) { return static_cast<uint32_t>((r << 7) + ((ch ^ (r & 7)) << 4)); }

// ---- epilogue math on 16 accumulator columns of one row: bias, activation, residual -> staging row (fp16 or fp32).
// The residual chunk (fp16 output only) sits in the staging slot that the result overwrites.
template <bool F32>
__device__ __forceinline__ void epilogue_math16(const GemmEpi& p, const uint32_t* v, int n, float slope, uint32_t stg_row, int lane,
                                                int col_in_group) {
  float f[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
      f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
    }
  }
  if constexpr (!F32) {
    const int ch0 = col_in_group >> 3;  // first of the two 16-byte chunks of these 16 fp16 columns
    if (p.act == ACT_GELU && p.R == nullptr) {
      // fc1 path: round the pre-activation to fp16 pairs and evaluate GELU on packed halves
      uint4 o[2];
      uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half2 x2 = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
        ow[j] = gelu_fast_h2(*reinterpret_cast<const uint32_t*>(&x2));
      }
      sts128(stg_row + (((ch0) ^ (lane & 7)) << 4), o[0]);
      sts128(stg_row + (((ch0 + 1) ^ (lane & 7)) << 4), o[1]);
      return;
    }
    if (p.act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = gelu_fast(f[j]);
    } else if (p.act == ACT_PRELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = f[j] >= 0.f ? f[j] : slope * f[j];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t slot = stg_row + (((ch0 + h) ^ (lane & 7)) << 4);
      if (p.R != nullptr) {
        const uint4 rr = lds128(slot);
        const __half2* r2 = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(r2[j]);
          f[h * 8 + 2 * j] += t.x;
          f[h * 8 + 2 * j + 1] += t.y;
        }
      }
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(f[h * 8 + 2 * j], f[h * 8 + 2 * j + 1]);
      sts128(slot, o);
    }
  } else {
    if (p.act == ACT_GELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = gelu_fast(f[j]);
    } else if (p.act == ACT_PRELU) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = f[j] >= 0.f ? f[j] : slope * f[j];
    }
    const int ch0 = col_in_group >> 2;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sts128(stg_row + (((ch0 + j) ^ (lane & 7)) << 4),
             make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]), __float_as_uint(f[4 * j + 3])));
  }
}

// One group of GC accumulator columns for the 32 rows of this warp; GC * esize is 32, 64 or 128 bytes per row.
template <int GC, bool F32>
__device__ __forceinline__ void epilogue_group(const GemmEpi& p, uint32_t stg, uint32_t taddr, int c, int64_t m_base,
                                               int rows_valid, int n0, int lane, float slope) {
  constexpr int ESIZE = F32 ? 4 : 2;
  constexpr int OCH = GC * ESIZE / 16;   // 16-byte chunks per output row (power of two: 2, 4 or 8)
  const uint32_t stg_row = stg + (lane << 7);
  if (!F32 && p.R != nullptr) {
    // coalesced residual load into the staging tile (consecutive lanes -> consecutive chunks of a row)
    const __half* rbase = p.R + m_base * p.ldr + n0 + c;
#pragma unroll
    for (int k = 0; k < OCH; ++k) {
      const int i = k * 32 + lane;
      const int r = i / OCH, ch = i % OCH;
      if (r < rows_valid) sts128(stg + stage_off(r, ch), __ldg(reinterpret_cast<const uint4*>(rbase + r * p.ldr) + ch));
    }
    __syncwarp();
  }
  if constexpr (GC >= 32) {
#pragma unroll
    for (int cc = 0; cc < GC; cc += 32) {
      uint32_t v[32];
      tmem_ld32(taddr + c + cc, v);
      tmem_ld_wait();
      epilogue_math16<F32>(p, v, n0 + c + cc, slope, stg_row, lane, cc);
      epilogue_math16<F32>(p, v + 16, n0 + c + cc + 16, slope, stg_row, lane, cc + 16);
    }
  } else {
    uint32_t v[16];
    tmem_ld16(taddr + c, v);
    tmem_ld_wait();
    epilogue_math16<F32>(p, v, n0 + c, slope, stg_row, lane, 0);
  }
  __syncwarp();
  uint8_t* cbase = static_cast<uint8_t*>(p.C) + (m_base * p.ldc + n0 + c) * ESIZE;
#pragma unroll
  for (int k = 0; k < OCH; ++k) {
    const int i = k * 32 + lane;
    const int r = i / OCH, ch = i % OCH;
    if (r < rows_valid) *reinterpret_cast<uint4*>(cbase + static_cast<int64_t>(r) * p.ldc * ESIZE + (ch << 4)) = lds128(stg + stage_off(r, ch));
  }
  __syncwarp();
}

template <bool F32>
__device__ __forceinline__ void epilogue_range(const GemmEpi& p, uint32_t stg, uint32_t taddr, int c_begin, int c_end, int64_t m_base,
                                               int rows_valid, int n0, int lane, float slope) {
  constexpr int GMAX = F32 ? 32 : 64;  // 128-byte staging rows
  int c = c_begin;
  for (; c + GMAX <= c_end; c += GMAX) epilogue_group<GMAX, F32>(p, stg, taddr, c, m_base, rows_valid, n0, lane, slope);
  if constexpr (!F32) {
    if (c + 32 <= c_end) { epilogue_group<32, F32>(p, stg, taddr, c, m_base, rows_valid, n0, lane, slope); c += 32; }
  }
  if (c + 16 <= c_end) epilogue_group<16, F32>(p, stg, taddr, c, m_base, rows_valid, n0, lane, slope);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tn_f16_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmW, const GemmEpi p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int block_n = p.block_n;
  const int stages = p.stages;
  const uint32_t stage_bytes = A_TILE_BYTES + block_n * BLOCK_K * 2;
  const int kb0 = (p.K0 + BLOCK_K - 1) / BLOCK_K;
  const int kb1 = (p.K1 + BLOCK_K - 1) / BLOCK_K;
  const int num_kb = kb0 + kb1;
  const uint32_t acc_cols = p.acc_cols;          // TMEM columns per accumulator stage (>= block_n)
  const uint32_t tmem_cols = p.tmem_cols;        // power of two >= 2 * acc_cols
  const int64_t total_tiles = p.total_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.K1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();
  // Programmatic dependent launch: only the threads that touch the predecessor's output wait for it.  The TMA producer first
  // issues the WEIGHT halves of its first ring stages (parameters, constant across the forward), then waits, then the activation
  // halves - the weight bytes (2/3 of a stage at 256-wide tiles) arrive under the previous kernel's tail.
  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // global k-block counter: the ring runs across tiles
      int s = 0;
      uint32_t ph = 0;
      uint32_t pre = 0;   // k-blocks of the first tile whose expect_tx + weight load were issued before the dependency wait
      if (static_cast<int64_t>(blockIdx.x) < total_tiles && !(p.dbg & 4)) {
        const int n0 = static_cast<int>(blockIdx.x % p.n_tiles) * block_n;
        pre = static_cast<uint32_t>(min(stages, num_kb));
        for (uint32_t kb = 0; kb < pre; ++kb) {
          uint8_t* sb = smem + kb * stage_bytes + A_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[kb], stage_bytes);
          const int kw = static_cast<int>(kb) < kb0 ? static_cast<int>(kb) * BLOCK_K : p.K0 + (static_cast<int>(kb) - kb0) * BLOCK_K;
          tma_load_2d(sb, &tmW, &full_bar[kb], kw, n0);
        }
      }
      pdl_wait();
      uint32_t ready = mbar_test(&empty_bar[0], 1);
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = static_cast<int>(tile % p.n_tiles) * block_n;
        const int m0 = static_cast<int>(tile / p.n_tiles) * BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          mbar_wait_hint(&empty_bar[s], ph ^ 1, ready);
          const int s_cur = s;
          if (++s == stages) { s = 0; ph ^= 1; }
          ready = mbar_test(&empty_bar[s], ph ^ 1);   // next slot's state, looked up under this slot's TMA issue
          uint8_t* sa = smem + s_cur * stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          if (p.dbg & 4) { mbar_arrive(&full_bar[s_cur]); continue; }
          const bool early = it < pre;   // expect_tx and the weight half are already in flight
          if (!early) mbar_arrive_expect_tx(&full_bar[s_cur], stage_bytes);
          if (kb < kb0) {
            tma_load_2d(sa, &tmA0, &full_bar[s_cur], kb * BLOCK_K, m0);
            if (!early) tma_load_2d(sb, &tmW, &full_bar[s_cur], kb * BLOCK_K, n0);
          } else {
            const int j = kb - kb0;
            tma_load_2d(sa, &tmA1, &full_bar[s_cur], j * BLOCK_K, m0);
            if (!early) tma_load_2d(sb, &tmW, &full_bar[s_cur], p.K0 + j * BLOCK_K, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer.  The WHOLE warp runs this loop convergently and one elected lane issues: with a single-lane branch around the
    // loop (`if (lane == 0)`) every value is thread-divergent for the compiler, which then moves each descriptor to the uniform
    // datapath with R2UR and wraps each tcgen05.mma / commit in an ELECT loop - ~180 clk per MMA instruction, i.e. the issue thread,
    // not the tensor pipe or the operand stream, paced every k-block (tools/gemm_dbg_big.sh: 720 clk per k-block for N = 128, 192
    // and 256 alike with the TMA loads switched off).
    const uint32_t idesc = umma_idesc_f16(BLOCK_M, block_n);
    uint32_t local = 0;
    int s = 0;
    uint32_t ph = 0;
    uint32_t ready = mbar_test(&full_bar[0], 0);
    uint32_t acc_ready = mbar_test(&tmem_empty_bar[0], 1);
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const uint32_t as = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      mbar_wait_hint(&tmem_empty_bar[as], aph ^ 1, acc_ready);   // epilogue has drained this accumulator stage
      acc_ready = mbar_test(&tmem_empty_bar[as ^ 1], (((local + 1) >> 1) & 1) ^ 1);   // next tile's stage, looked up early
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * acc_cols;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait_hint(&full_bar[s], ph, ready);
        const int s_cur = s;
        if (++s == stages) { s = 0; ph ^= 1; }
        ready = mbar_test(&full_bar[s], ph);   // next slot's state, looked up under this slot's MMA issue
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s_cur * stage_bytes);
        const uint32_t sb = sa + A_TILE_BYTES;
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sb);
        // valid K elements in this block (a K0/K1 tail shorter than 64 is zero-filled by TMA but not multiplied)
        const int kvalid = kb < kb0 ? min(BLOCK_K, p.K0 - kb * BLOCK_K) : min(BLOCK_K, p.K1 - (kb - kb0) * BLOCK_K);
        const int ksteps = (p.dbg & 2) ? 0 : (kvalid + 15) >> 4;
        if (elect_one()) {
          // advance 16 fp16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          } else {
            for (int k = 0; k < ksteps; ++k)
              umma_f16_ss(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[s_cur]);  // frees the smem slot once these MMAs have read it
          if (kb == num_kb - 1) tc_commit(&tmem_full_bar[as]);  // accumulator of this tile complete
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue: thread <-> one row of the tile, warp <-> (lane quadrant, column half)
    pdl_wait();   // residual reads / output writes below must follow the predecessor
    const int q = warp & 3;
    const int quarter = (warp - 2) >> 2;                    // 0..3: which slice of the tile columns
    const int units = block_n >> 4;                          // 16-column units, split as evenly as possible
    const int c_begin = (quarter * units >> 2) << 4;
    const int c_end = ((quarter + 1) * units >> 2) << 4;
    const int row = q * 32 + lane;
    const uint32_t stg = smem_u32(smem + static_cast<size_t>(stages) * stage_bytes + (warp - 2) * 4096);  // this warp's staging tile
    float slope = 0.f;
    if (p.act == ACT_PRELU) slope = __ldg(p.prelu);
    uint32_t local = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const uint32_t as = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      const int n0 = static_cast<int>(tile % p.n_tiles) * block_n;
      const int64_t m = static_cast<int64_t>(tile / p.n_tiles) * BLOCK_M + row;
      const uint32_t taddr = tmem_base + as * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      const int64_t m_base = m - lane;
      const int rows_valid = static_cast<int>(min(static_cast<int64_t>(32), p.M - m_base));  // <= 0 for fully out-of-range warps
      if (p.dbg & 1) { /* timing ablation: accumulator handed straight back */ }
      else if (p.out_f32) epilogue_range<true>(p, stg, taddr, c_begin, c_end, m_base, rows_valid, n0, lane, slope);
      else epilogue_range<false>(p, stg, taddr, c_begin, c_end, m_base, rows_valid, n0, lane, slope);
      // all TMEM reads of this warp are complete (wait::ld above): hand the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ CTA-pair variant
// Same pipeline with tcgen05 cta_group::2: two CTAs of a cluster (one TPC) work on one 256 x block_n tile.  Each CTA
// stages its own 128 rows of A and HALF of the weight tile (block_n / 2 rows), the leader's MMA warp issues M = 256
// instructions that read both CTAs' shared memory and write both CTAs' TMEM, and each CTA drains its own 128 rows.
// Per CTA this halves the weight bytes pulled through L2 -> smem and the smem a ring stage costs (deeper ring).
// Barriers: full[] lives in the leader (both CTAs' TMA loads complete on it), empty[] / tmem_full[] exist in both CTAs
// and are signalled by multicast commits, tmem_empty[] lives in the leader and collects both CTAs' epilogue warps.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    gemm_tn_f16_pair_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                            const __grid_constant__ CUtensorMap tmW, const GemmEpi p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar[2];
  __shared__ __align__(8) uint64_t tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0: leader (issues the MMAs)
  const int block_n = p.block_n;
  const int half_n = block_n >> 1;
  const int stages = p.stages;
  const uint32_t stage_bytes = A_TILE_BYTES + half_n * BLOCK_K * 2;
  const int kb0 = (p.K0 + BLOCK_K - 1) / BLOCK_K;
  const int kb1 = (p.K1 + BLOCK_K - 1) / BLOCK_K;
  const int num_kb = kb0 + kb1;
  const uint32_t acc_cols = p.acc_cols;
  const uint32_t tmem_cols = p.tmem_cols;
  const int64_t total_tiles = p.total_tiles;     // 256 x block_n tiles
  const int64_t cluster_id = blockIdx.x >> 1;
  const int64_t n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.K1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(&tmem_base_smem, tmem_cols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();   // barriers of both CTAs initialised before any remote arrive / peer TMA completion
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = cluster_id; tile < total_tiles; tile += n_clusters) {
        const int n0 = static_cast<int>(tile % p.n_tiles) * block_n + static_cast<int>(rank) * half_n;
        const int m0 = static_cast<int>(tile / p.n_tiles) * (2 * BLOCK_M) + static_cast<int>(rank) * BLOCK_M;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_bytes);   // both CTAs' bytes land on the leader's barrier
          if (kb < kb0) {
            tma_load_2d_pair(sa, &tmA0, &full_bar[s], kb * BLOCK_K, m0);
            tma_load_2d_pair(sb, &tmW, &full_bar[s], kb * BLOCK_K, n0);
          } else {
            const int j = kb - kb0;
            tma_load_2d_pair(sa, &tmA1, &full_bar[s], j * BLOCK_K, m0);
            tma_load_2d_pair(sb, &tmW, &full_bar[s], p.K0 + j * BLOCK_K, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {   // the leader's MMA warp walks the loop convergently, one elected lane issues (see gemm_tn_f16_kernel)
      const uint32_t idesc = umma_idesc_f16(2 * BLOCK_M, block_n);
      uint32_t it = 0;
      uint32_t local = 0;
      for (int64_t tile = cluster_id; tile < total_tiles; tile += n_clusters, ++local) {
        const uint32_t as = local & 1;
        const uint32_t aph = (local >> 1) & 1;
        mbar_wait(&tmem_empty_bar[as], aph ^ 1);   // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * acc_cols;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % stages;
          const uint32_t ph = (it / stages) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * stage_bytes);
          const uint32_t sb = sa + A_TILE_BYTES;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sb);
          const int kvalid = kb < kb0 ? min(BLOCK_K, p.K0 - kb * BLOCK_K) : min(BLOCK_K, p.K1 - (kb - kb0) * BLOCK_K);
          const int ksteps = (kvalid + 15) >> 4;
          if (elect_one()) {
            if (ksteps == 4) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            } else {
              for (int k = 0; k < ksteps; ++k)
                umma_f16_ss_pair(d_tmem, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
            tc_commit_pair(&empty_bar[s], 3);        // frees the slot in both CTAs
            if (kb == num_kb - 1) tc_commit_pair(&tmem_full_bar[as], 3);     // accumulator complete: wake both CTAs' epilogues
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int q = warp & 3;
    const int quarter = (warp - 2) >> 2;
    const int units = block_n >> 4;
    const int c_begin = (quarter * units >> 2) << 4;
    const int c_end = ((quarter + 1) * units >> 2) << 4;
    const int row = q * 32 + lane;
    const uint32_t stg = smem_u32(smem + static_cast<size_t>(stages) * stage_bytes + (warp - 2) * 4096);
    float slope = 0.f;
    if (p.act == ACT_PRELU) slope = __ldg(p.prelu);
    uint32_t local = 0;
    for (int64_t tile = cluster_id; tile < total_tiles; tile += n_clusters, ++local) {
      const uint32_t as = local & 1;
      const uint32_t aph = (local >> 1) & 1;
      const int n0 = static_cast<int>(tile % p.n_tiles) * block_n;
      const int64_t m = static_cast<int64_t>(tile / p.n_tiles) * (2 * BLOCK_M) + rank * BLOCK_M + row;
      const uint32_t taddr = tmem_base + as * acc_cols + (static_cast<uint32_t>(q * 32) << 16);
      mbar_wait(&tmem_full_bar[as], aph);
      tc_fence_after();
      const int64_t m_base = m - lane;
      const int rows_valid = static_cast<int>(min(static_cast<int64_t>(32), p.M - m_base));
      if (p.out_f32) epilogue_range<true>(p, stg, taddr, c_begin, c_end, m_base, rows_valid, n0, lane, slope);
      else epilogue_range<false>(p, stg, taddr, c_begin, c_end, m_base, rows_valid, n0, lane, slope);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[as], 0);   // the leader's MMA warp owns the accumulator hand-back
    }
  }
  tc_fence_before();
  cluster_sync_all();   // neither CTA may exit (or free TMEM) while the other can still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                     uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return fail(SUNET_E_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * 2) % 16 != 0)
    return fail(SUNET_E_ALIGN, "TMA operand needs a 16-byte aligned base and row stride (base=%p stride=%llu elems)", base,
                (unsigned long long)row_stride_elems);
  cuuint64_t gdim[2] = {inner, rows};
  cuuint64_t gstr[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(SUNET_E_DRIVER, "cuTensorMapEncodeTiled failed (CUresult %d) inner=%llu rows=%llu stride=%llu box_rows=%u",
                (int)r, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)row_stride_elems, box_rows);
  return 0;
}

static int num_sms() { return device_sms(); }

static int pick_block_n(int N, int64_t m_tiles, int K) {
  static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  // Measured on B200 (profiles/r01_gemm_*.log): wide tiles win whenever they still give ~0.6 tiles per SM (the MMA
  // needs N >= 128 to stay off the smem-read bound, and one wide epilogue beats several narrow ones); below that,
  // fall back to the widest tile <= 128 to spread the small-M stages over more SMs.
  const int64_t want = (6 * static_cast<int64_t>(num_sms())) / 10;
  for (int c : cand)
    if (N % c == 0 && m_tiles * (N / c) >= want) {
      // a single partial wave of 256-wide tiles: 192-wide tiles that still fit one wave finish sooner (stage-3 proj / fc2,
      // M = 4096, N = 768: 96 tiles -> 128 tiles, 31.0 -> 29.1 us; tools/gemm_stage_sweep.sh)
      if (c == 256 && N % 192 == 0 && m_tiles * (N / 256) < num_sms() && m_tiles * (N / 192) <= num_sms()) return 192;
      return c;
    }
  (void)K;
  for (int c : cand)
    if (N % c == 0 && c <= 128) return c;
  for (int c : cand)
    if (N % c == 0) return c;
  return 0;
}

int gemm_prepare(const GemmArgs& a, GemmOp* op) {
  if (a.M <= 0 || a.N <= 0 || a.K0 <= 0 || a.K1 < 0) return fail(SUNET_E_SHAPE, "gemm: bad shape M=%lld N=%d K0=%d K1=%d", (long long)a.M, a.N, a.K0, a.K1);
  if (a.N % 16 != 0) return fail(SUNET_E_SHAPE, "gemm: N=%d must be a multiple of 16", a.N);
  if (a.K0 % 16 != 0 || a.K1 % 16 != 0) return fail(SUNET_E_SHAPE, "gemm: K0=%d / K1=%d must be multiples of 16", a.K0, a.K1);
  if (a.ldc % 8 != 0 || (a.R && a.ldr % 8 != 0)) return fail(SUNET_E_ALIGN, "gemm: ldc/ldr must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(a.C) & 15) || (reinterpret_cast<uintptr_t>(a.R) & 15) ||
      (reinterpret_cast<uintptr_t>(a.bias) & 15))
    return fail(SUNET_E_ALIGN, "gemm: C/R/bias must be 16-byte aligned");
  if (a.M > (int64_t)0x7fffff00) return fail(SUNET_E_SHAPE, "gemm: M too large for 32-bit TMA coordinates");
  const bool pair = a.pair_mode == 2;   // cta_group::2: 256-row tiles shared by a CTA pair
  const int64_t m_tiles = pair ? (a.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) : (a.M + BLOCK_M - 1) / BLOCK_M;
  int bn = a.force_block_n ? a.force_block_n : pick_block_n(a.N, m_tiles, a.K0 + a.K1);
  if (pair && bn % 32 != 0) return fail(SUNET_E_SHAPE, "gemm: the CTA-pair path needs a tile width that is a multiple of 32 (bn=%d)", bn);
  if (bn == 0 || a.N % bn != 0 || bn % 16 != 0 || bn > 256) return fail(SUNET_E_SHAPE, "gemm: no tile width for N=%d (bn=%d)", a.N, bn);
  const int n_tiles = a.N / bn;
  if (m_tiles * n_tiles > 0x7fffffffLL) return fail(SUNET_E_SHAPE, "gemm: grid too large");
  const int num_kb = (a.K0 + BLOCK_K - 1) / BLOCK_K + (a.K1 + BLOCK_K - 1) / BLOCK_K;
  const int stage_bytes = A_TILE_BYTES + (pair ? bn / 2 : bn) * BLOCK_K * 2;
  int stages = (GEMM_MAX_DYN_SMEM - 1024 - EPI_WARPS * 4096) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  static const int stage_cap = [] { const char* cap = getenv("SUNET_GEMM_STAGES"); return cap ? atoi(cap) : 0; }();   // pipeline-depth experiments (tools/gemm_stage_sweep.sh); read once
  if (stage_cap >= 1 && stage_cap < stages) stages = stage_cap;
  if (stages < 1) stages = 1;
  SUNET_TRY(make_tmap_2d_f16(&op->tmA0, a.A0, a.K0, a.M, a.lda0, BLOCK_M));
  if (a.K1 > 0) SUNET_TRY(make_tmap_2d_f16(&op->tmA1, a.A1, a.K1, a.M, a.lda1, BLOCK_M));
  else op->tmA1 = op->tmA0;
  SUNET_TRY(make_tmap_2d_f16(&op->tmW, a.W, a.K0 + a.K1, a.N, a.ldw, pair ? bn / 2 : bn));
  GemmEpi& e = op->epi;
  e.M = a.M; e.N = a.N; e.K0 = a.K0; e.K1 = a.K1; e.block_n = bn; e.stages = stages;
  e.bias = a.bias; e.prelu = a.prelu; e.act = a.act; e.R = a.R; e.ldr = a.ldr; e.C = a.C; e.ldc = a.ldc;
  e.out_f32 = a.out_f32; e.n_tiles = n_tiles;
  e.total_tiles = m_tiles * n_tiles;
  e.acc_cols = bn <= 32 ? 32 : (bn <= 64 ? 64 : (bn <= 128 ? 128 : 256));
  e.tmem_cols = 2 * e.acc_cols;
  e.dbg = a.dbg;
  if (a.act == ACT_PRELU && a.prelu == nullptr) return fail(SUNET_E_ARG, "gemm: PReLU needs a slope pointer");
  if (a.out_f32 && a.R != nullptr) return fail(SUNET_E_ARG, "gemm: a residual with fp32 output is not supported");
  const int64_t tiles = m_tiles * n_tiles;
  op->pair = pair ? 1 : 0;
  if (pair) {
    const int64_t clusters = num_sms() / 2;
    op->grid = static_cast<unsigned>(2 * (tiles < clusters ? tiles : clusters));
  } else {
    op->grid = static_cast<unsigned>(tiles < num_sms() ? tiles : num_sms());
  }
  // never keep more ring stages than this CTA will ever fill
  const int64_t workers = pair ? op->grid / 2 : op->grid;
  const int64_t tiles_per_cta = (tiles + workers - 1) / workers;
  if (static_cast<int64_t>(stages) > tiles_per_cta * num_kb) { stages = static_cast<int>(tiles_per_cta * num_kb); e.stages = stages; }
  op->smem = stages * stage_bytes + 1024 + EPI_WARPS * 4096;
  op->flops = 2.0 * (double)a.M * a.N * (a.K0 + a.K1);
  return 0;
}

int gemm_launch(const GemmOp& op, cudaStream_t stream) {
  static DeviceOnce once;   // the shared-memory opt-in is per device
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(gemm_tn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_MAX_DYN_SMEM));
    SUNET_CUDA(cudaFuncSetAttribute(gemm_tn_f16_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_MAX_DYN_SMEM));
    once.done();
  }
  if (op.pair) SUNET_CUDA(launch_pdl(gemm_tn_f16_pair_kernel, dim3(op.grid), dim3(GEMM_THREADS), op.smem, stream, op.tmA0, op.tmA1, op.tmW, op.epi));
  else SUNET_CUDA(launch_pdl(gemm_tn_f16_kernel, dim3(op.grid), dim3(GEMM_THREADS), op.smem, stream, op.tmA0, op.tmA1, op.tmW, op.epi));
  return 0;
}

// ------------------------------------------------------------------------------------------ bring-up self test
__global__ void __launch_bounds__(128) umma_selftest_kernel(const __half* A, const __half* B, float* D, int N) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + A_TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__half*>(sa + sw128_offset(r, k)) = A[i];
  }
  for (int i = threadIdx.x; i < N * 64; i += 128) {
    const int r = i >> 6, k = i & 63;
    *reinterpret_cast<__half*>(sb + sw128_offset(r, k)) = B[i];
  }
  fence_proxy_async_smem();
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)N) tmem_cols <<= 1;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, N);
    const uint64_t ad = umma_desc_sw128(smem_u32(sa)), bd = umma_desc_sw128(smem_u32(sb));
    for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[row * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int umma_selftest(const __half* A, const __half* B, float* D, int N, cudaStream_t stream) {
  if (N % 16 || N < 16 || N > 256) return fail(SUNET_E_SHAPE, "selftest: bad N");
  const int smem = A_TILE_BYTES + N * 128 + 1024;
  SUNET_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, stream>>>(A, B, D, N);
  SUNET_CHECK_LAUNCH();
  return 0;
}

}  // namespace sunet
