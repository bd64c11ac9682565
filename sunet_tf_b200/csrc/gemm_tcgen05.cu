// Warp-specialised tcgen05 GEMM (sm_100a): TMA (128B swizzle) -> smem ring -> tcgen05.mma (fp32 accum in TMEM)
// -> tcgen05.ld epilogue (bias / GELU / PReLU / residual) -> global.
//   warp 0 : TMA producer (one elected lane)
//   warp 1 : TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5 : epilogue, warp (w & 3) owns TMEM lanes [32*(w&3), +32) == tile rows
// One 128 x block_n output tile per CTA; several CTAs co-reside per SM when the ring is short (small K),
// which overlaps one tile's epilogue with another tile's loads/MMAs.
#include "error.h"
#include "gemm.cuh"
#include "ptx.cuh"

namespace sunet {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;  // 64 fp16 = one 128-byte swizzle row
static constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
static constexpr int MAX_STAGES = 8;
static constexpr int GEMM_THREADS = 192;
static constexpr int GEMM_MAX_DYN_SMEM = 226 * 1024;  // 227 KB per CTA minus the static barriers

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__global__ void __launch_bounds__(GEMM_THREADS)
    gemm_tn_f16_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                       const __grid_constant__ CUtensorMap tmW, const GemmEpi p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[MAX_STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int block_n = p.block_n;
  const int stages = p.stages;
  const uint32_t stage_bytes = A_TILE_BYTES + block_n * BLOCK_K * 2;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int n0 = n_tile * block_n;
  const int64_t m0 = static_cast<int64_t>(m_tile) * BLOCK_M;
  const int kb0 = (p.K0 + BLOCK_K - 1) / BLOCK_K;
  const int kb1 = (p.K1 + BLOCK_K - 1) / BLOCK_K;
  const int num_kb = kb0 + kb1;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(block_n)) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    if (p.K1 > 0) tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&tmem_full_bar, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * stage_bytes;
        uint8_t* sb = sa + A_TILE_BYTES;
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        if (it < kb0) {
          tma_load_2d(sa, &tmA0, &full_bar[s], it * BLOCK_K, static_cast<int>(m0));
          tma_load_2d(sb, &tmW, &full_bar[s], it * BLOCK_K, n0);
        } else {
          const int j = it - kb0;
          tma_load_2d(sa, &tmA1, &full_bar[s], j * BLOCK_K, static_cast<int>(m0));
          tma_load_2d(sb, &tmW, &full_bar[s], p.K0 + j * BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BLOCK_M, block_n);
      for (int it = 0; it < num_kb; ++it) {
        const int s = it % stages;
        const uint32_t ph = (it / stages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * stage_bytes);
        const uint32_t sb = sa + A_TILE_BYTES;
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sb);
        // valid K elements in this block (a K0/K1 tail shorter than 64 is zero-filled by TMA but not multiplied)
        int kvalid;
        if (it < kb0) kvalid = min(BLOCK_K, p.K0 - it * BLOCK_K);
        else kvalid = min(BLOCK_K, p.K1 - (it - kb0) * BLOCK_K);
        const int ksteps = (kvalid + 15) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          // advance 16 fp16 = 32 bytes along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_f16_ss(tmem_base, adesc + static_cast<uint64_t>(2 * k), bdesc + static_cast<uint64_t>(2 * k), idesc,
                      (it > 0 || k > 0) ? 1u : 0u);
        }
        tc_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
      }
      tc_commit(&tmem_full_bar);   // accumulator complete
    }
  } else {
    // ---------------- epilogue: thread <-> one row of the tile
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int64_t m = m0 + row;
    const bool row_ok = m < p.M;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float slope = 0.f;
    if (p.act == ACT_PRELU) slope = __ldg(p.prelu);
    mbar_wait(&tmem_full_bar, 0);
    tc_fence_after();
    for (int c = 0; c < block_n; c += 16) {
      uint32_t v[16];
      tmem_ld16(taddr + c, v);
      tmem_ld_wait();
      if (row_ok) {
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
        const int n = n0 + c;
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
            f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
          }
        }
        if (p.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = gelu_erf(f[j]);
        } else if (p.act == ACT_PRELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = f[j] >= 0.f ? f[j] : slope * f[j];
        }
        if (p.R != nullptr) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.R + m * p.ldr + n);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 r = __ldg(rp + h);
            const __half2* r2 = reinterpret_cast<const __half2*>(&r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 t = __half22float2(r2[j]);
              f[h * 8 + 2 * j] += t.x;
              f[h * 8 + 2 * j + 1] += t.y;
            }
          }
        }
        if (p.out_f32) {
          float4* cp = reinterpret_cast<float4*>(static_cast<float*>(p.C) + m * p.ldc + n);
#pragma unroll
          for (int j = 0; j < 4; ++j) cp[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          uint4* cp = reinterpret_cast<uint4*>(static_cast<__half*>(p.C) + m * p.ldc + n);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 o;
            __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(f[h * 8 + 2 * j], f[h * 8 + 2 * j + 1]);
            cp[h] = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
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

static int pick_block_n(int N, int64_t m_tiles, int K) {
  static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
  // largest tile that still yields >= 2 waves of 148 SMs; otherwise the largest tile >= 96 (smem-read bound
  // below that for SS-mode MMAs), otherwise whatever divides N.
  for (int c : cand)
    if (N % c == 0 && m_tiles * (N / c) >= 296) return c;
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
  const int64_t m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  int bn = a.force_block_n ? a.force_block_n : pick_block_n(a.N, m_tiles, a.K0 + a.K1);
  if (bn == 0 || a.N % bn != 0 || bn % 16 != 0 || bn > 256) return fail(SUNET_E_SHAPE, "gemm: no tile width for N=%d (bn=%d)", a.N, bn);
  const int n_tiles = a.N / bn;
  if (m_tiles * n_tiles > 0x7fffffffLL) return fail(SUNET_E_SHAPE, "gemm: grid too large");
  const int num_kb = (a.K0 + BLOCK_K - 1) / BLOCK_K + (a.K1 + BLOCK_K - 1) / BLOCK_K;
  const int stage_bytes = A_TILE_BYTES + bn * BLOCK_K * 2;
  int stages = (GEMM_MAX_DYN_SMEM - 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > num_kb) stages = num_kb;
  if (stages < 1) stages = 1;
  SUNET_TRY(make_tmap_2d_f16(&op->tmA0, a.A0, a.K0, a.M, a.lda0, BLOCK_M));
  if (a.K1 > 0) SUNET_TRY(make_tmap_2d_f16(&op->tmA1, a.A1, a.K1, a.M, a.lda1, BLOCK_M));
  else op->tmA1 = op->tmA0;
  SUNET_TRY(make_tmap_2d_f16(&op->tmW, a.W, a.K0 + a.K1, a.N, a.ldw, bn));
  GemmEpi& e = op->epi;
  e.M = a.M; e.N = a.N; e.K0 = a.K0; e.K1 = a.K1; e.block_n = bn; e.stages = stages;
  e.bias = a.bias; e.prelu = a.prelu; e.act = a.act; e.R = a.R; e.ldr = a.ldr; e.C = a.C; e.ldc = a.ldc;
  e.out_f32 = a.out_f32; e.n_tiles = n_tiles;
  if (a.act == ACT_PRELU && a.prelu == nullptr) return fail(SUNET_E_ARG, "gemm: PReLU needs a slope pointer");
  op->grid = static_cast<unsigned>(m_tiles * n_tiles);
  op->smem = stages * stage_bytes + 1024;
  op->flops = 2.0 * (double)a.M * a.N * (a.K0 + a.K1);
  return 0;
}

int gemm_launch(const GemmOp& op, cudaStream_t stream) {
  static bool configured = false;  // per process; one device per process (one rank per GPU)
  if (!configured) {
    SUNET_CUDA(cudaFuncSetAttribute(gemm_tn_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_MAX_DYN_SMEM));
    configured = true;
  }
  gemm_tn_f16_kernel<<<op.grid, GEMM_THREADS, op.smem, stream>>>(op.tmA0, op.tmA1, op.tmW, op.epi);
  SUNET_CHECK_LAUNCH();
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
