// Window-attention core on the 5th-generation tensor cores: per (window, head)  S = Q K^T + relative-position bias,
// softmax over the 64 keys, O = P V  (SUNet_detail.py:118-135) with BOTH contractions on tcgen05 and S / O in TMEM.
// Built for head_dim 96 (SUNet stage 3: C = 768, 8 heads), where K = 96 fills six k-steps of the instruction and the
// token grid is a single 8x8 window per image: no cyclic shift (:186-189), no mask, window order == image order, so the
// operands are plain 2-D boxes of the [rows][3C] qkv matrix and arrive by TMA.
//
// Work unit = (pair of images, pair of heads): 128 token rows x 192 columns of each of q, k, v = three 128-byte-swizzled
// [128][64] k-blocks per operand (head A = columns 0..95 = block 0 + first half of block 1, head B = second half of block 1
// + block 2; every k-step of 16 stays inside one block).  The two images of a pair share one M = 128 instruction:
//   S_h  [128][128] = Q_h K_h^T           6 k-steps, N = 128: the off-diagonal 64x64 blocks (image a against image b) are
//                                          computed and ignored - the tensor pipe is idle otherwise, an M = 64 form saves nothing
//   P_h  [128][128] block-diagonal fp16   row r holds its 64 probabilities in k-block (r / 64), the other k-block stays zero
//   O_h  [128][128] = P_h V               8 k-steps; V is the B operand in MN-MAJOR form straight from the TMA tile
//                                          ([key][d] rows of 128 bytes: d contiguous, 8-key groups 1024 bytes apart = SBO,
//                                          64-column blocks one tile apart = LBO), N = 128 columns starting at block 0 (head A:
//                                          columns 0..95 valid) or block 1 (head B: columns 32..127 valid); O_h overwrites S_h in TMEM
// Warps: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2-5 = softmax / output of head A, 6-9 = head B (thread = token row:
// TMEM lane == row, the row's 64 bias values are one contiguous 256-byte line of the pre-expanded [head][64][64] table).
// The output tile is staged in the (consumed) Q buffer in the same swizzled layout and leaves through three TMA stores.
#include "attn_core_tc.cuh"

#include "device.h"
#include "error.h"
#include "gemm.cuh"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

namespace {

constexpr int TC_THREADS = 320;
constexpr int TILE = 128 * 128;          // one [128 rows][64 fp16] swizzled block
constexpr int OFF_Q = 0, OFF_K = 3 * TILE, OFF_V = 6 * TILE, OFF_P = 9 * TILE;
constexpr int TC_SMEM = 13 * TILE + 1024;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// MN-major shared-memory operand, 128-byte swizzle: rows (k index) of 64 contiguous MN elements, 8-row groups `sbo` bytes
// apart, 64-element MN blocks `lbo` bytes apart (canonical form ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo >> 4) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
constexpr uint32_t IDESC_B_MN = 1u << 16;   // instruction descriptor: B operand is MN-major

__global__ void __launch_bounds__(TC_THREADS, 1)
    attn_core_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmO,
                        const float* __restrict__ bias_exp, const int C, const int head_pairs, const int units) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qk_full[2], v_full, s_full[2], p_full[2], o_full[2], unit_done;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmO);
    mbar_init(&qk_full[0], 1);
    mbar_init(&qk_full[1], 1);
    mbar_init(&v_full, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128);
      mbar_init(&o_full[g], 1);
    }
    mbar_init(&unit_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  // the off-diagonal halves of the P tiles are never written again: zero once
  for (int i = threadIdx.x; i < 4 * TILE / 16; i += TC_THREADS) sts128(sbase + OFF_P + i * 16, make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      pdl_wait();
      uint32_t it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        if (it > 0) mbar_wait(&unit_done, (it - 1) & 1);   // the previous unit's output tile has left the Q buffer
        const int pair = u / head_pairs, hp = u - pair * head_pairs;
        const int row0 = pair * 128;
        const int cq = hp * 192, ck = C + cq, cv = 2 * C + cq;
        mbar_arrive_expect_tx(&qk_full[0], 4 * TILE);     // what head A needs: k-blocks 0 and 1 of q and k
        tma_load_2d(smem + OFF_Q, &tmQKV, &qk_full[0], cq, row0);
        tma_load_2d(smem + OFF_K, &tmQKV, &qk_full[0], ck, row0);
        tma_load_2d(smem + OFF_Q + TILE, &tmQKV, &qk_full[0], cq + 64, row0);
        tma_load_2d(smem + OFF_K + TILE, &tmQKV, &qk_full[0], ck + 64, row0);
        mbar_arrive_expect_tx(&qk_full[1], 2 * TILE);     // head B adds k-block 2
        tma_load_2d(smem + OFF_Q + 2 * TILE, &tmQKV, &qk_full[1], cq + 128, row0);
        tma_load_2d(smem + OFF_K + 2 * TILE, &tmQKV, &qk_full[1], ck + 128, row0);
        mbar_arrive_expect_tx(&v_full, 3 * TILE);
        for (int kb = 0; kb < 3; ++kb) tma_load_2d(smem + OFF_V + kb * TILE, &tmQKV, &v_full, cv + kb * 64, row0);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the warp runs the loop convergently, one elected lane issues (see gemm_tcgen05.cu)
    const uint32_t idesc_s = umma_idesc_f16(128, 128);
    const uint32_t idesc_o = idesc_s | IDESC_B_MN;
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const uint64_t q0 = umma_desc_sw128(sbase + OFF_Q), k0 = umma_desc_sw128(sbase + OFF_K);
      constexpr uint64_t T16 = TILE >> 4;   // one k-block further in the (addr >> 4) field
      mbar_wait(&qk_full[0], par);
      tc_fence_after();
      if (elect_one()) {   // S_A: columns 0..95 of the pair = k-block 0 (4 k-steps) + k-block 1 (k-steps 0, 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base, q0 + 2 * k, k0 + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_f16_ss(tmem_base, q0 + T16 + 2 * k, k0 + T16 + 2 * k, idesc_s, 1u);
        tc_commit(&s_full[0]);
      }
      __syncwarp();
      mbar_wait(&qk_full[1], par);
      tc_fence_after();
      if (elect_one()) {   // S_B: columns 96..191 = k-block 1 (k-steps 2, 3) + k-block 2
#pragma unroll
        for (int k = 2; k < 4; ++k) umma_f16_ss(tmem_base + 128, q0 + T16 + 2 * k, k0 + T16 + 2 * k, idesc_s, k > 2 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_base + 128, q0 + 2 * T16 + 2 * k, k0 + 2 * T16 + 2 * k, idesc_s, 1u);
        tc_commit(&s_full[1]);
      }
      __syncwarp();
      mbar_wait(&v_full, par);
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        mbar_wait(&p_full[g], par);   // P_g is in shared memory and every thread of the group has read S_g out of TMEM
        tc_fence_after();
        if (elect_one()) {
          const uint64_t p0 = umma_desc_sw128(sbase + OFF_P + g * 2 * TILE);
          const uint32_t vaddr = sbase + OFF_V + g * TILE;   // N = 128 columns of v from k-block g on
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // 16 keys per step: P k-block ks / 4, V rows 16 ks .. 16 ks + 15
            umma_f16_ss(tmem_base + g * 128, p0 + (ks >> 2) * T16 + 2 * (ks & 3), umma_desc_mn_sw128(vaddr + ks * 2048, TILE, 1024),
                        idesc_o, ks > 0 ? 1u : 0u);
          tc_commit(&o_full[g]);
        }
        __syncwarp();
      }
    }
  } else {
    // softmax + output: group g = head g of the pair, thread = token row (TMEM lane)
    pdl_wait();
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int img = r >> 6, i = r & 63;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t sw = static_cast<uint32_t>(r & 7);
    uint32_t it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      const int pair = u / head_pairs, hp = u - pair * head_pairs;
      const int head = hp * 2 + g;
      float s[64];
      {
        const float4* brow = reinterpret_cast<const float4*>(bias_exp + (static_cast<size_t>(head) * 64 + i) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 b = __ldg(brow + j);
          s[4 * j] = b.x; s[4 * j + 1] = b.y; s[4 * j + 2] = b.z; s[4 * j + 3] = b.w;
        }
      }
      mbar_wait(&s_full[g], par);
      tc_fence_after();
      {
        uint32_t v0[32], v1[32];
        const uint32_t ta = lane_addr + g * 128 + img * 64;   // this image's diagonal block
        tmem_ld32(ta, v0);
        tmem_ld32(ta + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[j] += __uint_as_float(v0[j]);
          s[32 + j] += __uint_as_float(v1[j]);
        }
      }
      float m = s[0];
#pragma unroll
      for (int j = 1; j < 64; ++j) m = fmaxf(m, s[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        s[j] = ex2f(s[j] - m);
        sum += s[j];
      }
      {
        const uint32_t prow = sbase + OFF_P + (g * 2 + img) * TILE + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          sts128(prow + ((static_cast<uint32_t>(c) ^ sw) << 4),
                 make_uint4(pack2(s[8 * c], s[8 * c + 1]), pack2(s[8 * c + 2], s[8 * c + 3]), pack2(s[8 * c + 4], s[8 * c + 5]),
                            pack2(s[8 * c + 6], s[8 * c + 7])));
      }
      tc_fence_before();          // the S reads above are complete: O may overwrite the columns
      fence_proxy_async_smem();   // P row visible to the tensor core's operand reads
      mbar_arrive(&p_full[g]);
      float inv;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sum));
      mbar_wait(&o_full[g], par);
      tc_fence_after();
      // O_g: 96 valid columns from column 0 (head A) or 32 (head B); they are columns 96 g .. 96 g + 95 of the unit's output tile,
      // i.e. 16-byte chunks 12 g .. 12 g + 11 of the three staged k-blocks
#pragma unroll
      for (int part = 0; part < 3; ++part) {
        uint32_t v[32];
        tmem_ld32(lane_addr + g * 128 + g * 32 + part * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int cg = g * 12 + part * 4 + c;
          const uint32_t dst = sbase + OFF_Q + (cg >> 3) * TILE + r * 128 + ((static_cast<uint32_t>(cg & 7) ^ sw) << 4);
          sts128(dst, make_uint4(pack2(__uint_as_float(v[8 * c]) * inv, __uint_as_float(v[8 * c + 1]) * inv),
                                 pack2(__uint_as_float(v[8 * c + 2]) * inv, __uint_as_float(v[8 * c + 3]) * inv),
                                 pack2(__uint_as_float(v[8 * c + 4]) * inv, __uint_as_float(v[8 * c + 5]) * inv),
                                 pack2(__uint_as_float(v[8 * c + 6]) * inv, __uint_as_float(v[8 * c + 7]) * inv)));
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, 256);     // both heads' rows are staged
      if (warp == 2 && lane == 0) {
        for (int kb = 0; kb < 3; ++kb) tma_store_2d(&tmO, smem + OFF_Q + kb * TILE, hp * 192 + kb * 64, pair * 128);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&unit_done);
      }
    }
    if (warp == 2 && lane == 0) tma_store_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

__global__ void expand_bias_kernel(const float* __restrict__ table, int heads, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * 4096) return;
  const int h = idx >> 12, i = (idx >> 6) & 63, j = idx & 63;
  // relative_position_index of an 8x8 window (SUNet_detail.py:96-105): (dy + 7) * 15 + (dx + 7)
  const int rel = ((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7);
  out[idx] = __ldg(table + rel * heads + h) * LOG2E;
}

__global__ void __launch_bounds__(128) umma_mn_selftest_kernel(const __half* A, const __half* Bt, float* D, uint32_t lbo, uint32_t sbo,
                                                               uint32_t desc_lbo, uint32_t desc_sbo) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;
  uint8_t* sb = smem + TILE;   // two [64 k][64 n] blocks, 8 KB each, stored `lbo` bytes apart
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 128 * 64; i += 128) *reinterpret_cast<__half*>(sa + sw128_offset(i >> 6, i & 63)) = A[i];
  for (int i = threadIdx.x; i < 64 * 128; i += 128) {
    const int k = i >> 7, n = i & 127;
    *reinterpret_cast<__half*>(sb + (n >> 6) * lbo + (k >> 3) * sbo + sw128_offset(k & 7, n & 63)) = Bt[i];
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 128) | IDESC_B_MN;
    const uint64_t ad = umma_desc_sw128(smem_u32(sa));
    for (int k = 0; k < 4; ++k)
      umma_f16_ss(tmem_base, ad + 2 * k, umma_desc_mn_sw128(smem_u32(sb) + k * 2 * sbo, desc_lbo, desc_sbo), idesc, k > 0);
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < 128; c += 16) {
    uint32_t v[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[row * 128 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace

bool attn_core_tc_supported(int C, int heads, int H, int W, int shift) {
  return heads > 0 && heads % 2 == 0 && C == heads * 96 && H == 8 && W == 8 && shift == 0;
}

int attn_core_tc_expand_bias(const float* table, int heads, float* bias_exp, cudaStream_t stream) {
  const int n = heads * 4096;
  expand_bias_kernel<<<(n + 255) / 256, 256, 0, stream>>>(table, heads, bias_exp);
  SUNET_CHECK_LAUNCH();
  return 0;
}

int attn_core_tc_launch(const __half* qkv, int64_t ld, __half* out, int64_t ldo, int64_t rows, int C, int heads, const float* bias_exp,
                        cudaStream_t stream) {
  if (heads <= 0 || heads % 2 || C != heads * 96) return fail(SUNET_E_SHAPE, "attn (tcgen05 core): C=%d heads=%d, needs head_dim 96 and an even head count", C, heads);
  if (rows <= 0 || rows % 64 || rows > 0x7fffffff) return fail(SUNET_E_SHAPE, "attn (tcgen05 core): %lld rows, needs whole 64-token windows", (long long)rows);
  if (!bias_exp) return fail(SUNET_E_ARG, "attn (tcgen05 core): expanded bias missing");
  static DeviceOnce once;
  if (once.need()) {
    SUNET_CUDA(cudaFuncSetAttribute(attn_core_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    once.done();
  }
  alignas(64) CUtensorMap tmQKV, tmO;
  SUNET_TRY(make_tmap_2d_f16(&tmQKV, qkv, static_cast<uint64_t>(3) * C, static_cast<uint64_t>(rows), static_cast<uint64_t>(ld), 128));
  SUNET_TRY(make_tmap_2d_f16(&tmO, out, static_cast<uint64_t>(C), static_cast<uint64_t>(rows), static_cast<uint64_t>(ldo), 128));
  const int head_pairs = heads / 2;
  const int64_t units = (rows + 127) / 128 * head_pairs;
  if (units > 0x7fffffff) return fail(SUNET_E_SHAPE, "attn (tcgen05 core): too many units");
  const int sms = device_sms();
  const unsigned grid = static_cast<unsigned>(units < sms ? units : sms);
  SUNET_CUDA(launch_pdl(attn_core_tc_kernel, dim3(grid), dim3(TC_THREADS), TC_SMEM, stream, tmQKV, tmO, bias_exp, C, head_pairs,
                        static_cast<int>(units)));
  return 0;
}

int umma_mn_selftest(const __half* A, const __half* Bt, float* D, uint32_t lbo, uint32_t sbo, uint32_t desc_lbo, uint32_t desc_sbo,
                     cudaStream_t stream) {
  if (lbo % 1024 || sbo % 1024 || lbo == 0 || sbo == 0) return fail(SUNET_E_ARG, "mn selftest: offsets must be multiples of 1024");
  const int smem = TILE + static_cast<int>(lbo + 8 * sbo) + 2048;   // block n at n * lbo, 8-row group j at j * sbo, 1024 bytes each
  if (smem > 200 * 1024) return fail(SUNET_E_ARG, "mn selftest: layout too large");
  SUNET_CUDA(cudaFuncSetAttribute(umma_mn_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_mn_selftest_kernel<<<1, 128, smem, stream>>>(A, Bt, D, lbo, sbo, desc_lbo, desc_sbo);
  SUNET_CHECK_LAUNCH();
  return 0;
}

}  // namespace sunet
