// Bandwidth-bound kernels: casts, LayerNorm, PatchMerging gather, folded PatchEmbed conv, Dual up-sample combine,
// folded tail stencil, and the one-off weight pre-pack / fold helpers.  All loads/stores are 8- or 16-byte vectors,
// consecutive lanes touch consecutive addresses.
#include "elementwise.cuh"

#include <stdlib.h>

#include "device.h"
#include "error.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace sunet {

static inline unsigned blocks_for(int64_t n, int per_block) { return static_cast<unsigned>((n + per_block - 1) / per_block); }

// ------------------------------------------------------------------------------------------------ casts
__global__ void cast_f32_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(in + i + 4));
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
    o2[0] = __floats2half2_rn(a.x, a.y); o2[1] = __floats2half2_rn(a.z, a.w);
    o2[2] = __floats2half2_rn(b.x, b.y); o2[3] = __floats2half2_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(out + i) = o;
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = __float2half_rn(in[j]);
  }
}
__global__ void cast_f16_f32_kernel(const __half* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + i));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
    *reinterpret_cast<float4*>(out + i) = make_float4(a.x, a.y, b.x, b.y);
    *reinterpret_cast<float4*>(out + i + 4) = make_float4(c.x, c.y, d.x, d.y);
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = __half2float(in[j]);
  }
}
int cast_f32_to_f16(const float* in, __half* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(SUNET_E_ALIGN, "cast: 16-byte alignment required");
  cast_f32_f16_kernel<<<blocks_for((n + 7) / 8, 256), 256, 0, s>>>(in, out, n);
  SUNET_CHECK_LAUNCH();
  return 0;
}
int cast_f16_to_f32(const __half* in, float* out, int64_t n, cudaStream_t s) {
  if (n <= 0) return 0;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(SUNET_E_ALIGN, "cast: 16-byte alignment required");
  cast_f16_f32_kernel<<<blocks_for((n + 7) / 8, 256), 256, 0, s>>>(in, out, n);
  SUNET_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// LPR lanes cooperate on one row; each lane owns up to MAXV 16-byte vectors (8 fp16).  MERGE: the logical row of 4C
// channels is the 2x2 gather of PatchMerging.
template <int LPR, int MAXV, bool MERGE>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ in, int64_t ld_in, __half* __restrict__ out,
                                                        int64_t ld_out, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int64_t M, int C, int H, int W, int Csrc) {
  pdl_wait();
  pdl_launch_dependents();
  const int sub = threadIdx.x % LPR;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) / LPR;
  const bool active = row < M;
  const int nvec = C >> 3;
  const __half* src_row = nullptr;
  int64_t mb = 0, mh = 0, mw = 0;
  if (active) {
    if (MERGE) {
      // (32-bit arithmetic: the callers bound M below 2^31; a 64-bit division is ~150 emulated instructions)
      const unsigned W2 = static_cast<unsigned>(W >> 1), H2 = static_cast<unsigned>(H >> 1), r32 = static_cast<unsigned>(row);
      // power-of-two grids (every SUNet stage) take shifts: the kernel is issue-bound and a 32-bit division is ~20 instructions
      const bool p2 = ((W2 & (W2 - 1)) | (H2 & (H2 - 1))) == 0;
      const unsigned rw = p2 ? r32 >> (__ffs(W2) - 1) : r32 / W2, bb = p2 ? rw >> (__ffs(H2) - 1) : rw / H2;
      mw = r32 - rw * W2; mh = rw - bb * H2; mb = bb;
    } else {
      src_row = in + row * ld_in;
    }
  }
  float x[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = sub + i * LPR;
    if (active && v < nvec) {
      const __half* p;
      if (MERGE) {
        const int col = v << 3;
        const int q = (col >= Csrc) + (col >= 2 * Csrc) + (col >= 3 * Csrc), off = col - q * Csrc;   // segment order TL, BL, TR, BR (SUNet_detail.py:312-316); col < 4 Csrc
        const int64_t y = 2 * mh + (q & 1), xx = 2 * mw + (q >> 1);
        p = in + ((mb * H + y) * W + xx) * Csrc + off;
      } else {
        p = src_row + (v << 3);
      }
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        x[i][2 * j] = f.x; x[i][2 * j + 1] = f.y;
        sum += f.x + f.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[i][j] = 0.f;
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = x[i][j] - mean; sq += d * d; }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + 1e-5f);
  if (!active) return;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = sub + i * LPR;
    if (v < nvec) {
      const int col = v << 3;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col)), b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 o;
      __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        o2[j] = __floats2half2_rn((x[i][2 * j] - mean) * rstd * g[2 * j] + b[2 * j], (x[i][2 * j + 1] - mean) * rstd * g[2 * j + 1] + b[2 * j + 1]);
      *reinterpret_cast<uint4*>(out + row * ld_out + col) = o;
    }
  }
}

#ifndef SUNET_LN_THIRDS
#define SUNET_LN_THIRDS 1   // 0: only the power-of-two lane / vector table (A/B runs)
#endif
template <bool MERGE>
static int launch_ln(const __half* in, int64_t ld_in, __half* out, int64_t ld_out, const float* gamma, const float* beta, int64_t M,
                     int C, int H, int W, int Csrc, cudaStream_t s) {
  if (C % 8 != 0 || C > 32 * 8 * 8) return fail(SUNET_E_SHAPE, "layernorm: C=%d must be a multiple of 8 and <= 2048", C);
  if ((ld_in % 8) || (ld_out % 8) || (reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(gamma) & 15) || (reinterpret_cast<uintptr_t>(beta) & 15))
    return fail(SUNET_E_ALIGN, "layernorm: 16-byte aligned rows / affine vectors required");
  const int nvec = C / 8;
  // Row widths of 3 * 2^k vectors (C = 96, 192, 384, 768 and the 4C rows of PatchMerging) take three vectors per lane on 2^k
  // lanes: every lane is active (the power-of-two table below leaves a quarter of them idle), the reductions are two steps shorter
  // and each thread has three independent 16-byte loads in flight.
#define LN_CASE(LPR, MAXV)                                                                                                          \
  SUNET_CUDA(launch_pdl(layernorm_kernel<LPR, MAXV, MERGE>, dim3(blocks_for(M * LPR, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, gamma, beta, \
                        M, C, H, W, Csrc))
  // (192 vectors - the stage-2 merge, 4096 rows - stay on <32, 8>: six vectors per lane measured 18.9 against 16.9 us there)
  if (SUNET_LN_THIRDS && (nvec == 12 || nvec == 24 || nvec == 48 || nvec == 96)) {
    switch (nvec) {
      case 12: LN_CASE(4, 3); break;
      case 24: LN_CASE(8, 3); break;
      case 48: LN_CASE(16, 3); break;
      default: LN_CASE(32, 3); break;
    }
    SUNET_CHECK_LAUNCH();
    return 0;
  }
#undef LN_CASE
  if (nvec <= 16) {
    SUNET_CUDA(launch_pdl(layernorm_kernel<16, 1, MERGE>, dim3(blocks_for(M * 16, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, gamma, beta, M, C, H, W, Csrc));
  } else if (nvec <= 32) {
    SUNET_CUDA(launch_pdl(layernorm_kernel<32, 1, MERGE>, dim3(blocks_for(M * 32, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, gamma, beta, M, C, H, W, Csrc));
  } else if (nvec <= 64) {
    SUNET_CUDA(launch_pdl(layernorm_kernel<32, 2, MERGE>, dim3(blocks_for(M * 32, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, gamma, beta, M, C, H, W, Csrc));
  } else if (nvec <= 128) {
    SUNET_CUDA(launch_pdl(layernorm_kernel<32, 4, MERGE>, dim3(blocks_for(M * 32, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, gamma, beta, M, C, H, W, Csrc));
  } else {
    SUNET_CUDA(launch_pdl(layernorm_kernel<32, 8, MERGE>, dim3(blocks_for(M * 32, 256)), dim3(256), 0, s, in, ld_in, out, ld_out, gamma, beta, M, C, H, W, Csrc));
  }
  SUNET_CHECK_LAUNCH();
  return 0;
}

int layernorm_f16(const __half* in, int64_t ld_in, __half* out, int64_t ld_out, const float* gamma, const float* beta, int64_t M,
                  int C, cudaStream_t s) {
  if (M <= 0) return 0;
  return launch_ln<false>(in, ld_in, out, ld_out, gamma, beta, M, C, 0, 0, C, s);
}

int merge_gather_ln_f16(const __half* in, __half* out, const float* gamma, const float* beta, int B, int H, int W, int C,
                        cudaStream_t s) {
  if ((H & 1) || (W & 1)) return fail(SUNET_E_SHAPE, "patch merging: grid %dx%d must be even", H, W);
  if (C % 8) return fail(SUNET_E_SHAPE, "patch merging: C=%d must be a multiple of 8", C);
  const int64_t M = static_cast<int64_t>(B) * (H / 2) * (W / 2);
  if (M >= (static_cast<int64_t>(1) << 31)) return fail(SUNET_E_SHAPE, "patch merging: %lld rows exceed the 32-bit row arithmetic", (long long)M);
  return launch_ln<true>(in, 0, out, 4 * C, gamma, beta, M, 4 * C, H, W, C, s);
}

// ------------------------------------------------------------------------------------------------ folded PatchEmbed
// CTA = 8x8 tokens; warp = one token row (8 tokens); lane = EPL output channels {lane, lane+32, ...}.
// Per tap: EPL conflict-free weight reads + 8 broadcast input reads feed 8*EPL FMAs.
template <int EPL, bool U8>
__global__ void __launch_bounds__(256) patch_embed_kernel(const void* __restrict__ img_v, int img_chans, int Himg, int Wimg,
                                                          const float* __restrict__ wfold, const float* __restrict__ bfold,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          __half* __restrict__ out) {
  pdl_wait();   // (the first kernel of a forward: orders it after the previous forward's readers of `out`)
  pdl_launch_dependents();
  constexpr int E = EPL * 32;
  constexpr int PITCH = 35;
  extern __shared__ float sm[];
  float* s_in = sm;                      // [3][34][PITCH]
  float* s_w = sm + 3 * 34 * PITCH;      // [108][E]
  const int GW = Wimg >> 2, GH = Himg >> 2;
  const int tiles_x = GW >> 3, tiles_y = GH >> 3;
  const int b = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x % (tiles_x * tiles_y);
  const int ty0 = (trem / tiles_x) * 8, tx0 = (trem % tiles_x) * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 108 * E; i += 256) s_w[i] = __ldg(wfold + i);
  const int y_base = 4 * ty0 - 1, x_base = 4 * tx0 - 1;
  for (int i = tid; i < 3 * 34 * 34; i += 256) {
    const int c = i / (34 * 34), r = (i / 34) % 34, col = i % 34;
    const int y = y_base + r, x = x_base + col;
    float v = 0.f;
    if (y >= 0 && y < Himg && x >= 0 && x < Wimg) {
      const int cs = img_chans == 1 ? 0 : c;   // grey input is repeated to 3 channels (model/SUNet.py:27-28)
      if (U8)   // PIL layout (B, H, W, chans); to_tensor's division by 255 (demo.py:71)
        v = static_cast<float>(__ldg(static_cast<const uint8_t*>(img_v) + ((static_cast<int64_t>(b) * Himg + y) * Wimg + x) * img_chans + cs)) / 255.f;
      else
        v = __ldg(static_cast<const float*>(img_v) + ((static_cast<int64_t>(b) * img_chans + cs) * Himg + y) * Wimg + x);
    }
    s_in[(c * 34 + r) * PITCH + col] = v;
  }
  __syncthreads();
  float acc[8][EPL];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[t][e] = 0.f;
  for (int c = 0; c < 3; ++c) {
    for (int u = 0; u < 6; ++u) {
      const float* in_row = s_in + (c * 34 + 4 * warp + u) * PITCH;
#pragma unroll
      for (int v = 0; v < 6; ++v) {
        const float* wk = s_w + (c * 36 + u * 6 + v) * E + lane;
        float w[EPL];
#pragma unroll
        for (int e = 0; e < EPL; ++e) w[e] = wk[32 * e];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float xv = in_row[4 * t + v];
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[t][e] = fmaf(xv, w[e], acc[t][e]);
        }
      }
    }
  }
  float bia[EPL], gam[EPL], bet[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    bia[e] = __ldg(bfold + lane + 32 * e);
    gam[e] = __ldg(gamma + lane + 32 * e);
    bet[e] = __ldg(beta + lane + 32 * e);
  }
  const int64_t tok_row0 = (static_cast<int64_t>(b) * GH + ty0 + warp) * GW + tx0;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) { acc[t][e] += bia[e]; sum += acc[t][e]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / E;
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) { const float d = acc[t][e] - mean; sq += d * d; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / E + 1e-5f);
    __half* orow = out + (tok_row0 + t) * E;
#pragma unroll
    for (int e = 0; e < EPL; ++e) orow[lane + 32 * e] = __float2half_rn((acc[t][e] - mean) * rstd * gam[e] + bet[e]);
  }
}

// Tensor-core form of the same folded conv (used when the token grid splits into 8 x 16 tiles): the 6x6x3 patch of a token is
// one row of an implicit [tokens][108 -> 112] fp16 matrix that is never built - the A fragments of mma.sync.m16n8k16 are read
// straight from the staged fp16 image tile (the k order (c, u, v) keeps every (k, k+1) pair adjacent in an image row, so a
// fragment register is one 32-bit shared load), B fragments by ldmatrix from the [E][112] fp16 weights, fp32 accumulators,
// bias + LayerNorm in registers, rows staged per warp so that the 16 tokens of a warp leave as one contiguous 16*E*2-byte run.
// CTA = 8 x 16 tokens (34 x 66 x 3 pixels); warp = one token row = one m16 tile.  Persistent over the tiles, 2 CTAs per SM.
__device__ __forceinline__ void pe_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void pe_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#ifndef SUNET_PE_VEC
#define SUNET_PE_VEC 1   // 16-byte image staging in patch_embed_mma_kernel (0: the 4-byte per-pixel form, for A/B runs)
#endif
constexpr int PE_KP = 112;   // 108 taps padded to 7 k-steps
constexpr int PE_WP = 120;   // weight row pitch in halves (240 B: the 8 rows of an ldmatrix land in 8 different 16-byte columns)
constexpr int PE_IP = 72;    // image row pitch in floats
constexpr int PE_IMG = 3 * 34 * PE_IP;   // floats per staged image tile
static_assert(PE_WP == PATCH_EMBED_WPK_PITCH, "pack_patch_embed_f16 layout");
template <int EPL> constexpr int pe_mma_smem() { return EPL * 32 * PE_WP * 2 + 2 * PE_IMG * 4 + 8 * 16 * (EPL * 32 + 8) * 2 + 3 * EPL * 32 * 4; }

__device__ __forceinline__ void pe_cp_async4(uint32_t dst, const void* src, int src_bytes) {   // src_bytes 0: zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// The pixels of one 8 x 16-token tile (3 x 34 x 66, zero outside the image) into a staging buffer.  fp32 input goes through cp.async
// (in flight while the previous tile is computed); 8-bit input is loaded, scaled by 1/255 (TF.to_tensor, demo.py:71) and stored.
// VEC (fp32 input, 16-byte aligned image rows): the tile is staged from the aligned column x_base - 3 in 16-byte chunks - 3 x 34 x 18
// cp.async instead of 3 x 34 x 66 four-byte ones, with 32-bit index arithmetic inside the image (the per-pixel form spent 58% of the
// kernel's instructions on this staging loop, ncu source page r09); the consumer then finds pixel x_base at column PE_XOFF = 3.
constexpr int PE_XOFF = 3;
__device__ __forceinline__ void pe_cp_async16(uint32_t dst, const void* src, int src_bytes) {   // src_bytes 0: zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
template <bool U8, bool VEC>
__device__ __forceinline__ void pe_stage_tile(const void* __restrict__ img_v, int img_chans, int Himg, int Wimg, int b, int y_base, int x_base,
                                              float* buf, int tid) {
  constexpr int NPIX = 3 * 34 * 66;
  if (!U8 && VEC) {
    static_assert(PE_IP == 72, "18 chunks of 4 floats per staged row");
    constexpr int NCHUNK = 3 * 34 * 18;
    const float* img_b = static_cast<const float*>(img_v) + static_cast<int64_t>(b) * img_chans * Himg * Wimg;
    const uint32_t dst0 = smem_u32(buf);
    const int x_al = x_base - PE_XOFF;   // a multiple of 4 (x_base = 64 tx - 1)
    for (int i = tid; i < NCHUNK; i += 256) {
      const int c = i / (34 * 18), rem = i - c * (34 * 18), r = rem / 18, ch = rem - r * 18;
      const int y = y_base + r, x = x_al + 4 * ch;
      // (Wimg is a multiple of 4: an aligned chunk lies wholly inside or wholly outside the image)
      const bool in = y >= 0 && y < Himg && x >= 0 && x < Wimg;
      const int cs = img_chans == 1 ? 0 : c;   // grey input is repeated to 3 channels (model/SUNet.py:27-28)
      const float* src = in ? img_b + static_cast<uint32_t>((cs * Himg + y) * Wimg + x) : img_b;
      pe_cp_async16(dst0 + ((c * 34 + r) * PE_IP + 4 * ch) * 4, src, in ? 16 : 0);
    }
  } else if (!U8) {
    const float* img = static_cast<const float*>(img_v);
    const uint32_t dst0 = smem_u32(buf);
    for (int i = tid; i < NPIX; i += 256) {
      const int c = i / (34 * 66), rem = i - c * (34 * 66), r = rem / 66, col = rem - r * 66;
      const int y = y_base + r, x = x_base + col;
      const bool in = y >= 0 && y < Himg && x >= 0 && x < Wimg;
      const int cs = img_chans == 1 ? 0 : c;   // grey input is repeated to 3 channels (model/SUNet.py:27-28)
      const float* src = in ? img + ((static_cast<int64_t>(b) * img_chans + cs) * Himg + y) * Wimg + x : img;
      pe_cp_async4(dst0 + ((c * 34 + r) * PE_IP + col) * 4, src, in ? 4 : 0);
    }
  } else {
    const uint8_t* img = static_cast<const uint8_t*>(img_v);
    constexpr int BATCH = 9;
#pragma unroll 1
    for (int i0 = 0; i0 < NPIX; i0 += 256 * BATCH) {
      float v[BATCH];
      int dst[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * 256 + tid;
        const int c = i / (34 * 66), rem = i - c * (34 * 66), r = rem / 66, col = rem - r * 66;
        const int y = y_base + r, x = x_base + col;
        v[k] = 0.f;
        dst[k] = i < NPIX ? (c * 34 + r) * PE_IP + col : -1;
        if (i < NPIX && y >= 0 && y < Himg && x >= 0 && x < Wimg) {
          const int cs = img_chans == 1 ? 0 : c;
          v[k] = static_cast<float>(__ldg(img + ((static_cast<int64_t>(b) * Himg + y) * Wimg + x) * img_chans + cs));
        }
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k)
        if (dst[k] >= 0) buf[dst[k]] = v[k] / 255.f;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int EPL, bool U8, bool VEC>
__global__ void __launch_bounds__(256, 2) patch_embed_mma_kernel(const void* __restrict__ img_v, int img_chans, int Himg, int Wimg,
                                                                 const __half* __restrict__ wpk, const float* __restrict__ bfold,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                 __half* __restrict__ out, int tiles_total) {
  constexpr int E = EPL * 32, NT = E / 8, OP = E + 8;
  extern __shared__ __align__(16) uint8_t pe_smem[];
  __half* s_w = reinterpret_cast<__half*>(pe_smem);                 // [E][PE_WP]
  float* s_img = reinterpret_cast<float*>(s_w + E * PE_WP);         // [2][3][34][PE_IP] fp32, double-buffered
  __half* s_out = reinterpret_cast<__half*>(s_img + 2 * PE_IMG);    // [8][16][OP]
  float* s_par = reinterpret_cast<float*>(s_out + 8 * 16 * OP);     // bias | gamma | beta
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tq = lane & 3;
  // parameters do not depend on the preceding kernel: staged before the dependency wait
  {   // wpk is already [E][PE_WP] fp16 (pack_patch_embed_f16): a straight 16-byte copy, all loads in flight at once
    constexpr int NV = E * PE_WP / 8, IT = (NV + 255) / 256;
    uint4 w[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k)
      if (tid + k * 256 < NV) w[k] = __ldg(reinterpret_cast<const uint4*>(wpk) + tid + k * 256);
#pragma unroll
    for (int k = 0; k < IT; ++k)
      if (tid + k * 256 < NV) reinterpret_cast<uint4*>(s_w)[tid + k * 256] = w[k];
  }
  for (int i = tid; i < E; i += 256) {
    s_par[i] = __ldg(bfold + i);
    s_par[E + i] = __ldg(gamma + i);
    s_par[2 * E + i] = __ldg(beta + i);
  }
  // this lane's pair offsets into the image tile: pair j = k / 2 = (c * 6 + u) * 3 + v / 2
  int aoff[7][2];
#pragma unroll
  for (int ks = 0; ks < 7; ++ks)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int j = 8 * ks + tq + 4 * h;
      if (j >= 54) j = 0;   // zero weights there; any staged value will do
      const int cu = j / 3, vp = j - 3 * cu, c = cu / 6, u = cu - 6 * c;
      aoff[ks][h] = (c * 34 + u) * PE_IP + 2 * vp;
    }
  pdl_wait();
  pdl_launch_dependents();
  const int GW = Wimg >> 2, GH = Himg >> 2;
  const int tiles_x = GW >> 4, tiles_y = GH >> 3, tiles_img = tiles_x * tiles_y;
  const uint32_t w_addr = smem_u32(s_w) + (((lane >> 4) * 8 + (lane & 7)) * PE_WP + 8 * ((lane >> 3) & 1)) * 2;
  int buf = 0;
  if (static_cast<int>(blockIdx.x) < tiles_total) {
    const int t = blockIdx.x, b = t / tiles_img, trem = t - b * tiles_img;
    pe_stage_tile<U8, VEC>(img_v, img_chans, Himg, Wimg, b, 32 * (trem / tiles_x) - 1, 64 * (trem % tiles_x) - 1, s_img, tid);
  }
  for (int tile = blockIdx.x; tile < tiles_total; tile += gridDim.x, buf ^= 1) {
    const int b = tile / tiles_img;
    const int trem = tile - b * tiles_img;
    const int ty0 = (trem / tiles_x) * 8, tx0 = (trem % tiles_x) * 16;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // this tile's pixels have landed; every warp is done reading the other buffer
    const int next = tile + gridDim.x;
    if (next < tiles_total) {
      const int nb = next / tiles_img, nrem = next - nb * tiles_img;
      pe_stage_tile<U8, VEC>(img_v, img_chans, Himg, Wimg, nb, 32 * (nrem / tiles_x) - 1, 64 * (nrem % tiles_x) - 1, s_img + (buf ^ 1) * PE_IMG, tid);
    }
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    // token (row warp, col g); col g + 8 is 32 floats further.  VEC: the tile starts PE_XOFF columns to the left, so a k-pair sits at
    // an odd column and is read as two 32-bit loads instead of one 64-bit load
    const float* a_base = s_img + buf * PE_IMG + (4 * warp) * PE_IP + 4 * g + (VEC ? PE_XOFF : 0);
    auto ld_pair = [](const float* q) {
      if constexpr (VEC) return make_float2(q[0], q[1]);
      else return *reinterpret_cast<const float2*>(q);
    };
#pragma unroll
    for (int ks = 0; ks < 7; ++ks) {
      const float2 f0 = ld_pair(a_base + aoff[ks][0]);
      const float2 f1 = ld_pair(a_base + 32 + aoff[ks][0]);
      const float2 f2 = ld_pair(a_base + aoff[ks][1]);
      const float2 f3 = ld_pair(a_base + 32 + aoff[ks][1]);
      uint32_t a[4];
      {
        const __half2 h0 = __floats2half2_rn(f0.x, f0.y), h1 = __floats2half2_rn(f1.x, f1.y);
        const __half2 h2 = __floats2half2_rn(f2.x, f2.y), h3 = __floats2half2_rn(f3.x, f3.y);
        a[0] = *reinterpret_cast<const uint32_t*>(&h0); a[1] = *reinterpret_cast<const uint32_t*>(&h1);
        a[2] = *reinterpret_cast<const uint32_t*>(&h2); a[3] = *reinterpret_cast<const uint32_t*>(&h3);
      }
#pragma unroll
      for (int nt = 0; nt < NT; nt += 2) {
        uint32_t bf[4];   // (nt, k lo), (nt, k hi), (nt + 1, k lo), (nt + 1, k hi)
        pe_ldsm_x4(bf, w_addr + (nt * 8 * PE_WP + ks * 16) * 2);
        pe_mma(acc[nt], a, bf[0], bf[1]);
        pe_mma(acc[nt + 1], a, bf[2], bf[3]);
      }
    }
    // bias + LayerNorm: token g lives in acc[.][0..1], token g + 8 in acc[.][2..3], a row is spread over the 4 lanes of a quad
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float2 bb = *reinterpret_cast<const float2*>(s_par + 8 * nt + 2 * tq);
      acc[nt][0] += bb.x; acc[nt][1] += bb.y; acc[nt][2] += bb.x; acc[nt][3] += bb.y;
      sum0 += acc[nt][0] + acc[nt][1];
      sum1 += acc[nt][2] + acc[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float mean0 = sum0 / E, mean1 = sum1 / E;
    float sq0 = 0.f, sq1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float d0 = acc[nt][0] - mean0, d1 = acc[nt][1] - mean0, d2 = acc[nt][2] - mean1, d3 = acc[nt][3] - mean1;
      sq0 += d0 * d0 + d1 * d1;
      sq1 += d2 * d2 + d3 * d3;
    }
    sq0 += __shfl_xor_sync(0xffffffffu, sq0, 1); sq0 += __shfl_xor_sync(0xffffffffu, sq0, 2);
    sq1 += __shfl_xor_sync(0xffffffffu, sq1, 1); sq1 += __shfl_xor_sync(0xffffffffu, sq1, 2);
    const float rstd0 = rsqrtf(sq0 / E + 1e-5f), rstd1 = rsqrtf(sq1 / E + 1e-5f);
    __half* so = s_out + warp * 16 * OP;
    __syncwarp();   // the previous tile's copy-out of this warp's staging rows is complete
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float2 ga = *reinterpret_cast<const float2*>(s_par + E + 8 * nt + 2 * tq);
      const float2 be = *reinterpret_cast<const float2*>(s_par + 2 * E + 8 * nt + 2 * tq);
      *reinterpret_cast<__half2*>(so + g * OP + 8 * nt + 2 * tq) =
          __floats2half2_rn((acc[nt][0] - mean0) * rstd0 * ga.x + be.x, (acc[nt][1] - mean0) * rstd0 * ga.y + be.y);
      *reinterpret_cast<__half2*>(so + (g + 8) * OP + 8 * nt + 2 * tq) =
          __floats2half2_rn((acc[nt][2] - mean1) * rstd1 * ga.x + be.x, (acc[nt][3] - mean1) * rstd1 * ga.y + be.y);
    }
    __syncwarp();
    // the warp's 16 tokens are consecutive in the token stream: one contiguous run of 16 * E halves
    __half* orow = out + ((static_cast<int64_t>(b) * GH + ty0 + warp) * GW + tx0) * E;
    constexpr int CPR = E / 8;   // 16-byte chunks per token
    for (int idx = lane; idx < 16 * CPR; idx += 32) {
      const int row = idx / CPR, ch = idx - row * CPR;
      *reinterpret_cast<uint4*>(orow + idx * 8) = *reinterpret_cast<const uint4*>(so + row * OP + ch * 8);
    }
  }
}

int patch_embed_fused(const void* img, int img_fmt, int img_chans, int B, int Himg, int Wimg, const float* wfold, const __half* wpk,
                      const float* bfold, const float* gamma, const float* beta, int E, __half* out, cudaStream_t s) {
  if (img_chans != 1 && img_chans != 3) return fail(SUNET_E_SHAPE, "patch embed: input must have 1 or 3 channels, got %d", img_chans);
  if (Himg % 32 || Wimg % 32) return fail(SUNET_E_SHAPE, "patch embed: image %dx%d must be a multiple of 32", Himg, Wimg);
  if (E % 32 || E > 128) return fail(SUNET_E_SHAPE, "patch embed: embed_dim %d must be a multiple of 32 and <= 128", E);
  static const bool pe_ffma = getenv("SUNET_PE_FFMA") != nullptr;   // read once per process, never on the hot call
  if (wpk != nullptr && Wimg % 64 == 0 && !pe_ffma) {   // tensor-core form: 8 x 16-token tiles
    const int tiles = B * (Himg / 32) * (Wimg / 64);
    const int sms = device_sms();
    const unsigned grid = static_cast<unsigned>(tiles < 2 * sms ? tiles : 2 * sms);
#define PE_MMA_T(EPL, U8, VEC)                                                                                               \
  {                                                                                                                          \
    SUNET_CUDA(cudaFuncSetAttribute(patch_embed_mma_kernel<EPL, U8, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, pe_mma_smem<EPL>())); \
    SUNET_CUDA(launch_pdl(patch_embed_mma_kernel<EPL, U8, VEC>, dim3(grid), dim3(256), pe_mma_smem<EPL>(), s, img, img_chans, Himg, Wimg,  \
                          wpk, bfold, gamma, beta, out, tiles));                                                          \
  }
    // 16-byte staging needs 16-byte aligned image rows (Wimg % 64 == 0 holds here) and one image below 2^31 elements
    const bool vec = SUNET_PE_VEC && (reinterpret_cast<uintptr_t>(img) & 15) == 0 && static_cast<int64_t>(img_chans) * Himg * Wimg < (1LL << 31);
#define PE_MMA(EPL) if (img_fmt == IMG_U8_NHWC) PE_MMA_T(EPL, true, false) else if (vec) PE_MMA_T(EPL, false, true) else PE_MMA_T(EPL, false, false)
    switch (E / 32) {
      case 1: PE_MMA(1); break;
      case 2: PE_MMA(2); break;
      case 3: PE_MMA(3); break;
      default: PE_MMA(4); break;
    }
#undef PE_MMA
#undef PE_MMA_T
    SUNET_CHECK_LAUNCH();
    return 0;
  }
  const unsigned grid = static_cast<unsigned>(B) * (Himg / 32) * (Wimg / 32);
  const int smem = (3 * 34 * 35 + 108 * E) * 4;
#define PE_LAUNCH_T(EPL, U8)                                                                                    \
  {                                                                                                             \
    SUNET_CUDA(cudaFuncSetAttribute(patch_embed_kernel<EPL, U8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    SUNET_CUDA(launch_pdl(patch_embed_kernel<EPL, U8>, dim3(grid), dim3(256), smem, s, img, img_chans, Himg, Wimg, wfold, bfold, gamma, beta, out));   \
  }
#define PE_LAUNCH(EPL)                                                                                          \
  if (img_fmt == IMG_U8_NHWC) PE_LAUNCH_T(EPL, true) else PE_LAUNCH_T(EPL, false)
  switch (E / 32) {
    case 1: PE_LAUNCH(1); break;
    case 2: PE_LAUNCH(2); break;
    case 3: PE_LAUNCH(3); break;
    default: PE_LAUNCH(4); break;
  }
#undef PE_LAUNCH
  SUNET_CHECK_LAUNCH();
  return 0;
}

__global__ void im2col_patch_kernel(const float* __restrict__ img, int Cin, int Himg, int Wimg, int P, __half* __restrict__ out,
                                    int64_t total) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int K = Cin * P * P;
  const int64_t m = i / K;
  const int k = static_cast<int>(i - m * K);
  const int c = k / (P * P), ky = (k / P) % P, kx = k % P;
  const int GW = Wimg / P, GH = Himg / P;
  const int px = static_cast<int>(m % GW), py = static_cast<int>((m / GW) % GH);
  const int64_t b = m / (static_cast<int64_t>(GW) * GH);
  out[i] = __float2half_rn(__ldg(img + ((b * Cin + c) * Himg + py * P + ky) * Wimg + px * P + kx));
}
int im2col_patch(const float* img, int B, int Cin, int Himg, int Wimg, int P, __half* out, cudaStream_t s) {
  if (Himg % P || Wimg % P) return fail(SUNET_E_SHAPE, "im2col: image %dx%d not divisible by patch %d", Himg, Wimg, P);
  const int64_t total = static_cast<int64_t>(B) * (Himg / P) * (Wimg / P) * Cin * P * P;
  im2col_patch_kernel<<<blocks_for(total, 256), 256, 0, s>>>(img, Cin, Himg, Wimg, P, out, total);
  SUNET_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ Dual up-sample
// align_corners=False taps along one axis (nn.Upsample, SUNet_detail.py:351/:362)
__device__ __forceinline__ void bilinear_tap(int d, int r, int n, int& i0, int& i1, float& lam) {
  const float src = fmaxf((d + 0.5f) / r - 0.5f, 0.f);
  i0 = static_cast<int>(src);
  i1 = min(i0 + 1, n - 1);
  lam = src - i0;
}

// IdxT = unsigned when the element count fits 32 bits (every SUNet shape): the five divisions of the index decomposition are
// ~20 instructions apiece in 32 bits and ~150 in 64 (emulated) - the 64-bit form was issue-bound at 2.3 TB/s
template <typename IdxT>
__global__ void __launch_bounds__(256) upsample_combine_kernel(const __half* __restrict__ Yp, const __half* __restrict__ Z,
                                                               void* __restrict__ out, int out_f32, int H, int W, int Co, int r,
                                                               int64_t total) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t i64 = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i64 >= total) return;
  const IdxT i = static_cast<IdxT>(i64);
  const IdxT nv = static_cast<IdxT>(Co >> 3);
  const IdxT pix_t = i / nv;
  const int v = static_cast<int>(i - pix_t * nv);
  const IdxT OWt = static_cast<IdxT>(W * r), OHt = static_cast<IdxT>(H * r);
  const IdxT rowt = pix_t / OWt;
  const int X = static_cast<int>(pix_t - rowt * OWt);
  const IdxT bt = rowt / OHt;
  const int Y = static_cast<int>(rowt - bt * OHt);
  const int64_t pix = static_cast<int64_t>(pix_t);
  const int64_t b = static_cast<int64_t>(bt);
  const int h = Y / r, ii = Y % r, w = X / r, jj = X % r;
  const uint4 yp = __ldg(reinterpret_cast<const uint4*>(Yp + (((b * H + h) * W + w) * (r * r) + ii * r + jj) * Co + (v << 3)));
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_tap(Y, r, H, y0, y1, ly);
  bilinear_tap(X, r, W, x0, x1, lx);
  const __half* zb = Z + b * H * W * Co + (v << 3);
  const uint4 z00 = __ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(y0) * W + x0) * Co));
  const uint4 z01 = __ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(y0) * W + x1) * Co));
  const uint4 z10 = __ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(y1) * W + x0) * Co));
  const uint4 z11 = __ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(y1) * W + x1) * Co));
  const __half2 *p = reinterpret_cast<const __half2*>(&yp), *a = reinterpret_cast<const __half2*>(&z00),
                *bq = reinterpret_cast<const __half2*>(&z01), *c = reinterpret_cast<const __half2*>(&z10),
                *d = reinterpret_cast<const __half2*>(&z11);
  float f[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fp = __half22float2(p[j]), fa = __half22float2(a[j]), fb = __half22float2(bq[j]), fc = __half22float2(c[j]),
                 fd = __half22float2(d[j]);
    // same evaluation order as the separable form: rows first, then columns
    const float tx = (fa.x * (1.f - ly) + fc.x * ly), ux = (fb.x * (1.f - ly) + fd.x * ly);
    const float ty = (fa.y * (1.f - ly) + fc.y * ly), uy = (fb.y * (1.f - ly) + fd.y * ly);
    f[2 * j] = fp.x + tx * (1.f - lx) + ux * lx;
    f[2 * j + 1] = fp.y + ty * (1.f - lx) + uy * lx;
  }
  if (out_f32) {
    float* o = static_cast<float*>(out) + pix * Co + (v << 3);
    *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(static_cast<__half*>(out) + pix * Co + (v << 3)) = o;
  }
}

// The same arithmetic with the index decomposition taken from the launch geometry: grid = (pixel groups of an output row, output
// row, image), block = (16-byte vectors of a pixel, pixels), R a compile-time constant.  The flat form above spends ~55% of its
// instructions on five runtime divisions and 64-bit addresses and is issue-bound (75% of the issue slots busy at 2.6 TB/s, ncu r10).
#ifndef SUNET_UPC_ROWS
#define SUNET_UPC_ROWS 1   // 0: always the flat-index kernel (A/B runs)
#endif
template <int R>
__global__ void __launch_bounds__(256) upsample_combine_rows_kernel(const __half* __restrict__ Yp, const __half* __restrict__ Z,
                                                                    void* __restrict__ out, int out_f32, int H, int W, int Co) {
  pdl_wait();
  pdl_launch_dependents();
  const int v8 = static_cast<int>(threadIdx.x) << 3;
  const int X = blockIdx.x * blockDim.y + threadIdx.y;
  const int OW = W * R, OH = H * R;
  if (X >= OW) return;
  const int Y = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int h = Y / R, ii = Y % R, w = X / R, jj = X % R;
  const uint4 yp = __ldg(reinterpret_cast<const uint4*>(Yp + (((b * H + h) * W + w) * (R * R) + ii * R + jj) * Co + v8));
  int y0, y1, x0, x1;
  float ly, lx;
  bilinear_tap(Y, R, H, y0, y1, ly);
  bilinear_tap(X, R, W, x0, x1, lx);
  const __half* zr0 = Z + (b * H + y0) * W * Co + v8;
  const __half* zr1 = Z + (b * H + y1) * W * Co + v8;
  const uint4 z00 = __ldg(reinterpret_cast<const uint4*>(zr0 + x0 * Co));
  const uint4 z01 = __ldg(reinterpret_cast<const uint4*>(zr0 + x1 * Co));
  const uint4 z10 = __ldg(reinterpret_cast<const uint4*>(zr1 + x0 * Co));
  const uint4 z11 = __ldg(reinterpret_cast<const uint4*>(zr1 + x1 * Co));
  const __half2 *p = reinterpret_cast<const __half2*>(&yp), *a = reinterpret_cast<const __half2*>(&z00),
                *bq = reinterpret_cast<const __half2*>(&z01), *c = reinterpret_cast<const __half2*>(&z10),
                *d = reinterpret_cast<const __half2*>(&z11);
  float f[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 fp = __half22float2(p[j]), fa = __half22float2(a[j]), fb = __half22float2(bq[j]), fc = __half22float2(c[j]),
                 fd = __half22float2(d[j]);
    // the flat kernel's expressions, term for term (bit-identical results)
    const float tx = (fa.x * (1.f - ly) + fc.x * ly), ux = (fb.x * (1.f - ly) + fd.x * ly);
    const float ty = (fa.y * (1.f - ly) + fc.y * ly), uy = (fb.y * (1.f - ly) + fd.y * ly);
    f[2 * j] = fp.x + tx * (1.f - lx) + ux * lx;
    f[2 * j + 1] = fp.y + ty * (1.f - lx) + uy * lx;
  }
  const int64_t o_off = ((b * OH + Y) * OW + X) * Co + v8;
  if (out_f32) {
    float* o = static_cast<float*>(out) + o_off;
    *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    uint4 o;
    __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
    *reinterpret_cast<uint4*>(static_cast<__half*>(out) + o_off) = o;
  }
}

// r = 2, one thread per (low-res pixel, 16-byte vector): its 2 x 2 output pixels share the 3 x 3 neighbourhood of Z, so the nine
// vectors are loaded and converted once (the per-output kernels above load 16 and convert 16 for the same four outputs) and the six
// vertical interpolations are shared by the two output columns: ~100 instead of ~240 instructions per output vector in a kernel that is
// issue-bound (83% of the issue slots, ncu r10).  Per output the expressions are those of the kernels above, term for term, with the
// same operands: for output row 2h the taps are rows (h - 1, h) with weight 0.75 - or, at h = 0, rows (0, 1) with weight 0 - and for
// row 2h + 1 rows (h, min(h + 1, H - 1)) with weight 0.25 (bilinear_tap); columns alike.
#ifndef SUNET_UPC_QUAD
#define SUNET_UPC_QUAD 1   // 0: upsample_combine_rows_kernel also for r = 2 (A/B runs)
#endif
__device__ __forceinline__ void upc_unpack(const uint4& u, float (&f)[8]) {
  const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __half22float2(h2[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}
__global__ void __launch_bounds__(256) upsample_combine_quad_kernel(const __half* __restrict__ Yp, const __half* __restrict__ Z,
                                                                    void* __restrict__ out, int out_f32, int H, int W, int Co) {
  pdl_wait();
  pdl_launch_dependents();
  const int v8 = static_cast<int>(threadIdx.x) << 3;
  const int w = blockIdx.x * blockDim.y + threadIdx.y;
  if (w >= W) return;
  const int h = blockIdx.y;
  const int64_t b = blockIdx.z;
  const int OW = 2 * W, OH = 2 * H;
  int ya0, ya1, yb0, yb1, xa0, xa1, xb0, xb1;
  float lya, lyb, lxa, lxb;
  bilinear_tap(2 * h, 2, H, ya0, ya1, lya);
  bilinear_tap(2 * h + 1, 2, H, yb0, yb1, lyb);
  bilinear_tap(2 * w, 2, W, xa0, xa1, lxa);
  bilinear_tap(2 * w + 1, 2, W, xb0, xb1, lxb);
  // rows ya0 <= ya1' <= yb1 with ya1' = yb0 = h: the three distinct rows / columns of the neighbourhood
  const int rr[3] = {ya0, h, yb1}, cc[3] = {xa0, w, xb1};
  const bool top = ya1 != h, left = xa1 != w;   // h = 0 / w = 0: the second tap of the even output row / column is row / column 1 (weight 0)
  const __half* zb = Z + b * H * W * Co + v8;
  // vertical interpolation per neighbourhood column: V[ii][c] = Z[y0(ii)][c] (1 - ly(ii)) + Z[y1(ii)][c] ly(ii)
  float V[2][3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float z0[8], z1[8], z2[8];
    upc_unpack(__ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(rr[0]) * W + cc[c]) * Co)), z0);
    upc_unpack(__ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(rr[1]) * W + cc[c]) * Co)), z1);
    upc_unpack(__ldg(reinterpret_cast<const uint4*>(zb + (static_cast<int64_t>(rr[2]) * W + cc[c]) * Co)), z2);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a1 = top ? z2[e] : z1[e];
      V[0][c][e] = z0[e] * (1.f - lya) + a1 * lya;
      V[1][c][e] = z1[e] * (1.f - lyb) + z2[e] * lyb;
    }
  }
  const __half* yrow = Yp + ((b * H + h) * W + w) * 4 * Co + v8;
#pragma unroll
  for (int ii = 0; ii < 2; ++ii) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      float fp[8], f[8];
      upc_unpack(__ldg(reinterpret_cast<const uint4*>(yrow + (ii * 2 + jj) * Co)), fp);
      const float lx = jj ? lxb : lxa;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float tx = jj ? V[ii][1][e] : V[ii][0][e];
        const float ux = jj ? V[ii][2][e] : (left ? V[ii][2][e] : V[ii][1][e]);
        f[e] = fp[e] + tx * (1.f - lx) + ux * lx;
      }
      const int64_t o_off = ((b * OH + 2 * h + ii) * OW + 2 * w + jj) * Co + v8;
      if (out_f32) {
        float* o = static_cast<float*>(out) + o_off;
        *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(f[4], f[5], f[6], f[7]);
      } else {
        uint4 o;
        __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) o2[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(static_cast<__half*>(out) + o_off) = o;
      }
    }
  }
}

int upsample_combine(const __half* Yp, const __half* Z, void* out, int out_f32, int B, int H, int W, int Co, int r, cudaStream_t s) {
  if (Co % 8) return fail(SUNET_E_SHAPE, "upsample: Co=%d must be a multiple of 8", Co);
  const int nv = Co / 8;
  if (SUNET_UPC_QUAD && r == 2 && nv <= 256 && H <= 65535 && B <= 65535) {
    int px = 1;
    while (2 * px * nv <= 256 && 2 * px <= W) px *= 2;
    const dim3 block(nv, px), grid((W + px - 1) / px, H, B);
    SUNET_CUDA(launch_pdl(upsample_combine_quad_kernel, grid, block, 0, s, Yp, Z, out, out_f32, H, W, Co));
    SUNET_CHECK_LAUNCH();
    return 0;
  }
  if (SUNET_UPC_ROWS && (r == 2 || r == 4) && nv <= 256 && H * r <= 65535 && B <= 65535) {
    int px = 1;   // pixels per block, a power of two (no ragged last block on the power-of-two SUNet rows); consecutive threads
    while (2 * px * nv <= 256) px *= 2;   // cover consecutive 16-byte vectors of the output row
    const dim3 block(nv, px), grid((W * r + px - 1) / px, H * r, B);
    if (r == 2) SUNET_CUDA(launch_pdl(upsample_combine_rows_kernel<2>, grid, block, 0, s, Yp, Z, out, out_f32, H, W, Co));
    else SUNET_CUDA(launch_pdl(upsample_combine_rows_kernel<4>, grid, block, 0, s, Yp, Z, out, out_f32, H, W, Co));
    SUNET_CHECK_LAUNCH();
    return 0;
  }
  const int64_t total = static_cast<int64_t>(B) * H * r * W * r * (Co / 8);
  if (total < (static_cast<int64_t>(1) << 31))
    SUNET_CUDA(launch_pdl(upsample_combine_kernel<unsigned>, dim3(blocks_for(total, 256)), dim3(256), 0, s, Yp, Z, out, out_f32, H, W, Co, r, total));
  else
    SUNET_CUDA(launch_pdl(upsample_combine_kernel<unsigned long long>, dim3(blocks_for(total, 256)), dim3(256), 0, s, Yp, Z, out, out_f32, H, W, Co, r, total));
  SUNET_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ folded tail stencil
// MODE 0: fp32 NCHW; 1: 8-bit NHWC, rint(clamp(v, 0, 1) * 255); 2: fp32 NCHW + validation epilogue (EvalEpilogue)
//
// CTA = 8 x 8 tokens = 32 x 32 output pixels of one image.  Phase 1 stages the used taps (OC * 9 of the NT columns) of the
// 34 x 34 pixel rows of Qp around the tile and of the 10 x 10 tokens of Rb the bilinear taps can reach, as 16-byte row reads;
// phase 2 sums the 9 taps per output pixel from shared memory (tap stride OC * 9 words: conflict-free for odd OC) and stores
// whole 128-byte output rows.  (The direct form - 45 scalar global loads per pixel - ran at 1.4 TB/s of the 268 MB it reads.)
constexpr int TS_T = 8;                 // tokens per tile side
constexpr int TS_P = 4 * TS_T + 2;      // staged pixel rows / columns (one-pixel halo)
constexpr int TS_R = TS_T + 2;          // staged Rb tokens per side

template <int MODE, int OC>
__global__ void __launch_bounds__(256) tail_stencil_kernel(const float* __restrict__ Qp, const float* __restrict__ Rb,
                                                           void* __restrict__ out_v, EvalEpilogue ev, int H, int W, int NT) {
  extern __shared__ __align__(16) float ts_smem[];
  constexpr int P = OC * 9;                   // used taps per row
  float* sQ = ts_smem;                        // [TS_P][TS_P][P]
  float* sR = ts_smem + TS_P * TS_P * P;      // [TS_R][TS_R][P]
  pdl_wait();
  pdl_launch_dependents();
  const int OW = 4 * W, OH = 4 * H;
  const int tiles_x = (W + TS_T - 1) / TS_T, tiles_y = (H + TS_T - 1) / TS_T;
  const int64_t b = blockIdx.x / (tiles_x * tiles_y);
  const int trem = blockIdx.x % (tiles_x * tiles_y);
  const int th0 = (trem / tiles_x) * TS_T, tw0 = (trem % tiles_x) * TS_T;
  const int py0 = 4 * th0 - 1, px0 = 4 * tw0 - 1;    // image coordinates of staged pixel (0, 0)
  const int tid = threadIdx.x;
  constexpr int nvec = (P + 3) >> 2;
  {   // register batches: all 16-byte loads of a batch are in flight before the first shared-memory store
    constexpr int TOTAL = TS_P * TS_P * nvec, BATCH = 7;
#pragma unroll 1
    for (int i0 = 0; i0 < TOTAL; i0 += 256 * BATCH) {
      float4 q[BATCH];
      int dst[BATCH];
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        const int i = i0 + k * 256 + tid;
        const int pix = i / nvec, v = i - pix * nvec;
        const int ry = pix / TS_P, rx = pix - ry * TS_P;
        const int ny = py0 + ry, nx = px0 + rx;
        q[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        dst[k] = i < TOTAL ? pix * P + 4 * v : -1;
        if (i < TOTAL && ny >= 0 && ny < OH && nx >= 0 && nx < OW)
          q[k] = __ldg(reinterpret_cast<const float4*>(Qp + ((((b * H + (ny >> 2)) * W + (nx >> 2)) << 4) + ((ny & 3) << 2) + (nx & 3)) * NT) + v);
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k) {
        if (dst[k] < 0) continue;
        const int v4 = (dst[k] % P);   // 4 * v
        float* d = sQ + dst[k];
        const float qv[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (v4 + e < P) d[e] = qv[e];
      }
    }
  }
  for (int i = tid; i < TS_R * TS_R * nvec; i += 256) {
    const int tok = i / nvec, v = i - tok * nvec;
    const int ry = tok / TS_R, rx = tok - ry * TS_R;
    const int ty = min(max(th0 - 1 + ry, 0), H - 1), tx = min(max(tw0 - 1 + rx, 0), W - 1);
    const float4 q = __ldg(reinterpret_cast<const float4*>(Rb + ((b * H + ty) * W + tx) * NT) + v);
    float* d = sR + tok * P + 4 * v;
    const float qv[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (4 * v + k < P) d[k] = qv[k];
  }
  __syncthreads();
  double part[4] = {0.0, 0.0, 0.0, 0.0};
  const int lx_ = tid & 31;
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int ly_ = (tid >> 5) + 8 * k;
    const int y = 4 * th0 + ly_, x = 4 * tw0 + lx_;
    if (y >= OH || x >= OW) continue;
    int yy0[3], yy1[3], xx0[3], xx1[3];
    float ly[3], lx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      bilinear_tap(min(max(y + d - 1, 0), OH - 1), 4, H, yy0[d], yy1[d], ly[d]);
      bilinear_tap(min(max(x + d - 1, 0), OW - 1), 4, W, xx0[d], xx1[d], lx[d]);
      yy0[d] = min(max(yy0[d] - (th0 - 1), 0), TS_R - 1); yy1[d] = min(max(yy1[d] - (th0 - 1), 0), TS_R - 1);
      xx0[d] = min(max(xx0[d] - (tw0 - 1), 0), TS_R - 1); xx1[d] = min(max(xx1[d] - (tw0 - 1), 0), TS_R - 1);
    }
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int ny = y + dy - 1;
      if (ny < 0 || ny >= OH) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int nx = x + dx - 1;
        if (nx < 0 || nx >= OW) continue;
        const int t = dy * 3 + dx;
        const float* q = sQ + ((ly_ + dy) * TS_P + lx_ + dx) * P;
        const float* r00 = sR + (yy0[dy] * TS_R + xx0[dx]) * P;
        const float* r01 = sR + (yy0[dy] * TS_R + xx1[dx]) * P;
        const float* r10 = sR + (yy1[dy] * TS_R + xx0[dx]) * P;
        const float* r11 = sR + (yy1[dy] * TS_R + xx1[dx]) * P;
        for (int oc = 0; oc < OC; ++oc) {
          const int col = oc * 9 + t;
          const float top = r00[col] * (1.f - ly[dy]) + r10[col] * ly[dy];
          const float bot = r01[col] * (1.f - ly[dy]) + r11[col] * ly[dy];
          acc[oc] += q[col] + top * (1.f - lx[dx]) + bot * lx[dx];
        }
      }
    }
    if (MODE == 1) {
      uint8_t* o = static_cast<uint8_t*>(out_v) + ((b * OH + y) * OW + x) * OC;
      for (int oc = 0; oc < OC; ++oc) o[oc] = static_cast<uint8_t>(__float2int_rn(fminf(fmaxf(acc[oc], 0.f), 1.f) * 255.f));
      continue;
    }
    float* out = static_cast<float*>(out_v);
    for (int oc = 0; oc < OC; ++oc) out[((b * OC + oc) * OH + y) * OW + x] = acc[oc];
    if (MODE == 2) {
      // train.py:437-443: luminance target, prob = sigmoid(logits), se = (logits - target)^2, weighted sums, Charbonnier
      const int64_t plane = static_cast<int64_t>(OH) * OW, pix = static_cast<int64_t>(y) * OW + x;
      const float w = ev.weight ? __ldg(ev.weight + b * plane + pix) : 1.f;
      for (int oc = 0; oc < OC; ++oc) {
        float t;
        if (ev.target_chans == OC) {
          t = __ldg(ev.target + (b * OC + oc) * plane + pix);
        } else {   // 3 -> 1
          const float* tp = ev.target + b * 3 * plane + pix;
          t = 0.2989f * __ldg(tp) + 0.5870f * __ldg(tp + plane) + 0.1140f * __ldg(tp + 2 * plane);
        }
        if (ev.prob) ev.prob[(b * OC + oc) * plane + pix] = 1.f / (1.f + expf(-acc[oc]));
        const float d = acc[oc] - t, se = d * d;
        part[0] += se;
        part[1] += static_cast<double>(se * w);
        part[2] += w;
        part[3] += static_cast<double>(sqrtf(se + ev.eps * ev.eps) * w);
      }
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
    if (blockIdx.x == 0 && threadIdx.x == 4) atomicAdd(ev.sums + 4, static_cast<double>(gridDim.x / (tiles_x * tiles_y)) * OH * OW * OC);
  }
}

int tail_stencil(const float* Qp, const float* Rb, void* out, int out_fmt, const EvalEpilogue* ev, int B, int H, int W, int OC, int NT,
                 cudaStream_t s) {
  if (OC < 1 || OC > 3 || NT < OC * 9 || NT % 4) return fail(SUNET_E_SHAPE, "tail: out_chans=%d (1..3) NT=%d", OC, NT);
  const dim3 grid(static_cast<unsigned>(B) * ((H + TS_T - 1) / TS_T) * ((W + TS_T - 1) / TS_T)), block(256);
  const int smem = (TS_P * TS_P + TS_R * TS_R) * OC * 9 * 4;
  EvalEpilogue e;
#define TS_LAUNCH_OC(MODE, OCV)                                                                                           \
  {                                                                                                                       \
    SUNET_CUDA(cudaFuncSetAttribute(tail_stencil_kernel<MODE, OCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));  \
    SUNET_CUDA(launch_pdl(tail_stencil_kernel<MODE, OCV>, grid, block, smem, s, Qp, Rb, out, e, H, W, NT));               \
  }
#define TS_LAUNCH(MODE)                                                                                                   \
  if (OC == 1) TS_LAUNCH_OC(MODE, 1) else if (OC == 2) TS_LAUNCH_OC(MODE, 2) else TS_LAUNCH_OC(MODE, 3)
  if (ev) {
    if (out_fmt != IMG_F32_NCHW) return fail(SUNET_E_ARG, "tail: the validation epilogue needs fp32 output");
    if (!ev->target || !ev->sums) return fail(SUNET_E_ARG, "tail: validation epilogue without target / sums");
    if (ev->target_chans != OC && !(ev->target_chans == 3 && OC == 1))
      return fail(SUNET_E_SHAPE, "tail: target has %d channels, output %d", ev->target_chans, OC);
    e = *ev;
    TS_LAUNCH(2)
  } else if (out_fmt == IMG_U8_NHWC) {
    TS_LAUNCH(1)
  } else {
    TS_LAUNCH(0)
  }
#undef TS_LAUNCH
#undef TS_LAUNCH_OC
  SUNET_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------ pre-pack helpers
__global__ void pack_weight_kernel(const float* __restrict__ src, __half* __restrict__ dst, int N, int K, int scale_rows, float scale) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(N) * K) return;
  const int n = static_cast<int>(i / K);
  dst[i] = __float2half_rn(src[i] * (n < scale_rows ? scale : 1.f));
}
int pack_weight_f16(const float* src, __half* dst, int N, int K, int scale_rows, float scale, cudaStream_t s) {
  pack_weight_kernel<<<blocks_for(static_cast<int64_t>(N) * K, 256), 256, 0, s>>>(src, dst, N, K, scale_rows, scale);
  SUNET_CHECK_LAUNCH();
  return 0;
}
// dst [N][K + N] = [ half(src) | I ]: the residual add of a square-ish projection becomes a second K segment of the GEMM
// (y = A W^T + R I), so the residual tile is prefetched by the TMA ring instead of being loaded in the epilogue
__global__ void pack_weight_residual_kernel(const float* __restrict__ src, __half* __restrict__ dst, int N, int K) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int KK = K + N;
  if (i >= static_cast<int64_t>(N) * KK) return;
  const int n = static_cast<int>(i / KK), k = static_cast<int>(i % KK);
  dst[i] = k < K ? __float2half_rn(src[static_cast<int64_t>(n) * K + k]) : __float2half_rn(k - K == n ? 1.f : 0.f);
}
int pack_weight_residual_f16(const float* src, __half* dst, int N, int K, cudaStream_t s) {
  pack_weight_residual_kernel<<<blocks_for(static_cast<int64_t>(N) * (K + N), 256), 256, 0, s>>>(src, dst, N, K);
  SUNET_CHECK_LAUNCH();
  return 0;
}
__global__ void pack_weight_shuffle_kernel(const float* __restrict__ src, __half* __restrict__ dst, int Cq, int rr, int K) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(Cq) * rr * K) return;
  const int n = static_cast<int>(i / K), k = static_cast<int>(i % K);
  const int c = n / rr, ij = n % rr;
  dst[(static_cast<int64_t>(ij) * Cq + c) * K + k] = __float2half_rn(src[i]);
}
int pack_weight_shuffle_f16(const float* src, __half* dst, int Cq, int rr, int K, cudaStream_t s) {
  pack_weight_shuffle_kernel<<<blocks_for(static_cast<int64_t>(Cq) * rr * K, 256), 256, 0, s>>>(src, dst, Cq, rr, K);
  SUNET_CHECK_LAUNCH();
  return 0;
}
__global__ void scale_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, int scale_n, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * (i < scale_n ? scale : 1.f);
}
int scale_copy_f32(const float* src, float* dst, int n, int scale_n, float scale, cudaStream_t s) {
  scale_copy_kernel<<<blocks_for(n, 256), 256, 0, s>>>(src, dst, n, scale_n, scale);
  SUNET_CHECK_LAUNCH();
  return 0;
}
__global__ void matmul_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, float* __restrict__ C,
                                  int ldc, int m, int n, int k) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(m) * n) return;
  const int r = static_cast<int>(i / n), c = static_cast<int>(i % n);
  float acc = 0.f;
  for (int j = 0; j < k; ++j) acc = fmaf(A[static_cast<int64_t>(r) * lda + j], B[static_cast<int64_t>(j) * ldb + c], acc);
  C[static_cast<int64_t>(r) * ldc + c] = acc;
}
int matmul_f32(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int m, int n, int k, cudaStream_t s) {
  matmul_f32_kernel<<<blocks_for(static_cast<int64_t>(m) * n, 256), 256, 0, s>>>(A, lda, B, ldb, C, ldc, m, n, k);
  SUNET_CHECK_LAUNCH();
  return 0;
}
__global__ void fold_patch_embed_kernel(const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                                        const float* __restrict__ b2, int Cin, int E, float* __restrict__ wfold,
                                        float* __restrict__ bfold) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int Kf = Cin * 36;
  if (i < Kf * E) {
    const int k = i / E, e = i % E;
    const int c = k / 36, u = (k / 6) % 6, v = k % 6;
    double acc = 0.0;
    for (int m = 0; m < E; ++m)
      for (int ky = 0; ky < 4; ++ky) {
        const int dy = u - ky;
        if (dy < 0 || dy > 2) continue;
        for (int kx = 0; kx < 4; ++kx) {
          const int dx = v - kx;
          if (dx < 0 || dx > 2) continue;
          acc += static_cast<double>(w2[((e * E + m) * 4 + ky) * 4 + kx]) * w1[((m * Cin + c) * 3 + dy) * 3 + dx];
        }
      }
    wfold[i] = static_cast<float>(acc);
  }
  if (i < E) {
    double acc = b2[i];
    for (int m = 0; m < E; ++m) {
      double sw = 0.0;
      for (int q = 0; q < 16; ++q) sw += w2[(i * E + m) * 16 + q];
      acc += sw * b1[m];
    }
    bfold[i] = static_cast<float>(acc);
  }
}
int fold_patch_embed(const float* w1, const float* b1, const float* w2, const float* b2, int Cin, int E, float* wfold, float* bfold,
                     cudaStream_t s) {
  fold_patch_embed_kernel<<<blocks_for(static_cast<int64_t>(Cin) * 36 * E, 128), 128, 0, s>>>(w1, b1, w2, b2, Cin, E, wfold, bfold);
  SUNET_CHECK_LAUNCH();
  return 0;
}
__global__ void pack_patch_embed_f16_kernel(const float* __restrict__ wfold, __half* __restrict__ wpk, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= E * PE_WP) return;
  const int e = i / PE_WP, k = i - e * PE_WP;
  wpk[i] = __float2half_rn(k < 108 ? wfold[k * E + e] : 0.f);
}
int pack_patch_embed_f16(const float* wfold, __half* wpk, int E, cudaStream_t s) {
  pack_patch_embed_f16_kernel<<<blocks_for(static_cast<int64_t>(E) * PE_WP, 128), 128, 0, s>>>(wfold, wpk, E);
  SUNET_CHECK_LAUNCH();
  return 0;
}
__global__ void fold_tail_taps_kernel(const float* __restrict__ Wo, const float* __restrict__ A, __half* __restrict__ G, int OC, int E,
                                      int NT) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NT * E) return;
  const int row = i / E, c = i % E;
  float acc = 0.f;
  if (row < OC * 9) {
    const int oc = row / 9, t = row % 9;
    double a = 0.0;
    for (int m = 0; m < E; ++m) a += static_cast<double>(Wo[(oc * E + m) * 9 + t]) * A[m * E + c];
    acc = static_cast<float>(a);
  }
  G[i] = __float2half_rn(acc);
}
int fold_tail_taps(const float* Wo, const float* A, __half* G, int OC, int E, int NT, cudaStream_t s) {
  fold_tail_taps_kernel<<<blocks_for(static_cast<int64_t>(NT) * E, 128), 128, 0, s>>>(Wo, A, G, OC, E, NT);
  SUNET_CHECK_LAUNCH();
  return 0;
}

}  // namespace sunet
