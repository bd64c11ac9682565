// Window-attention core: per (window, head)  S = Q K^T + rel-pos bias (+ shifted-window mask),
// softmax over the 64 keys, O = P V.   Restates SUNet_detail.py:118-135 for 8x8 windows.
//
// The cyclic shift (torch.roll, :238/:255) and window_partition / window_reverse (:27-56) are never
// materialised: q/k/v rows are gathered from the image-order token tensor with
//     src = ((wr*8 + r + shift) mod H, (wc*8 + c + shift) mod W)
// and O is scattered back through the same map.  q arrives pre-multiplied by qk_scale (folded into the
// qkv weights at pre-pack).
//
// Round-1 implementation: register-resident flash-style core on mma.sync.m16n8k16 (fp16 in, fp32 accum);
// the linear layers around it (>= 90% of the block FLOPs) run on tcgen05 (gemm_tcgen05.cu).
// CTA = one window x HPC heads, 8 warps; a warp owns (head, 16-row query tiles).
#include "attn_core.cuh"
#include "error.h"
#include "ptx.cuh"

namespace sunet {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int HD, int HPC>
__global__ void __launch_bounds__(256) attn_core_kernel(const AttnCoreArgs p) {
  constexpr int HD_PAD = (HD + 15) / 16 * 16;
  constexpr int LDS = HD_PAD + 8;           // fp16 elements per smem row (pad breaks ldmatrix bank conflicts)
  constexpr int KS = HD_PAD / 16;           // k-steps of Q K^T
  constexpr int NO = HD_PAD / 8;            // n-tiles of O
  constexpr int WPH = 8 / HPC;              // warps per head
  constexpr int MT = 4 / WPH;               // 16-row query tiles per warp
  constexpr int SEG = HPC * HD;             // contiguous fp16 per token per q/k/v segment handled by this CTA
  constexpr int CHUNKS = SEG / 8;           // 16-byte chunks per segment
  static_assert(SEG % 8 == 0, "segment must be 16-byte granular");
  constexpr int TBL = 232;                  // 225 padded

  extern __shared__ __align__(16) uint8_t smem_raw[];
  __half* sQ = reinterpret_cast<__half*>(smem_raw);            // [HPC][64][LDS]
  __half* sK = sQ + HPC * 64 * LDS;
  __half* sV = sK + HPC * 64 * LDS;
  float* sTbl = reinterpret_cast<float*>(sV + HPC * 64 * LDS);  // [HPC][TBL]
  __shared__ long long sRow[64];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int win = blockIdx.x;
  const int hg = blockIdx.y;
  const int nWc = p.W >> 3, nWr = p.H >> 3;
  const int nW = nWr * nWc;
  const int wimg = win % nW;
  const int wr = wimg / nWc, wc = wimg % nWc;

  if (tid < 64) {
    long long row;
    if (p.windowed_input) {
      row = static_cast<long long>(win) * 64 + tid;
    } else {
      const int b = win / nW;
      const int r = (wr * 8 + (tid >> 3) + p.shift) % p.H;
      const int c = (wc * 8 + (tid & 7) + p.shift) % p.W;
      row = (static_cast<long long>(b) * p.H + r) * p.W + c;
    }
    sRow[tid] = row;
  }
  // bias table slice for this CTA's heads: table is [225][heads]
  for (int i = tid; i < HPC * 225; i += 256) {
    const int h = i / 225, e = i % 225;
    sTbl[h * TBL + e] = __ldg(p.bias_table + e * p.heads + hg * HPC + h);
  }
  // zero the K-padding columns (and the unused tail) once
  if constexpr (HD_PAD != HD) {
    constexpr int PADW = (HD_PAD - HD) / 2;  // half2 words per row
    for (int i = tid; i < 3 * HPC * 64 * PADW; i += 256) {
      const int row = i / PADW, w = i % PADW;
      reinterpret_cast<uint32_t*>(sQ + row * LDS + HD)[w] = 0u;
    }
  }
  __syncthreads();

  // ---- gather q/k/v rows: 16-byte global loads, 4-byte smem scatter (head boundaries are 4-byte granular)
  for (int i = tid; i < 64 * 3 * CHUNKS; i += 256) {
    const int t = i / (3 * CHUNKS);
    const int rem = i % (3 * CHUNKS);
    const int seg = rem / CHUNKS, ch = rem % CHUNKS;
    const __half* src = p.qkv + sRow[t] * p.ld + seg * p.C + hg * SEG + ch * 8;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    __half* base = (seg == 0 ? sQ : (seg == 1 ? sK : sV));
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = ch * 8 + j * 2;       // element within the segment
      const int h = e / HD, c = e % HD;
      *reinterpret_cast<uint32_t*>(base + (h * 64 + t) * LDS + c) = w[j];
    }
  }
  __syncthreads();

  const int h = warp % HPC;
  const int mbase = (warp / HPC) * MT;
  const __half* q_h = sQ + h * 64 * LDS;
  const __half* k_h = sK + h * 64 * LDS;
  const __half* v_h = sV + h * 64 * LDS;
  const float* tbl = sTbl + h * TBL;
  const int g = lane >> 2, tq = lane & 3;
  constexpr float LOG2E = 1.4426950408889634f;

  for (int mi = 0; mi < MT; ++mi) {
    const int mt = mbase + mi;
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldsm_x4(qa[ks], smem_u32(q_h + (mt * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8));
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
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
    // ---- bias + mask, rows i0 = mt*16+g and i1 = i0+8, keys j = nt*8 + 2*tq + {0,1}
    const int i0 = mt * 16 + g, i1 = i0 + 8;
    const int ri0 = i0 >> 3, ci0 = i0 & 7, ri1 = i1 >> 3, ci1 = i1 & 7;
    const int sb = 8 - p.shift;   // tokens with r (c) >= sb wrapped around from the other image edge
    const bool mrow = p.mask_mode == 1 && wr == nWr - 1;
    const bool mcol = p.mask_mode == 1 && wc == nWc - 1;
    const float* mexp = p.mask_mode == 2 ? p.mask + static_cast<long long>(win % p.mask_nw) * 4096 : nullptr;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = nt * 8 + 2 * tq + e;
        const int rj = j >> 3, cj = j & 7;   // rj == nt
        float b0 = tbl[(ri0 - rj + 7) * 15 + (ci0 - cj + 7)];
        float b1 = tbl[(ri1 - rj + 7) * 15 + (ci1 - cj + 7)];
        if (mrow) {
          if ((ri0 >= sb) != (rj >= sb)) b0 -= 100.f;
          if ((ri1 >= sb) != (rj >= sb)) b1 -= 100.f;
        }
        if (mcol) {
          // the reference mask is 0 / -100 per pair (regions differ in row OR col split): apply at most once
          if ((ci0 >= sb) != (cj >= sb) && !(mrow && (ri0 >= sb) != (rj >= sb))) b0 -= 100.f;
          if ((ci1 >= sb) != (cj >= sb) && !(mrow && (ri1 >= sb) != (rj >= sb))) b1 -= 100.f;
        }
        if (mexp) {
          b0 += __ldg(mexp + i0 * 64 + j);
          b1 += __ldg(mexp + i1 * 64 + j);
        }
        s[nt][e] += b0;
        s[nt][2 + e] += b1;
      }
    }
    // ---- softmax over 64 keys (each row lives in the 4 lanes of a quad)
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
    const float mm0 = m0 * LOG2E, mm1 = m1 * LOG2E;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] * LOG2E - mm0);
      s[nt][1] = exp2f(s[nt][1] * LOG2E - mm0);
      s[nt][2] = exp2f(s[nt][2] * LOG2E - mm1);
      s[nt][3] = exp2f(s[nt][3] * LOG2E - mm1);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
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
    // ---- normalise and park O in this tile's own Q rows (already consumed into registers)
    __syncwarp();
    __half* o_h = sQ + h * 64 * LDS;
#pragma unroll
    for (int n = 0; n < NO; ++n) {
      const int c = n * 8 + 2 * tq;
      if (c < HD) {
        *reinterpret_cast<uint32_t*>(o_h + i0 * LDS + c) = pack_half2(o[n][0] * inv0, o[n][1] * inv0);
        *reinterpret_cast<uint32_t*>(o_h + i1 * LDS + c) = pack_half2(o[n][2] * inv1, o[n][3] * inv1);
      }
    }
  }
  __syncthreads();
  // ---- scatter O rows back (16-byte coalesced stores; heads are concatenated in order, :135)
  for (int i = tid; i < 64 * CHUNKS; i += 256) {
    const int t = i / CHUNKS, ch = i % CHUNKS;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = ch * 8 + j * 2;
      const int hh = e / HD, c = e % HD;
      w[j] = *reinterpret_cast<const uint32_t*>(sQ + (hh * 64 + t) * LDS + c);
    }
    *reinterpret_cast<uint4*>(p.out + sRow[t] * p.ldo + hg * SEG + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

template <int HD, int HPC>
static int launch_core(const AttnCoreArgs& a, int64_t windows, cudaStream_t stream) {
  constexpr int HD_PAD = (HD + 15) / 16 * 16;
  constexpr int LDS = HD_PAD + 8;
  const int smem = 3 * HPC * 64 * LDS * 2 + HPC * 232 * 4;
  static bool configured = false;
  if (!configured) {
    SUNET_CUDA(cudaFuncSetAttribute(attn_core_kernel<HD, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(static_cast<unsigned>(windows), a.heads / HPC);
  attn_core_kernel<HD, HPC><<<grid, 256, smem, stream>>>(a);
  SUNET_CHECK_LAUNCH();
  return 0;
}

int attn_core_launch(const AttnCoreArgs& a, cudaStream_t stream) {
  if (a.C % a.heads != 0) return fail(SUNET_E_SHAPE, "attn: C=%d not divisible by heads=%d", a.C, a.heads);
  if (a.heads % 4 != 0) return fail(SUNET_E_SHAPE, "attn: heads=%d must be a multiple of 4", a.heads);
  if (a.H % 8 || a.W % 8) return fail(SUNET_E_SHAPE, "attn: token grid %dx%d must be a multiple of the 8x8 window", a.H, a.W);
  if (a.ld % 8 || a.ldo % 8 || (reinterpret_cast<uintptr_t>(a.qkv) & 15) || (reinterpret_cast<uintptr_t>(a.out) & 15))
    return fail(SUNET_E_ALIGN, "attn: qkv/out need 16-byte aligned rows");
  if (a.mask_mode == 2 && (a.mask == nullptr || a.mask_nw <= 0)) return fail(SUNET_E_ARG, "attn: explicit mask missing");
  const int hd = a.C / a.heads;
  const int64_t windows = a.windowed_input ? a.num_windows : static_cast<int64_t>(a.B) * (a.H / 8) * (a.W / 8);
  if (windows <= 0 || windows > 0x7fffffff) return fail(SUNET_E_SHAPE, "attn: bad window count");
  const bool h8 = a.heads % 8 == 0;
  switch (hd) {
    case 12: return h8 ? launch_core<12, 8>(a, windows, stream) : launch_core<12, 4>(a, windows, stream);
    case 24: return h8 ? launch_core<24, 8>(a, windows, stream) : launch_core<24, 4>(a, windows, stream);
    case 48: return h8 ? launch_core<48, 8>(a, windows, stream) : launch_core<48, 4>(a, windows, stream);
    case 96: return launch_core<96, 4>(a, windows, stream);
    case 16: return h8 ? launch_core<16, 8>(a, windows, stream) : launch_core<16, 4>(a, windows, stream);
    case 32: return h8 ? launch_core<32, 8>(a, windows, stream) : launch_core<32, 4>(a, windows, stream);
    default: return fail(SUNET_E_SHAPE, "attn: head_dim %d not instantiated (12/16/24/32/48/96)", hd);
  }
}

}  // namespace sunet
