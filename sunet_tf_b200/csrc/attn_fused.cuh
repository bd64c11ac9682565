// Fused window attention front half for one Swin block (see attn_fused.cu):
//   norm1 -> cyclic shift -> window partition -> qkv Linear -> (q*scale) k^T + bias (+ mask) -> softmax -> @ v
//   -> window reverse -> un-shift          (SUNet_detail.py:233-257 with WindowAttention.forward :107-135)
// x and out are fp16 image-order token rows [B*H*W][C]; `out` holds the per-head attention output that feeds `proj`.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

struct AttnFusedPack {
  int C = 0, heads = 0;
  __half* w = nullptr;          // [3C][attn_fused_w_pitch(C)] qkv weight (+ the bias columns): rows permuted to [head group][q|k|v][head][d], LN gamma folded, q rows scaled
  float* hconst = nullptr;      // float [3C]: folded bias (b + W beta, q rows scaled) in the same row order
  const float* table = nullptr; // relative_position_bias_table fp32 [225][heads] (device)
  alignas(64) CUtensorMap tmW;
};

bool attn_fused_supported(int C, int heads);
// row pitch (elements) of the packed weight buffer w: C, or C rounded up to whole 64-column k-blocks where the qkv bias rides the MMA (C = 96)
int attn_fused_w_pitch(int C);
// gamma/beta: norm1; wqkv [3C][C], bqkv [3C] or null; qscale = qk_scale * log2(e)
int attn_fused_prepack(AttnFusedPack* p, int C, int heads, float qscale, const float* gamma, const float* beta, const float* wqkv,
                       const float* bqkv, const float* table, cudaStream_t stream);
int attn_fused_launch(const AttnFusedPack& p, const __half* x, __half* out, int B, int H, int W, int shift, cudaStream_t stream);

}  // namespace sunet
