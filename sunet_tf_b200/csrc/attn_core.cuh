// Window-attention core launcher (see attn_core.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

struct AttnCoreArgs {
  const __half* qkv = nullptr;  // [rows][3C] fp16, q|k|v, each ordered [head][hd]; q already multiplied by qk_scale
  int64_t ld = 0;
  __half* out = nullptr;        // [rows][C] fp16, heads concatenated
  int64_t ldo = 0;
  int B = 0, H = 0, W = 0;      // images and token grid (image-order rows: (b*H + y)*W + x)
  int C = 0, heads = 0;
  int shift = 0;                // cyclic shift (0 or window/2); folded into the gather/scatter addresses
  const float* bias_table = nullptr;  // relative_position_bias_table, fp32 [225][heads]
  int mask_mode = 0;            // 0 none, 1 closed-form SW-MSA mask of the (H, W, shift) grid, 2 explicit tensor
  const float* mask = nullptr;  // explicit mask fp32 [mask_nw][64][64] (WindowAttention.forward(x, mask))
  int mask_nw = 0;
  int windowed_input = 0;       // 1: rows are already window-ordered (row = window*64 + token), no gather map
  int64_t num_windows = 0;      // only for windowed_input
};

int attn_core_launch(const AttnCoreArgs& a, cudaStream_t stream);

}  // namespace sunet
