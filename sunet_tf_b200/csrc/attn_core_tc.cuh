// Window-attention core on tcgen05 for head_dim 96 (stage 3 of SUNet: C = 768, 8 heads, one 8x8 window per image).  See attn_core_tc.cu.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

// true when the (C, heads, token grid, shift) of a block is the case the tcgen05 core is built for: head_dim 96, an even number of
// heads, and a token grid that IS one window (H = W = 8, so SUNet_detail.py:186-189 has set shift_size to 0 and the window order
// of :27-56 is the image order)
bool attn_core_tc_supported(int C, int heads, int H, int W, int shift);

// relative_position_bias_table fp32 [225][heads] (SUNet_detail.py:93-105) -> bias_exp fp32 [heads][64][64], times log2(e)
int attn_core_tc_expand_bias(const float* table, int heads, float* bias_exp, cudaStream_t stream);

// qkv [rows][3C] fp16 (q | k | v, each [head][96]; q pre-multiplied by qk_scale * log2(e)), rows = images * 64 in image order
// -> out [rows][C] fp16, heads concatenated (SUNet_detail.py:118-135)
int attn_core_tc_launch(const __half* qkv, int64_t ld, __half* out, int64_t ldo, int64_t rows, int C, int heads, const float* bias_exp,
                        cudaStream_t stream);

// bring-up: D[128][128] = A[128][64] * Bt[64][128] with B read from shared memory as an MN-major operand (Bt rows are 128-byte
// swizzled rows of 64 fp16, two column blocks `lbo` bytes apart, 8-row groups `sbo` bytes apart; desc_* are the values written
// into the descriptor fields - equal to the layout's in the product, separate here so that bring-up can probe the semantics)
int umma_mn_selftest(const __half* A, const __half* Bt, float* D, uint32_t lbo, uint32_t sbo, uint32_t desc_lbo, uint32_t desc_sbo,
                     cudaStream_t stream);

}  // namespace sunet
