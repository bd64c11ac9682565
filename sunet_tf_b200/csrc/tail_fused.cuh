// Fused pixel-shuffle branch of the final x4 up-sample (see tail_fused.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

bool tail_up_fused_supported(int E, int NT, int W);   // W: width of the token grid (the staged bilinear taps are sized for <= 128)
// T [M][96] fp16 tokens (after norm_up), M = B * H * W in image order; w_p0 [16 * 96][96] fp16 (rows ordered sub-pixel major);
// g_p [16][96] fp16 folded tap maps (out_chans = 1: 9 taps, padded); Rb [M][16] fp32 = the bilinear branch's tap maps per token;
// strips: tail_strips_floats(M) fp32 partial output sums, consumed by tail_finish
int tail_up_fused_launch(const __half* T, const __half* w_p0, const __half* g_p, const float* slope, const float* Rb, float* strips, int B,
                         int H, int W, cudaStream_t stream);
inline int64_t tail_strips_floats(int64_t M) { return (M + 31) / 32 * 32 * 96; }
// strips -> out: fp32 NCHW (B, 1, 4H, 4W), or - out_fmt IMG_U8_NHWC - 8-bit, or fp32 + the validation epilogue (see elementwise.cuh)
struct EvalEpilogue;
int tail_finish(const float* strips, void* out, int out_fmt, const EvalEpilogue* ev, int B, int H, int W, cudaStream_t stream);

}  // namespace sunet
