// Fused pixel-shuffle branch of the final x4 up-sample (see tail_fused.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

bool tail_up_fused_supported(int E, int NT);
// T [M][96] fp16 tokens (after norm_up); w_p0 [16 * 96][96] fp16 (rows ordered sub-pixel major); g_p [16][96] fp16 folded tap maps;
// out [M * 16][16] fp32
int tail_up_fused_launch(const __half* T, const __half* w_p0, const __half* g_p, const float* slope, float* out, int64_t M, cudaStream_t stream);

}  // namespace sunet
