// Fused  x + Mlp(LayerNorm(x))  for one Swin block (SUNet_detail.py:262 -> :18-24), one kernel, sm_100a.
//   y[m, :] = x[m, :] + fc2( GELU( fc1( LN(x[m, :]) ) ) )
// The hidden activation (4C wide) never leaves the SM: fc1 accumulates into TMEM, the GELU epilogue re-packs it as
// the fp16 A operand of fc2 in shared memory, fc2 accumulates into TMEM, and the residual add happens on the way out.
// LayerNorm is folded into fc1 (exact algebra, see mlp_fused.cu) so the raw fp16 token tile is the MMA operand.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

struct MlpFusedPack {
  int C = 0;                 // channels (96 or 192); hidden = 4C
  __half* w1g = nullptr;     // [4C][C]  fp16( fc1.weight * gamma[k] )
  __half* w2 = nullptr;      // [C][4C]  fp16( fc2.weight )
  float* hconst = nullptr;   // [4C][2]  (s_n = sum_k w1g[n,k],  b1f_n = fc1.bias[n] + sum_k fc1.weight[n,k] * beta[k])
  float* b2 = nullptr;       // [C]
  __half* wp = nullptr;      // [C][C]   fp16( attn.proj.weight )   (proj + shortcut + MLP variant)
  float* bp = nullptr;       // [C]      attn.proj.bias
  __half* w1h = nullptr;     // [4C][mlp_fused_w1h_pitch(C)]  fp16( 0.5 * fc1.weight * gamma[k] )  (+ the bias columns)   (proj variant: LayerNorm runs in the kernel)
  float* hbias = nullptr;    // [4C]     0.5 * (fc1.bias + fc1.weight beta)
  int has_proj = 0;
  alignas(64) CUtensorMap tmW1;
  alignas(64) CUtensorMap tmW2;
  alignas(64) CUtensorMap tmWp;
  alignas(64) CUtensorMap tmW1h;
};

bool mlp_fused_supported(int C);
// row pitch (elements) of the w1h pack buffer: C, or C rounded up to whole 64-column k-blocks when the fc1 bias rides the MMA (C = 96)
int mlp_fused_w1h_pitch(int C);
// fp32 parameters (device) -> pack; w1g / w2 / hconst / b2 must already be allocated by the caller
int mlp_fused_prepack(MlpFusedPack* p, int C, const float* gamma, const float* beta, const float* w1, const float* b1,
                      const float* w2, const float* b2, cudaStream_t stream);
// optional: attn.proj weights (+ the norm2 / fc1 parameters again, packed differently) for the proj + shortcut + MLP kernel;
// wp / bp / w1h / hbias allocated by the caller
int mlp_fused_set_proj(MlpFusedPack* p, const float* wp, const float* bp, const float* gamma, const float* beta, const float* w1,
                       const float* b1, cudaStream_t stream);
// out = x1 + Mlp(LN(x1)),  x1 = shortcut + attn_out Wp^T + bp   (SUNet_detail.py:136, :261-262); out may alias shortcut, not attn_out
int mlp_proj_fused_launch(const MlpFusedPack& p, const __half* attn_out, const __half* shortcut, __half* out, int64_t M, cudaStream_t stream);
// x, out: [M][C] fp16 row-major (out may alias x)
int mlp_fused_launch(const MlpFusedPack& p, const __half* x, __half* out, int64_t M, cudaStream_t stream);

}  // namespace sunet
