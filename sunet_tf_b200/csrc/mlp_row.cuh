// mlp.fc1 -> GELU -> mlp.fc2 -> + residual with whole 128-token ROWS per CTA, for the stage whose rows are too wide for the
// chunk-fused kernel of mlp_fused.cu (C = 384: the fc2 accumulator alone takes 384 of the 512 TMEM columns).
//   X[m, :] = R[m, :] + fc2( GELU( fc1( T[m, :] ) ) )          (SUNet_detail.py:18-24, :262)
// T = norm2(x1) comes from proj_ln.cu (which also writes the residual R = x1).  The [M][4C] hidden activation (50 MB at
// M = 16384) never exists in HBM: fc1 accumulates 128 hidden units at a time into TMEM, the GELU epilogue re-packs them as the
// fp16 A operand of fc2 in shared memory, fc2 accumulates into the [128][C] accumulator that stays in TMEM for the whole row.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

struct MlpRowPack {
  int C = 0;
  const __half* w1h = nullptr;    // [4C][C]  fp16( 0.5 * fc1.weight )   (the GELU epilogue takes u = x / 2, act.cuh)
  const float* hbias = nullptr;   // [4C]     0.5 * fc1.bias
  const __half* w2 = nullptr;     // [C][4C]  fp16( fc2.weight )
  const float* b2 = nullptr;      // [C] or null
  // proj + shortcut + norm2 in front (mlp_row_set_proj):
  const __half* wp = nullptr;     // [C][C]   fp16( attn.proj.weight )
  const float* bp = nullptr;      // [C] or null
  const __half* w1hg = nullptr;   // [4C][C]  fp16( 0.5 * fc1.weight * norm2.weight[k] )
  const float* hbiasg = nullptr;  // [4C][2]  (s_n = sum_k w1hg[n,k],  c_n = 0.5 * (fc1.bias + fc1.weight norm2.bias)): norm2 folded into fc1
  int has_proj = 0;
  alignas(64) CUtensorMap tmW1;
  alignas(64) CUtensorMap tmW2;
  alignas(64) CUtensorMap tmWp;
  alignas(64) CUtensorMap tmW1g;
};

bool mlp_row_supported(int C);
// encodes the weight tensor maps (pointers must be set)
int mlp_row_prepare(MlpRowPack* p);
// T, R, X: [M][C] fp16 row-major, 16-byte aligned; X may alias R (each element is read and later written by the same thread),
// not T (other CTAs' tiles are still being read)
int mlp_row_launch(const MlpRowPack& p, const __half* T, const __half* R, __half* X, int64_t M, cudaStream_t stream);
// The whole back half of a Swin block in the same kernel (SUNet_detail.py:136 proj, :261 first residual, :262 norm2 + Mlp + second
// residual):  x1 = shortcut + attn_out Wp^T + bp ;  X = x1 + fc2(GELU(fc1(LN(x1)))).  w1 / b1 / gamma / beta: fp32 device parameters;
// w1hg [4C][C] fp16 and hbiasg [4C][2] fp32 are caller-allocated pack buffers.  X may alias shortcut, not attn_out.
int mlp_row_set_proj(MlpRowPack* p, const __half* wp, const float* bp, const float* gamma, const float* beta, const float* w1, const float* b1,
                     __half* w1hg, float* hbiasg, cudaStream_t stream);
int mlp_row_proj_launch(const MlpRowPack& p, const __half* attn_out, const __half* shortcut, __half* X, int64_t M, cudaStream_t stream);

}  // namespace sunet
