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
  alignas(64) CUtensorMap tmW1;
  alignas(64) CUtensorMap tmW2;
};

bool mlp_row_supported(int C);
// encodes the weight tensor maps (pointers must be set)
int mlp_row_prepare(MlpRowPack* p);
// T, R, X: [M][C] fp16 row-major, 16-byte aligned; X may alias R (each element is read and later written by the same thread),
// not T (other CTAs' tiles are still being read)
int mlp_row_launch(const MlpRowPack& p, const __half* T, const __half* R, __half* X, int64_t M, cudaStream_t stream);

}  // namespace sunet
