// attn.proj + bias + shortcut, then norm2, as one tcgen05 kernel for the stages whose MLP runs as separate GEMMs (C = 384):
//   X1 = A * Wp^T + bp + R            (SUNet_detail.py:136, :261)        fp16 [M][C]
//   T  = LayerNorm(X1) * g + b        (:262, the input of mlp.fc1)       fp16 [M][C]
// One CTA owns whole 128-token rows (N = C accumulator columns in TMEM), so the row statistics never leave the SM and the
// separate LayerNorm launch (and its re-read of X1) disappears.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sunet {

struct ProjLnPack {
  const __half* w = nullptr;      // [C][C] fp16 (attn.proj.weight)
  const float* bias = nullptr;    // [C] or null
  const float* gamma = nullptr;   // norm2
  const float* beta = nullptr;
  int C = 0;
};

bool proj_ln_supported(int C);
// Same pipeline without the LayerNorm half: X = A[M][Kdim] * W[C][Kdim]^T + bias + R with whole rows per CTA (mlp.fc2 + second
// residual at C = 384: one wave of 128 full-row tiles instead of two partial waves of 128 x 192 tiles).
bool row_gemm_supported(int C, int Kdim);
int row_gemm_residual_launch(const __half* W, const float* bias, int C, int Kdim, const __half* A, const __half* R, __half* X,
                             int64_t M, cudaStream_t stream);
int proj_ln_launch(const ProjLnPack& p, const __half* A, const __half* R, __half* X1, __half* T, int64_t M, cudaStream_t stream);

}  // namespace sunet
