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
int proj_ln_launch(const ProjLnPack& p, const __half* A, const __half* R, __half* X1, __half* T, int64_t M, cudaStream_t stream);

}  // namespace sunet
