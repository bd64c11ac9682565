// Activation helpers shared by the GEMM epilogue and the fused MLP kernel.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace sunet {

// GELU(x) = x * Phi(x), Phi(x) = 0.5 * (1 + erf(x / sqrt 2)) (nn.GELU default, SUNet_detail.py:9).
// Phi is evaluated as 0.5 + 0.5 * tanh(u * (c0 + c1 u^2 + c2 u^4)), u = clamp(x, +-7): a minimax fit of atanh(erf(x/sqrt2))
// with max |dPhi| = 5.0e-5 (fit) + 2.4e-4 (tanh.approx.f32, 2^-11 relative); ~9 instructions instead of ~40 for erff().
// The hidden activation is stored in fp16 (2^-11 relative) anyway; measured effect on the SUNet output < 1e-4.
__device__ __forceinline__ float gelu_fast(float x) {
  const float u = fminf(fmaxf(x, -7.0f), 7.0f);
  const float t = u * u;
  float p = fmaf(t, -3.5151847398e-04f, 3.7005657172e-02f);
  p = fmaf(t, p, 7.9750787105e-01f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u * p));
  const float h = 0.5f * x;
  return fmaf(h, th, h);
}

// GELU(x) from u = x / 2 (the caller pre-scales), in the fused MLP kernels whose GELU pass is FMA-issue-bound.
// SUNET_GELU_TERMS 3:  u + u * tanh(u * (2 c0 + 8 c1 u^2 + 32 c2 u^4)), the fit of gelu_fast (max |dGELU| 2.5e-5): 6 FMA-pipe
//                      instructions + one MUFU per element (the clamp keeps the odd polynomial monotone);
// SUNET_GELU_TERMS 2:  u + u * tanh(u * (a + b u^2)) with the minimax pair (a, b) = (1.60031416, 0.27760715): max |dGELU| 2.7e-4 over
//                      all x - below the 4.9e-4 half-ulp of the fp16 hidden activation it is stored in around |x| ~ 1.7 where the
//                      maximum sits - monotone without a clamp: 4 FMA-pipe instructions + one MUFU per element.
#ifndef SUNET_GELU_TERMS
#define SUNET_GELU_TERMS 2   // measured: 7.756 -> 7.619 ms per batch; whole-model max-abs 3.0e-4 / 3.4e-4 / 4.0e-4 / 1.0e-3 (init, stress, 64 distinct, outlier) against 3.3e-4 / 3.1e-4 / 4.7e-4 / 8.6e-4 with 3 terms (bar 2e-3)
#endif
__device__ __forceinline__ float gelu_half_arg(float u) {
#if SUNET_GELU_TERMS == 2
  const float p = fmaf(u * u, 2.7760715e-01f, 1.60031416f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u * p));
  return fmaf(u, th, u);
#else
  const float t = fminf(u * u, 12.25f);
  float p = fmaf(t, -1.1248591167e-02f, 2.9604525738e-01f);
  p = fmaf(t, p, 1.5950157421f);
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u * p));
  return fmaf(u, th, u);
#endif
}

// Packed-half2 variant for the fc1 epilogue (the result is stored as fp16 anyway): 5 instructions per element.
// t = min(u^2, 49) keeps the odd polynomial monotone for any |x| (tanh saturates to +-1 long before).
__device__ __forceinline__ uint32_t gelu_fast_h2(uint32_t xbits) {
  const __half2 x = *reinterpret_cast<const __half2*>(&xbits);
  const __half2 t = __hmin2(__hmul2(x, x), __float2half2_rn(49.0f));
  __half2 p = __hfma2(t, __float2half2_rn(-3.5151847398e-04f), __float2half2_rn(3.7005657172e-02f));
  p = __hfma2(t, p, __float2half2_rn(7.9750787105e-01f));
  const __half2 z = __hmul2(x, p);
  uint32_t th;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(th) : "r"(*reinterpret_cast<const uint32_t*>(&z)));
  const __half2 h = __hmul2(x, __float2half2_rn(0.5f));
  const __half2 g = __hfma2(h, *reinterpret_cast<const __half2*>(&th), h);
  return *reinterpret_cast<const uint32_t*>(&g);
}

}  // namespace sunet
