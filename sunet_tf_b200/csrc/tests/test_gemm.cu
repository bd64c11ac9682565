// Bring-up ladder for the tcgen05 GEMM (run on a B200 through gpurun):
//   1. umma_selftest: hand-swizzled smem -> one UMMA -> TMEM -> global   (descriptors / TMEM lane map)
//   2. gemm on small shapes vs a CPU fp32 reference on the same fp16-rounded operands
//   3. timing of the SUNet shapes (B=64) with CUDA events
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../error.h"
#include "../gemm.cuh"

using namespace sunet;

#define CK(x)                                                                    \
  do {                                                                           \
    cudaError_t e = (x);                                                         \
    if (e != cudaSuccess) {                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                   \
    }                                                                            \
  } while (0)

static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

static int g_fail = 0;

static void fill_half(std::vector<__half>& h, std::vector<float>& f, size_t n, float scale) {
  h.resize(n);
  f.resize(n);
  for (size_t i = 0; i < n; ++i) {
    h[i] = __float2half(frand() * scale);
    f[i] = __half2float(h[i]);
  }
}

static void test_selftest(int N) {
  std::vector<__half> hA, hB;
  std::vector<float> fA, fB;
  fill_half(hA, fA, 128 * 64, 2.f);
  fill_half(hB, fB, (size_t)N * 64, 2.f);
  __half *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * N * 4));
  int rc = umma_selftest(dA, dB, dD, N, 0);
  if (rc) { printf("selftest N=%d launch rc=%d %s\n", N, rc, last_error_buf()); g_fail++; return; }
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  int bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)fA[i * 64 + k] * fB[j * 64 + k];
      double err = fabs(ref - D[i * N + j]);
      if (!(err <= 1e-3)) { if (bad < 5) printf("  selftest mismatch (%d,%d): got %f ref %f\n", i, j, D[i * N + j], ref); bad++; }
      if (err > maxerr) maxerr = err;
    }
  printf("selftest N=%d: maxerr %.3g bad %d -> %s\n", N, maxerr, bad, bad ? "FAIL" : "ok");
  if (bad) g_fail++;
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

struct Case {
  int64_t M; int N, K0, K1; int bias, act, res, f32; int bn; int pair = 0;
};

static void test_gemm(const Case& c, bool check, int time_iters) {
  const int K = c.K0 + c.K1;
  std::vector<__half> hA0, hA1, hW, hR;
  std::vector<float> fA0, fA1, fW, fR, bias(c.N);
  fill_half(hA0, fA0, (size_t)c.M * c.K0, 2.f);
  if (c.K1) fill_half(hA1, fA1, (size_t)c.M * c.K1, 2.f);
  fill_half(hW, fW, (size_t)c.N * K, 0.5f);
  if (c.res) fill_half(hR, fR, (size_t)c.M * c.N, 2.f);
  for (auto& b : bias) b = frand();
  float slope = 0.25f;
  __half *dA0 = nullptr, *dA1 = nullptr, *dW = nullptr, *dR = nullptr;
  float *dBias = nullptr, *dSlope = nullptr;
  void* dC = nullptr;
  CK(cudaMalloc(&dA0, hA0.size() * 2));
  CK(cudaMemcpy(dA0, hA0.data(), hA0.size() * 2, cudaMemcpyHostToDevice));
  if (c.K1) { CK(cudaMalloc(&dA1, hA1.size() * 2)); CK(cudaMemcpy(dA1, hA1.data(), hA1.size() * 2, cudaMemcpyHostToDevice)); }
  CK(cudaMalloc(&dW, hW.size() * 2));
  CK(cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
  if (c.res) { CK(cudaMalloc(&dR, hR.size() * 2)); CK(cudaMemcpy(dR, hR.data(), hR.size() * 2, cudaMemcpyHostToDevice)); }
  CK(cudaMalloc(&dBias, c.N * 4));
  CK(cudaMemcpy(dBias, bias.data(), c.N * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dSlope, 4));
  CK(cudaMemcpy(dSlope, &slope, 4, cudaMemcpyHostToDevice));
  const size_t csz = (size_t)c.M * c.N * (c.f32 ? 4 : 2);
  CK(cudaMalloc(&dC, csz));
  CK(cudaMemset(dC, 0xff, csz));
  GemmArgs a;
  a.A0 = dA0; a.lda0 = c.K0; a.K0 = c.K0; a.A1 = dA1; a.lda1 = c.K1; a.K1 = c.K1;
  a.W = dW; a.ldw = K; a.M = c.M; a.N = c.N;
  a.bias = c.bias ? dBias : nullptr; a.act = c.act; a.prelu = dSlope; a.R = dR; a.ldr = c.N;
  a.C = dC; a.ldc = c.N; a.out_f32 = c.f32; a.force_block_n = c.bn; a.pair_mode = c.pair; a.dbg = getenv("SUNET_GEMM_DBG") ? atoi(getenv("SUNET_GEMM_DBG")) : 0;
  GemmOp op;
  int rc = gemm_prepare(a, &op);
  if (rc) { printf("gemm prepare rc=%d: %s\n", rc, last_error_buf()); g_fail++; return; }
  rc = gemm_launch(op, 0);
  if (rc) { printf("gemm launch rc=%d: %s\n", rc, last_error_buf()); g_fail++; return; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("gemm M=%lld N=%d K=%d+%d: CUDA error %s\n", (long long)c.M, c.N, c.K0, c.K1, cudaGetErrorString(e)); exit(3); }
  char tag[160];
  snprintf(tag, sizeof tag, "gemm M=%lld N=%d K=%d+%d bias%d act%d res%d f32%d bn=%d st=%d grid=%u%s", (long long)c.M, c.N, c.K0, c.K1,
           c.bias, c.act, c.res, c.f32, op.epi.block_n, op.epi.stages, op.grid, op.pair ? " PAIR" : "");
  if (check) {
    std::vector<uint8_t> hC(csz);
    CK(cudaMemcpy(hC.data(), dC, csz, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad = 0;
    // full check for small problems, sampled rows otherwise
    const int64_t row_step = c.M > 2048 ? 37 : 1;
    for (int64_t i = 0; i < c.M; i += row_step)
      for (int j = 0; j < c.N; ++j) {
        double ref = 0;
        for (int k = 0; k < c.K0; ++k) ref += (double)fA0[i * c.K0 + k] * fW[(size_t)j * K + k];
        for (int k = 0; k < c.K1; ++k) ref += (double)fA1[i * c.K1 + k] * fW[(size_t)j * K + c.K0 + k];
        if (c.bias) ref += bias[j];
        if (c.act == ACT_GELU) ref = 0.5 * ref * (1.0 + erf(ref / sqrt(2.0)));  // kernel uses the tanh-form fit: |dGELU| <= 3e-4*|x|
        if (c.act == ACT_PRELU) ref = ref >= 0 ? ref : slope * ref;
        if (c.res) ref += fR[i * c.N + j];
        float got = c.f32 ? reinterpret_cast<float*>(hC.data())[i * c.N + j]
                          : __half2float(reinterpret_cast<__half*>(hC.data())[i * c.N + j]);
        double tol = c.f32 ? 2e-3 : (2e-3 + fabs(ref) * 2e-3) + (c.act == ACT_GELU ? 2e-3 : 0.0);
        double err = fabs(ref - got);
        if (!(err <= tol)) { if (bad < 5) printf("  mismatch (%lld,%d): got %f ref %f\n", (long long)i, j, got, ref); bad++; }
        if (err > maxerr) maxerr = err;
      }
    printf("%s: maxerr %.3g bad %d -> %s\n", tag, maxerr, bad, bad ? "FAIL" : "ok");
    if (bad) g_fail++;
  }
  if (time_iters > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) gemm_launch(op, 0);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < time_iters; ++i) gemm_launch(op, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / time_iters;
    const double bytes = (double)c.M * K * 2 + (double)c.N * K * 2 + (double)csz + (c.res ? (double)c.M * c.N * 2 : 0);
    printf("%s: %.1f us  %.1f TFLOP/s  %.0f GB/s(min traffic)\n", tag, us, op.flops / us * 1e-6, bytes / us * 1e-3);
  }
  cudaFree(dA0); cudaFree(dA1); cudaFree(dW); cudaFree(dR); cudaFree(dBias); cudaFree(dSlope); cudaFree(dC);
}

int main(int argc, char** argv) {
  if (argc >= 10 && !strcmp(argv[1], "one")) {  // one M N K act res f32 bn iters
    Case c{atoll(argv[2]), atoi(argv[3]), atoi(argv[4]), 0, 1, atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8])};
    if (getenv("SUNET_GEMM_PAIR")) c.pair = atoi(getenv("SUNET_GEMM_PAIR"));
    test_gemm(c, false, atoi(argv[9]));
    return 0;
  }
  const bool do_time = argc > 1 && !strcmp(argv[1], "time");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  test_selftest(64);
  test_selftest(96);
  test_selftest(256);
  test_selftest(16);
  if (g_fail) { printf("SELFTEST FAILED - stopping before TMA gemm\n"); }
  const Case small[] = {
      {128, 64, 64, 0, 0, 0, 0, 1, 0},       // one tile, one k-block, fp32 out
      {256, 96, 64, 0, 0, 0, 0, 0, 0},
      {128, 96, 128, 0, 0, 0, 0, 1, 0},      // two k-blocks
      {1000, 288, 96, 0, 1, 0, 0, 0, 0},     // K tail (96), M tail
      {512, 384, 96, 0, 1, ACT_GELU, 0, 0, 0},
      {512, 96, 384, 0, 1, 0, 1, 0, 0},      // residual, 6 k-blocks
      {64, 768, 768, 0, 1, 0, 1, 0, 0},      // M < 128
      {300, 96, 96, 96, 1, 0, 0, 0, 0},      // concat K0+K1 with tails
      {640, 768, 384, 384, 1, 0, 0, 0, 0},
      {512, 16, 96, 0, 0, 0, 0, 1, 0},       // tail taps N=16 fp32
      {512, 1536, 96, 0, 0, ACT_PRELU, 0, 0, 256},
      {256, 3072, 768, 0, 1, ACT_GELU, 0, 0, 256},
      {256, 768, 3072, 0, 1, 0, 1, 0, 128},  // long K: ring wraps many times
      {4096, 192, 192, 0, 1, 0, 1, 0, 192},
      {384, 48, 96, 0, 0, 0, 0, 0, 0},
  };
  for (const Case& c : small) test_gemm(c, true, 0);
  for (Case c : small) {   // the same ladder on CTA pairs (cta_group::2) where the tile width allows it
    c.pair = 2;
    if (c.bn == 0) c.bn = c.N % 256 == 0 ? 256 : (c.N % 192 == 0 ? 192 : (c.N % 128 == 0 ? 128 : (c.N % 96 == 0 ? 96 : (c.N % 64 == 0 ? 64 : (c.N % 32 == 0 ? 32 : 0)))));
    if (c.bn == 0 || c.bn % 32) continue;
    test_gemm(c, true, 0);
  }
  if (do_time && !g_fail) {
    const Case big[] = {
        {262144, 288, 96, 0, 1, 0, 0, 0, 0},   {262144, 288, 96, 0, 1, 0, 0, 0, 96}, {262144, 288, 96, 0, 1, 0, 0, 0, 48},
        {262144, 96, 96, 0, 1, 0, 1, 0, 0},    {262144, 96, 96, 0, 1, 0, 1, 0, 48},
        {262144, 384, 96, 0, 1, ACT_GELU, 0, 0, 0}, {262144, 384, 96, 0, 1, ACT_GELU, 0, 0, 128}, {262144, 384, 96, 0, 1, ACT_GELU, 0, 0, 96},
        {262144, 384, 96, 0, 1, ACT_GELU, 0, 0, 192}, {262144, 384, 96, 0, 1, 0, 0, 0, 192},
        {262144, 96, 384, 0, 1, 0, 1, 0, 0},   {262144, 96, 384, 0, 1, 0, 1, 0, 48},
        {65536, 576, 192, 0, 1, 0, 0, 0, 0},   {65536, 576, 192, 0, 1, 0, 0, 0, 96}, {65536, 192, 192, 0, 1, 0, 1, 0, 0}, {65536, 192, 192, 0, 1, 0, 1, 0, 96},
        {65536, 768, 192, 0, 1, ACT_GELU, 0, 0, 0}, {65536, 768, 192, 0, 1, ACT_GELU, 0, 0, 128}, {65536, 192, 768, 0, 1, 0, 1, 0, 0}, {65536, 192, 768, 0, 1, 0, 1, 0, 96},
        {16384, 1152, 384, 0, 1, 0, 0, 0, 0},  {16384, 384, 384, 0, 1, 0, 1, 0, 0}, {16384, 384, 384, 0, 1, 0, 1, 0, 64},
        {16384, 1536, 384, 0, 1, ACT_GELU, 0, 0, 0}, {16384, 1536, 384, 0, 1, ACT_GELU, 0, 0, 128}, {16384, 384, 1536, 0, 1, 0, 1, 0, 0}, {16384, 384, 1536, 0, 1, 0, 1, 0, 64},
        {4096, 2304, 768, 0, 1, 0, 0, 0, 0},   {4096, 2304, 768, 0, 1, 0, 0, 0, 256}, {4096, 768, 768, 0, 1, 0, 1, 0, 0}, {4096, 768, 768, 0, 1, 0, 1, 0, 64},
        {4096, 3072, 768, 0, 1, ACT_GELU, 0, 0, 0}, {4096, 3072, 768, 0, 1, ACT_GELU, 0, 0, 256}, {4096, 768, 3072, 0, 1, 0, 1, 0, 0},
        {4096, 768, 3072, 0, 1, 0, 1, 0, 64}, {4096, 768, 3072, 0, 1, 0, 1, 0, 256},
        {8192, 8192, 8192, 0, 0, 0, 0, 0, 256},
        {262144, 1536, 96, 0, 0, ACT_PRELU, 0, 0, 0}, {4194304, 16, 96, 0, 0, 0, 0, 1, 0},
    };
    for (const Case& c : big) test_gemm(c, false, 20);
    printf("---- CTA pairs\n");
    for (Case c : big) {
      c.pair = 2;
      if (c.bn == 0) c.bn = c.N % 256 == 0 ? 256 : (c.N % 192 == 0 ? 192 : (c.N % 128 == 0 ? 128 : (c.N % 96 == 0 ? 96 : 0)));
      if (c.bn == 0 || c.bn % 32) continue;
      test_gemm(c, false, 20);
    }
  }
  printf(g_fail ? "RESULT: %d FAILURES\n" : "RESULT: ALL OK\n", g_fail);
  return g_fail ? 1 : 0;
}
