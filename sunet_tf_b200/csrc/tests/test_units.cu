// Micro-measurements of the two non-tensor units the fused kernels lean on (B200, sm_100a):
//   * MUFU issue rate (ex2.f32, tanh.f32, tanh.f16, rcp) per SM
//   * tcgen05.ld (TMEM -> registers) bandwidth per SM as a function of the number of reading warps and the load width
// Build: tools/build_tests.sh ; run on a B200: build/test_units
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../ptx.cuh"

using namespace sunet;

template <int OP>
__global__ void __launch_bounds__(1024, 1) mufu_kernel(float* out, int iters, long long* cycles) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i + 1);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (OP == 1) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      else if (OP == 3) { uint32_t u = __float_as_uint(a[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      else if (OP == 4) { uint32_t u = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); a[i] = __uint_as_float(u); }
      else a[i] = fmaf(a[i], 1.0001f, 0.5f);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int WIDTH>
__global__ void __launch_bounds__(512, 1) tmem_ld_kernel(float* out, int iters, long long* cycles) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t taddr = tmem_base_smem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 512; c += 128) {
      const uint32_t col = c + (warp >> 2) * 32;   // 4 column groups of 32 per 128-column slice
      if (WIDTH == 32) {
        uint32_t v[32];
        tmem_ld32(taddr + col, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      } else if (WIDTH == 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          tmem_ld8(taddr + col + 8 * j, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) acc ^= v[i];
        }
      } else {  // 4 x8 loads per wait
        uint32_t v[32];
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld8(taddr + col + 8 * j, *reinterpret_cast<uint32_t(*)[8]>(&v[8 * j]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base_smem, 512);
  }
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  long long h[148];
  const int iters = 2000;
  const char* names[6] = {"ex2.f32", "tanh.f32", "rcp.f32", "tanh.f16x2", "ex2.f16x2", "ffma"};
  for (int op = 0; op < 6; ++op) {
    for (int threads : {128, 256, 512, 1024}) {
      switch (op) {
        case 0: mufu_kernel<0><<<148, threads>>>(out, iters, cyc); break;
        case 1: mufu_kernel<1><<<148, threads>>>(out, iters, cyc); break;
        case 2: mufu_kernel<2><<<148, threads>>>(out, iters, cyc); break;
        case 3: mufu_kernel<3><<<148, threads>>>(out, iters, cyc); break;
        case 4: mufu_kernel<4><<<148, threads>>>(out, iters, cyc); break;
        default: mufu_kernel<5><<<148, threads>>>(out, iters, cyc); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double ops = 8.0 * iters * threads;
      printf("%-10s threads/SM %4d: %.2f lane-ops/clk/SM (%.1f clk per warp instruction per scheduler)\n", names[op], threads, ops / h[0],
             h[0] / (8.0 * iters * (threads / 32) / 4.0));
    }
  }
  for (int width : {32, 8, 0}) {
    for (int threads : {128, 256, 512}) {
      const int it2 = 2000;
      if (width == 32) tmem_ld_kernel<32><<<148, threads>>>(out, it2, cyc);
      else if (width == 8) tmem_ld_kernel<8><<<148, threads>>>(out, it2, cyc);
      else tmem_ld_kernel<0><<<148, threads>>>(out, it2, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = 4.0 * it2 * (threads / 32) * 32.0 * 32.0 * 4.0;   // per warp per iteration: 4 loads of 32 lanes x 32 columns x 4 B
      printf("tcgen05.ld %s warps %2d: %.1f B/clk/SM\n", width == 32 ? "x32      " : (width == 8 ? "x8       " : "4*x8/wait"), threads / 32, bytes / h[0]);
    }
  }
  return 0;
}
