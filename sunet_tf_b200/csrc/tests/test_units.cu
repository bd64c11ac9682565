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
      else if (OP == 5) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      else if (OP == 6) { uint32_t u; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(u | 0x3c000000u); }
      else if (OP == 7) { unsigned short hbits = static_cast<unsigned short>(__float_as_uint(a[i]) >> 13); asm volatile("cvt.f32.f16 %0, %1;" : "=f"(a[i]) : "h"(hbits)); }
      else if (OP == 8) { uint32_t u = __float_as_uint(a[i]); asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(u) : "r"(0x3c003c00u)); a[i] = __uint_as_float(u); }
      else if (OP == 9) { uint32_t u = __float_as_uint(a[i]); asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u) : "r"(0x42004200u)); a[i] = __uint_as_float(u); }
      else { uint32_t u = __float_as_uint(a[i]); asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(u) : "r"(__float_as_uint(a[(i + 1) & 7]))); a[i] = __uint_as_float(u); }
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

// MUFU + FMA mix: NM ex2 and NF independent FFMA per iteration - do the two pipes overlap (time = max) or serialise (sum)?
template <int NM, int NF>
__global__ void __launch_bounds__(1024, 1) mix_kernel(float* out, int iters, long long* cycles) {
  float a[8], b[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i + 1);
#pragma unroll
  for (int i = 0; i < 16; ++i) b[i] = 0.001f * (threadIdx.x + i + 1);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < NM / 4; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(r * (NM / 4) + i) & 7]));
#pragma unroll
      for (int i = 0; i < NF / 4; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[(r * (NF / 4) + i) & 15]) : "f"(1.0001f), "f"(0.5f));
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// legacy warp-level MMA rate: 8 independent m16n8k16 (fp16 in, fp32 accumulate) chains per warp
__global__ void __launch_bounds__(1024, 1) hmma_kernel(float* out, int iters, long long* cycles) {
  float d[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  const uint32_t a0 = 0x3c003c00u + threadIdx.x, b0 = 0x38003800u;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a0 + 1), "r"(a0 + 2), "r"(a0 + 3), "r"(b0), "r"(b0 + 1));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Cost of the barrier plumbing of a tcgen05 pipeline, one thread: MODE 0 = tcgen05.commit issue rate (no wait),
// 1 = commit -> mbarrier wait round trip, 2 = plain mbarrier.arrive -> wait round trip, 3 = try_wait on an already complete phase
template <int MODE>
__global__ void __launch_bounds__(128, 1) commit_kernel(int iters, long long* cycles) {
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_base_smem, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (MODE == 0) {
      for (int i = 0; i < iters; ++i) tc_commit(&bar);   // iters is even: the barrier is back at parity 0
      tc_commit(&bar);
      mbar_wait(&bar, 0);
    } else if (MODE == 1) {
      for (int i = 0; i < iters; ++i) { tc_commit(&bar); mbar_wait(&bar, i & 1); }
    } else if (MODE == 2) {
      for (int i = 0; i < iters; ++i) { mbar_arrive(&bar); mbar_wait(&bar, i & 1); }
    } else if (MODE == 3) {
      mbar_arrive(&bar);
      for (int i = 0; i < iters; ++i) mbar_wait(&bar, 0);
    } else {
      mbar_arrive(&bar);
      uint32_t acc = 0;
      for (int i = 0; i < iters; ++i) {
        uint32_t ok;
        if (MODE == 4)
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        else if (MODE == 5)
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        else if (MODE == 6)
          asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        else {   // 4 independent test_waits in flight (latency vs throughput)
          uint32_t o2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(o2[j]) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
          ok = o2[0] + o2[1] + o2[2] + o2[3];
        }
        if (ok == 0) break;   // dependent use every iteration
        acc += ok;
      }
      if (acc == 12345) cycles[1] = acc;
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base_smem, 32); }
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
  const char* names[11] = {"ex2.f32", "tanh.f32", "rcp.f32", "tanh.f16x2", "ex2.f16x2", "ffma", "cvt.f16x2.f32", "cvt.f32.f16", "hfma2", "hmin2", "prmt"};
  for (int op = 0; op < 11; ++op) {
    for (int threads : {128, 256, 512, 1024}) {
      switch (op) {
        case 0: mufu_kernel<0><<<148, threads>>>(out, iters, cyc); break;
        case 1: mufu_kernel<1><<<148, threads>>>(out, iters, cyc); break;
        case 2: mufu_kernel<2><<<148, threads>>>(out, iters, cyc); break;
        case 3: mufu_kernel<3><<<148, threads>>>(out, iters, cyc); break;
        case 4: mufu_kernel<4><<<148, threads>>>(out, iters, cyc); break;
        case 5: mufu_kernel<5><<<148, threads>>>(out, iters, cyc); break;
        case 6: mufu_kernel<6><<<148, threads>>>(out, iters, cyc); break;
        case 7: mufu_kernel<7><<<148, threads>>>(out, iters, cyc); break;
        case 8: mufu_kernel<8><<<148, threads>>>(out, iters, cyc); break;
        case 9: mufu_kernel<9><<<148, threads>>>(out, iters, cyc); break;
        default: mufu_kernel<10><<<148, threads>>>(out, iters, cyc); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double ops = 8.0 * iters * threads;
      printf("%-10s threads/SM %4d: %.2f lane-ops/clk/SM (%.1f clk per warp instruction per scheduler)\n", names[op], threads, ops / h[0],
             h[0] / (8.0 * iters * (threads / 32) / 4.0));
    }
  }
  for (int threads : {128, 256, 512, 1024}) {
    hmma_kernel<<<148, threads>>>(out, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per = static_cast<double>(h[0]) / (8.0 * iters * (threads / 128.0));
    printf("mma.sync.m16n8k16 f16->f32, %2d warps/SM: %.2f clk per instruction per scheduler = %.0f dense TFLOP/s at 1.9 GHz\n", threads / 32, per,
           4096.0 / per * 4 * 148 * 1.9e9 / 1e12);
  }
  for (int mode = 0; mode < 8; ++mode) {
    const int n = 1000;
    switch (mode) {
      case 0: commit_kernel<0><<<148, 128>>>(n, cyc); break;
      case 1: commit_kernel<1><<<148, 128>>>(n, cyc); break;
      case 2: commit_kernel<2><<<148, 128>>>(n, cyc); break;
      case 3: commit_kernel<3><<<148, 128>>>(n, cyc); break;
      case 4: commit_kernel<4><<<148, 128>>>(n, cyc); break;
      case 5: commit_kernel<5><<<148, 128>>>(n, cyc); break;
      case 6: commit_kernel<6><<<148, 128>>>(n, cyc); break;
      default: commit_kernel<7><<<148, 128>>>(n, cyc); break;
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    static const char* nm[8] = {"tcgen05.commit issue + deferred waits", "tcgen05.commit -> mbarrier wait round trip", "mbarrier.arrive -> wait round trip", "try_wait on a completed phase", "test_wait (acquire) completed phase", "try_wait.relaxed completed phase", "test_wait.relaxed completed phase", "4 x test_wait in flight"};
    printf("%-44s: %.0f clk per iteration\n", nm[mode], static_cast<double>(h[0]) / n);
  }
  for (int cfg = 0; cfg < 4; ++cfg) {
    for (int threads : {128, 512}) {
      int nm = 0, nf = 0;
      switch (cfg) {
        case 0: nm = 8; nf = 0; mix_kernel<8, 0><<<148, threads>>>(out, iters, cyc); break;
        case 1: nm = 0; nf = 64; mix_kernel<0, 64><<<148, threads>>>(out, iters, cyc); break;
        case 2: nm = 8; nf = 64; mix_kernel<8, 64><<<148, threads>>>(out, iters, cyc); break;
        default: nm = 8; nf = 32; mix_kernel<8, 32><<<148, threads>>>(out, iters, cyc); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("mix %d ex2 + %2d ffma per iteration, %2d warps/SM: %.1f clk per iteration per scheduler-warp (MUFU alone %d, FMA alone %d)\n", nm, nf,
             threads / 32, static_cast<double>(h[0]) / iters / (threads / 128.0), nm * 8, nf);
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
