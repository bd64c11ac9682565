// What bounds the L2 -> shared-memory operand stream of the tcgen05 GEMMs at ~42 B/clk/SM?
// One CTA per SM runs a TMA-only ring (no MMA): a producer thread issues 2-D tile loads of `box_rows` x 64 fp16 (128-byte rows, SW128)
// into a 4-stage ring, a consumer thread waits for each stage and frees it.  Modes:
//   0  every CTA streams its own rows of a big matrix                (distinct lines: L2 -> SM bandwidth, all traffic unique)
//   1  every CTA loads the SAME tile sequence                        (one copy in L2 serves all SMs: is the cap on the L2 side?)
//   2  clusters of CS CTAs, each CTA loads 1/CS of the tile and MULTICASTS it to the whole cluster (SM ingest without L2 reads)
// Prints bytes per clock per SM (landed in shared memory).  Build: tools/build_tests.sh; run on a B200: build/test_ingest
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../ptx.cuh"

using namespace sunet;

static constexpr int STAGES = 4;

__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// box_rows rows per stage in this CTA's smem; iters stages per CTA
__global__ void __launch_bounds__(128, 1) ingest_kernel(const __grid_constant__ CUtensorMap tm, int mode, int box_rows, int part_rows,
                                                        int iters, int rows_total, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = box_rows * 128;
  uint32_t csize = 1, crank = 0;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csize));
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], mode == 2 ? csize : 1); }
    fence_mbar_init();
  }
  if (csize > 1) cluster_sync_all(); else __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {            // producer
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
      mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
      int row0;
      if (mode == 0) row0 = ((blockIdx.x * iters + it) * box_rows) % (rows_total - box_rows);
      else row0 = (it * box_rows) % (rows_total - box_rows);
      if (mode == 2) {
        // this CTA fetches rows [crank * part_rows, +part_rows) of the tile and multicasts them to the same offset in every CTA
        tma_load_2d_mc(smem + s * stage_bytes + crank * part_rows * 128, &tm, &full_bar[s], 0, row0 + crank * part_rows,
                       static_cast<uint16_t>((1u << csize) - 1));
      } else {
        for (int r = 0; r < box_rows; r += part_rows) tma_load_2d(smem + s * stage_bytes + r * 128, &tm, &full_bar[s], 0, row0 + r);
      }
    }
  } else if (threadIdx.x == 32) {    // consumer
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      if (mode == 2) { for (uint32_t c = 0; c < csize; ++c) mbar_arrive_cluster(&empty_bar[s], c); }   // every CTA's copy of this stage is consumed
      else mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (csize > 1) cluster_sync_all();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = reinterpret_cast<PFN_encodeTiled>(fn);
  const int rows_total = 1 << 20;                 // 1M rows x 64 fp16 = 128 MB
  void* buf;
  cudaMalloc(&buf, static_cast<size_t>(rows_total) * 128);
  cudaMemset(buf, 1, static_cast<size_t>(rows_total) * 128);
  long long* dcyc;
  cudaMalloc(&dcyc, 148 * sizeof(long long));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 2000;
  struct Cfg { int mode, box_rows, part_rows, cluster; const char* what; };
  const Cfg cfgs[] = {
      {0, 384, 128, 1, "distinct rows per CTA, 48 KB stage (3 x 128-row boxes)"},
      {0, 256, 256, 1, "distinct rows per CTA, 32 KB stage (1 x 256-row box)"},
      {1, 384, 128, 1, "same tile for all CTAs, 48 KB stage"},
      {1, 256, 256, 1, "same tile for all CTAs, 32 KB stage"},
      {2, 256, 128, 2, "cluster 2 multicast, 32 KB stage (each CTA fetches 128 rows)"},
      {2, 256, 64, 4, "cluster 4 multicast, 32 KB stage (each CTA fetches 64 rows)"},
      {2, 384, 96, 4, "cluster 4 multicast, 48 KB stage (each CTA fetches 96 rows)"},
  };
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {64, static_cast<cuuint64_t>(rows_total)};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, static_cast<cuuint32_t>(c.part_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int grid = (sms / c.cluster) * c.cluster;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = STAGES * c.box_rows * 128 + 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) {
      cudaError_t e = cudaLaunchKernelEx(&cfg, ingest_kernel, tm, c.mode, c.box_rows, c.part_rows, iters, rows_total, dcyc);
      if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[148];
    cudaMemcpy(h, dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; ++i) avg += h[i];
    avg /= grid;
    const double bytes = static_cast<double>(iters) * c.box_rows * 128;
    printf("%-72s grid %3d: %7.1f B/clk/SM into smem (%.0f cycles per %d KB stage)\n", c.what, grid, bytes / avg, avg / iters, c.box_rows * 128 / 1024);
  }
  return 0;
}
