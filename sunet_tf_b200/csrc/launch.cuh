// Kernel launch helper: programmatic dependent launch (PDL) on the forward path.
//
// The SUNet forward is ~330 strictly dependent launches per batch of 64; with ordinary stream ordering every launch pays
// the grid drain of its predecessor plus the launch latency plus its own prologue (barrier init, TMEM allocation,
// descriptor prefetch) - about 5 us per tcgen05 GEMM launch (measured with csrc/tests/test_gemm).  With
// cudaLaunchAttributeProgrammaticStreamSerialization the CTAs of kernel N+1 are scheduled as soon as the CTAs of kernel N
// exit, run their prologue, and block in griddepcontrol.wait until ALL of kernel N has completed and flushed.
//
// Contract for every kernel launched through launch_pdl(): call pdl_wait() (ptx.cuh) before the first global-memory access
// (reads of the predecessor's output AND writes that could race with its reads), then pdl_launch_dependents().
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include <utility>

namespace sunet {

inline bool pdl_enabled() {
  static const bool on = getenv("SUNET_NO_PDL") == nullptr;
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace sunet
