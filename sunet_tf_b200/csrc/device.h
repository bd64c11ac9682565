// Per-device host state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are properties of a DEVICE, not of
// the process: a second device in the same process (model.to("cuda:1"), nn.DataParallel replicas) needs its own opt-in and its
// own grid size.  Both are keyed on cudaGetDevice() here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace sunet {

inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

// SM count of the current device (cached per device)
inline int device_sms() {
  static std::atomic<int> cache[64];
  const int d = current_device() & 63;
  int n = cache[d].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d) != cudaSuccess || n <= 0) n = 148;
    cache[d].store(n, std::memory_order_relaxed);
  }
  return n;
}

// `static DeviceOnce once; if (once.need()) { ...cudaFuncSetAttribute...; once.done(); }` runs the body once per device.
// A race between two host threads on the same device only repeats the (idempotent) body.
struct DeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool need() const { return ((mask.load(std::memory_order_acquire) >> (current_device() & 63)) & 1ull) == 0; }
  void done() { mask.fetch_or(1ull << (current_device() & 63), std::memory_order_release); }
};

}  // namespace sunet
