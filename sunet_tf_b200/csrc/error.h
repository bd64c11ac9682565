// Error plumbing for the C-ABI: 0 = OK, <0 = argument/shape/alignment error, >0 = cudaError_t.
// Messages are thread-local and fetched through sunet_last_error(); nothing throws across the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

namespace sunet {

enum {
  SUNET_OK = 0,
  SUNET_E_ARG = -1,
  SUNET_E_SHAPE = -2,
  SUNET_E_ALIGN = -3,
  SUNET_E_WORKSPACE = -4,
  SUNET_E_STATE = -5,
  SUNET_E_DRIVER = -6,
};

char* last_error_buf();  // thread-local, 512 bytes

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define SUNET_CUDA(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::sunet::fail(static_cast<int>(_e), "%s:%d %s: %s", __FILE__, __LINE__, #expr,        \
                           cudaGetErrorString(_e));                                               \
  } while (0)

#define SUNET_TRY(expr)        \
  do {                         \
    int _rc = (expr);          \
    if (_rc != 0) return _rc;  \
  } while (0)

#define SUNET_CHECK_LAUNCH() SUNET_CUDA(cudaGetLastError())

}  // namespace sunet
