// common.cuh — error plumbing shared by the translation units of libbsvd_b200.so.
#pragma once
#include <cuda_runtime.h>

namespace bsvd {

// defined in bsvd_capi.cu: records the message bsvd_last_error() returns, yields 1
int fail(const char* fmt, ...);
#define CUDA_TRY(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

}  // namespace bsvd
