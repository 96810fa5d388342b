// Shared host-side helpers: error reporting for the C ABI and launch checks.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/jpdse_b200.h"

namespace jpdse {

// thread-local message behind jpdse_last_error()
char* last_error_buffer();
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();  // clear the sticky launch-config error so later calls are not poisoned
    return fail(JPDSE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return JPDSE_OK;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace jpdse
