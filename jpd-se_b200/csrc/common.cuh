// Shared host-side helpers: error reporting for the C ABI and launch checks.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
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

// SM count of the CURRENT device (cached per device: a process may touch several GPUs)
inline int num_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (n[dev] == 0) {
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    v = v > 0 ? v : 148;
    // JPDSE_RESERVE_SMS=n (experiment; read once per process): size every grid for n fewer SMs, so that a concurrent
    // library kernel (NCCL under the overlapped gradient all-reduce) has SMs of its own instead of delaying the
    // statically scheduled persistent CTAs that would have run there
    const char* e = getenv("JPDSE_RESERVE_SMS");
    const int r = e ? atoi(e) : 0;
    if (r > 0 && r < v) v -= r;
    n[dev] = v;
  }
  return n[dev];
}

// JPDSE_PDL=1 (read once per process): the ResnetBlock conv and the InstanceNorm apply kernel are launched with the
// programmatic-stream-serialization attribute, so each one's CTAs are placed and run their prologue while the previous
// kernel drains (the kernels call griddepcontrol.wait before they touch global memory)
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JPDSE_PDL");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// One-time per-DEVICE setup (cudaFuncSetAttribute applies to the current device only). The design is one process
// per GPU, but a process that touches several devices must not inherit "already configured" from the first.
struct DeviceOnce {
  unsigned long long mask = 0;
  int dev = 0;
  bool first_use() {
    cudaGetDevice(&dev);
    return dev < 0 || dev >= 64 || !((mask >> dev) & 1ull);
  }
  void done() {
    if (dev >= 0 && dev < 64) mask |= 1ull << dev;
  }
};

}  // namespace jpdse
