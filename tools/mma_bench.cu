// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and of how the
// accumulators / A tiles are reused. Shared memory is left uninitialised-but-finite (zeros); only timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_bench tools/mma_bench.cu -I jpd-se_b200/csrc
#include <cstdio>
#include <cstdlib>
#include "ptx.cuh"

using namespace jpdse;

// mode 0: all MMAs accumulate into ONE tmem region, same A/B tile (pure issue/throughput)
// mode 1: rotate over 7 accumulator slots per A k-step (the row-stationary pattern: same A, different B + D)
// mode 2: rotate A tiles (4 K-steps of a 16 KB tile), one accumulator
template <int N>
__global__ void __launch_bounds__(128, 1) mma_bench_kernel(int iters, int mode, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (120 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc<1>(&tmem_slot, 512);
    tmem_relinquish<1>();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint32_t sa = smem_u32(smem);
    const uint32_t sb = sa + 16384;
    const uint64_t adesc = umma_smem_desc_sw128(sa);
    const uint64_t bdesc = umma_smem_desc_sw128(sb);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem_base, adesc + k * 2, bdesc + k * 2, idesc, 1u);
      } else if (mode == 1) {
#pragma unroll
        for (int s = 0; s < 7; ++s)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16<1>(tmem_base + s * N, adesc + k * 2, bdesc + (s * N * 128 >> 4) + k * 2, idesc, 1u);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem_base, adesc + ((it & 1) * 16384 >> 4) + k * 2, bdesc + k * 2, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

template <int N>
void run(int mode, int grid) {
  const int iters = 2000;
  long long* d;
  cudaMalloc(&d, sizeof(long long) * grid);
  cudaFuncSetAttribute(mma_bench_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  mma_bench_kernel<N><<<grid, 128, 128 * 1024>>>(iters, mode, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("N=%d mode=%d: %s\n", N, mode, cudaGetErrorString(e));
    exit(1);
  }
  long long h[256];
  cudaMemcpy(h, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const int per_iter = mode == 1 ? 28 : 4;
  printf("N=%3d mode=%d grid=%3d: %8.1f cycles/MMA  (ideal %d)\n", N, mode, grid, double(mx) / (double(iters) * per_iter),
         128 * N / 256);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode != 1) run<256>(mode, grid);
      if (mode != 1) run<128>(mode, grid);
      run<64>(mode, grid);
      run<32>(mode, grid);
      run<16>(mode, grid);
    }
  }
  return 0;
}
