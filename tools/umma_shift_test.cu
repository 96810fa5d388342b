// Experiment: can a SWIZZLE_128B K-major UMMA operand start at a 128-byte row that is NOT 1024-byte aligned
// (i.e. a tile shifted by r rows inside a larger shared-memory tile), and which value must the descriptor's
// base_offset field (bits 49..51) carry?  Needed for "one halo tile in smem, all kw taps" convolutions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jpd-se_b200/csrc -o tools/umma_shift_test.bin tools/umma_shift_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace jpdse;

constexpr int kRowsA = 160;  // rows of the big A tile in smem
constexpr int kN = 64;

__global__ void __launch_bounds__(128, 1) shift_kernel(int shift_rows, int base_off_mode, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sa = smem;                 // kRowsA rows x 128 B, 128B-swizzled by ABSOLUTE row index
  uint8_t* sb = smem + 24 * 1024;     // 64 rows x 128 B
  // A[row][k] = ((row * 7 + k * 3) % 17 - 8) / 8 ; B[n][k] = ((n * 5 + k) % 13 - 6) / 4   (exact in bf16)
  for (int i = threadIdx.x; i < kRowsA * 64; i += blockDim.x) {
    const int row = i / 64, k = i % 64;
    const float v = float((row * 7 + k * 3) % 17 - 8) / 8.f;
    const int chunk = k / 8, within = k % 8;
    const int phys = (chunk ^ (row & 7));
    reinterpret_cast<__nv_bfloat16*>(sa + row * 128 + phys * 16)[within] = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < kN * 64; i += blockDim.x) {
    const int row = i / 64, k = i % 64;
    const float v = float((row * 5 + k) % 13 - 6) / 4.f;
    const int chunk = k / 8, within = k % 8;
    const int phys = (chunk ^ (row & 7));
    reinterpret_cast<__nv_bfloat16*>(sb + row * 128 + phys * 16)[within] = __float2bfloat16(v);
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc<1>(&tmem_slot, 64);
    tmem_relinquish<1>();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(sa) + shift_rows * 128;
    uint64_t adesc = umma_smem_desc_sw128(a_addr);
    if (base_off_mode == 1) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sb));
    for (int k = 0; k < 4; ++k) umma_bf16<1>(tmem_base, adesc + k * 2, bdesc + k * 2, umma_idesc_bf16(128, kN), k != 0);
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int ch = 0; ch < kN / 32; ++ch) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * kN + ch * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tmem_base, 64);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * kN * sizeof(float));
  float* h = (float*)malloc(128 * kN * sizeof(float));
  cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int mode = 0; mode < 2; ++mode) {
    for (int shift : {0, 1, 2, 3, 5, 8, 9, 17}) {
      cudaMemset(d, 0, 128 * kN * sizeof(float));
      shift_kernel<<<1, 128, 64 * 1024>>>(shift, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("shift %d mode %d: CUDA error %s\n", shift, mode, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(h, d, 128 * kN * sizeof(float), cudaMemcpyDeviceToHost);
      int bad = 0;
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < kN; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k)
            ref += double(float(((m + shift) * 7 + k * 3) % 17 - 8) / 8.f) * double(float((n * 5 + k) % 13 - 6) / 4.f);
          const double err = fabs(ref - h[m * kN + n]);
          if (err > 1e-3) ++bad;
          if (err > maxerr) maxerr = err;
        }
      printf("shift %2d rows, base_offset %s: %s (bad %d / %d, max err %.3g)\n", shift, mode ? "=(addr>>7)&7" : "=0",
             bad == 0 ? "CORRECT" : "wrong", bad, 128 * kN, maxerr);
    }
  }
  return 0;
}
