// Experiment: tcgen05.mma with MN-major (pixel-major / NHWC) operands, as needed by the weight-gradient GEMM
//   dW[co][ci] = sum_pixels dY[pixel][co] * X[pixel][ci]
// where both operands sit in shared memory exactly as a TMA box {64 channels, K pixels} with SWIZZLE_128B writes them:
// rows = pixels (K), 128 bytes = 64 channels (M or N) per row.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I jpd-se_b200/csrc -o tools/umma_mnmajor_test.bin tools/umma_mnmajor_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include "ptx.cuh"
using namespace jpdse;

constexpr int kK = 64;   // pixels per tile
constexpr int kM = 128;  // two 64-channel atoms
constexpr int kN = 64;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1) mn_kernel(int variant, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sa = smem;              // 2 atoms x (64 rows x 128 B)
  uint8_t* sb = smem + 16 * 1024;  // 1 atom
  // A[k][m] = ((k * 3 + m * 5) % 19 - 9) / 8 ; B[k][n] = ((k * 7 + n) % 11 - 5) / 4
  for (int i = threadIdx.x; i < kK * kM; i += blockDim.x) {
    const int k = i / kM, m = i % kM;
    const float v = float((k * 3 + m * 5) % 19 - 9) / 8.f;
    const int atom = m / 64, mm = m % 64, chunk = mm / 8, within = mm % 8;
    reinterpret_cast<__nv_bfloat16*>(sa + atom * 8192 + k * 128 + ((chunk ^ (k & 7)) * 16))[within] = __float2bfloat16(v);
  }
  for (int i = threadIdx.x; i < kK * kN; i += blockDim.x) {
    const int k = i / kN, n = i % kN;
    const float v = float((k * 7 + n) % 11 - 5) / 4.f;
    const int chunk = n / 8, within = n % 8;
    reinterpret_cast<__nv_bfloat16*>(sb + k * 128 + ((chunk ^ (k & 7)) * 16))[within] = __float2bfloat16(v);
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
    // variant 0: LBO = atom pitch (8192), SBO = 1024 ; variant 1: swapped roles
    const uint32_t lbo = variant == 0 ? 8192u : 1024u, sbo = variant == 0 ? 1024u : 8192u;
    const uint64_t adesc = desc_mn_sw128(smem_u32(sa), lbo, sbo);
    const uint64_t bdesc = desc_mn_sw128(smem_u32(sb), lbo, sbo);
    const uint32_t idesc = umma_idesc_bf16(kM, kN) | (1u << 15) | (1u << 16);  // A and B MN-major
    for (int k = 0; k < kK / 16; ++k)  // 16 pixel rows per MMA = 2048 bytes
      umma_bf16<1>(tmem_base, adesc + static_cast<uint64_t>((k * 2048) >> 4), bdesc + static_cast<uint64_t>((k * 2048) >> 4),
                   idesc, k != 0);
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
  cudaMalloc(&d, kM * kN * sizeof(float));
  float* h = (float*)malloc(kM * kN * sizeof(float));
  cudaFuncSetAttribute(mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(d, 0, kM * kN * sizeof(float));
    mn_kernel<<<1, 128, 64 * 1024>>>(variant, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(h, d, kM * kN * sizeof(float), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < kM; ++m)
      for (int n = 0; n < kN; ++n) {
        double ref = 0;
        for (int k = 0; k < kK; ++k) ref += double(float((k * 3 + m * 5) % 19 - 9) / 8.f) * double(float((k * 7 + n) % 11 - 5) / 4.f);
        if (fabs(ref - h[m * kN + n]) > 1e-3) ++bad;
      }
    printf("MN-major variant %d (LBO=%s): %s (bad %d / %d); D[0][0..3] = %g %g %g %g\n", variant,
           variant == 0 ? "atom pitch, SBO=1024" : "1024, SBO=atom pitch", bad == 0 ? "CORRECT" : "wrong", bad, kM * kN, h[0],
           h[1], h[2], h[3]);
  }
  return 0;
}
