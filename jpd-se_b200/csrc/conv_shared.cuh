// Declarations shared by the convolution translation units (conv_igemm.cu, conv_rowstat.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "common.cuh"

namespace jpdse {

// which kernel executes a conv descriptor (decided once, in conv_geom; packing follows it)
enum ConvPath {
  kPathIgemm = 0,    // generic per-tap implicit GEMM (conv_igemm.cu)
  kPathRowHead = 1,  // row-stationary 7x7, (kw,cout) polyphase columns, bias+tanh (conv_rowstat.cu)
  kPathRowStem = 2,  // row-stationary 7x7, window-K, 32-channel splits, raw+stats (conv_rowstat.cu)
  kPathConvtFused = 3,  // ConvTranspose with all four output phases per tile and halo-shared A boxes (conv_convt.cu)
  kPathPair = 4         // 3x3 stride-1 conv on a CTA pair, tcgen05 cta_group::2, 256 x 256 tiles (conv_pair.cu)
};

struct ConvGeom {
  int out_h, out_w;    // output spatial dims
  int gemm_h, gemm_w;  // pixel grid the GEMM M dimension runs over (== out dims except ConvT: input dims)
  int ktot;            // packed K per packed row
  int rows;            // packed rows
  int bn;              // igemm N tile
  int cpt;             // igemm: 64-element chunks per tap
  int path;            // ConvPath
};

int conv_geom(const jpdse_conv_desc* d, ConvGeom* g);

// bf16 tensor map, 128-byte swizzle, zero OOB fill. dims/box innermost first; strides (bytes) for dims 1..rank-1.
int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box);

bool convt_fused_applicable(const jpdse_conv_desc* d);
int convt_fused_pack(const jpdse_conv_desc* d, const float* w, void* w_packed, cudaStream_t stream);
int convt_fused_forward(const jpdse_conv_desc* d, const void* x, const void* w_packed, void* y, double* stats,
                        cudaStream_t stream);

bool pair_conv_applicable(const jpdse_conv_desc* d);
int pair_conv_forward(const jpdse_conv_desc* d, const void* x, const void* w_packed, void* y, double* stats,
                      cudaStream_t stream);

int rowconv_forward(const jpdse_conv_desc* d, bool head, const void* x, const void* w_packed, const float* bias, void* y,
                    double* stats, cudaStream_t stream);

// Column sums over the 32 lanes of a warp for 32 per-lane values: after the butterfly lane l holds
// sum over lanes of v[l]. 31 shuffles instead of 32*5.
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = upper ? v[j] : v[j + s];
      const float keep = upper ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

}  // namespace jpdse
