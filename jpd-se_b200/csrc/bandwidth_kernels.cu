// HBM-bound kernels of the JPD-SE hot path: input build (one-hot + edges + concat + reflect pad),
// InstanceNorm apply (+ReLU, +residual, +reflect pad), layout converts and the ctu/quantizers
// forward passes. All are coalesced, 16-byte-vectorised, grid sized from the SM count.
#include <cuda_bf16.h>

#include <cstdint>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace jpdse {

__device__ __forceinline__ int reflect_index(int i, int n) {
  // ReflectionPad2d: -k -> k, (n-1+k) -> (n-1-k); edge sample not repeated
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------ input build
// label id with the reference's `.long()` semantics (truncate toward zero); -1 = not a valid channel
__device__ __forceinline__ int load_label(const void* p, int dtype, size_t i, int num_labels) {
  long long v;
  if (dtype == 0) {
    const float f = static_cast<const float*>(p)[i];
    if (!(f > -9.0e18f && f < 9.0e18f)) return -1;  // NaN / inf / out of int64 range
    v = static_cast<long long>(f);
  } else if (dtype == 1) {
    v = static_cast<const uint8_t*>(p)[i];
  } else {
    v = static_cast<const long long*>(p)[i];
  }
  return (v >= 0 && v < num_labels) ? static_cast<int>(v) : -1;
}

template <typename T>
__device__ __forceinline__ bool edge_at(const T* t, int h, int w, int H, int W) {
  // Pix2PixHDModel.get_edges: both pixels of every horizontally / vertically differing pair
  const T* row = t + static_cast<size_t>(h) * W;
  const T c = row[w];
  bool e = false;
  if (w > 0) e |= (c != row[w - 1]);
  if (w < W - 1) e |= (c != row[w + 1]);
  if (h > 0) e |= (c != row[w - W]);
  if (h < H - 1) e |= (c != row[w + W]);
  return e;
}

__device__ __forceinline__ bool load_edge(const void* inst, int dtype, size_t img_off, int h, int w, int H, int W) {
  switch (dtype) {
    case 0: return edge_at(static_cast<const int*>(inst) + img_off, h, w, H, W);
    case 1: return edge_at(static_cast<const short*>(inst) + img_off, h, w, H, W);
    case 2: return edge_at(static_cast<const long long*>(inst) + img_off, h, w, H, W);
    default: return edge_at(static_cast<const float*>(inst) + img_off, h, w, H, W);
  }
}

// image pixel: float32 (already normalised) or uint8 with the loader's normalisation fused in --
// ToTensor (x / 255) then Normalize ((x - mean) / std), float32, IEEE division (ctu/data/base_dataset.py transforms)
struct ImgNorm {
  int u8;
  float mean[3], std[3];
};
__device__ __forceinline__ float load_px(const void* image, const ImgNorm& nm, size_t idx, int c) {
  if (!nm.u8) return static_cast<const float*>(image)[idx];
  const float v = static_cast<float>(static_cast<const uint8_t*>(image)[idx]);
  return __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), nm.mean[c]), nm.std[c]);
}

constexpr int kBuildThreads = 256;

// One thread per padded output pixel; rows are staged in shared memory so the NHWC write is a
// contiguous stream of 16-byte vectors.
__global__ void __launch_bounds__(kBuildThreads)
build_input_nhwc_kernel(const void* __restrict__ label, int label_dtype, const void* __restrict__ inst, int inst_dtype,
                        const void* __restrict__ image, ImgNorm nm, int B, int H, int W, int num_labels, int pad, int c_pad,
                        __nv_bfloat16* __restrict__ out, int* __restrict__ bad) {
  extern __shared__ uint4 s_rows[];  // [kBuildThreads][c_pad/8]
  // uint8 images: the normalised value of each of the 256 byte values, per channel, computed ONCE per block with the
  // loader's exact float32 formula (two IEEE divisions) -- the per-pixel work is then a shared-memory lookup
  __shared__ float s_lut[3][256];
  if (nm.u8) {
    for (int i = threadIdx.x; i < 3 * 256; i += kBuildThreads) {
      const int c = i >> 8;
      s_lut[c][i & 255] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), 255.0f), nm.mean[c]), nm.std[c]);
    }
    __syncthreads();
  }
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const size_t total = static_cast<size_t>(B) * Hp * Wp;
  const int vpp = c_pad / 8;
  for (size_t base = static_cast<size_t>(blockIdx.x) * kBuildThreads; base < total;
       base += static_cast<size_t>(gridDim.x) * kBuildThreads) {
    const size_t pix = base + threadIdx.x;
    if (pix < total) {
      const int pw = static_cast<int>(pix % Wp);
      const int ph = static_cast<int>((pix / Wp) % Hp);
      const int b = static_cast<int>(pix / (static_cast<size_t>(Wp) * Hp));
      const int h = reflect_index(ph - pad, H), w = reflect_index(pw - pad, W);
      const size_t img_off = static_cast<size_t>(b) * H * W;
      const size_t src = img_off + static_cast<size_t>(h) * W + w;
      const int lab = load_label(label, label_dtype, src, num_labels);
      if (lab < 0 && bad != nullptr && ph >= pad && ph < H + pad && pw >= pad && pw < W + pad) atomicAdd(bad, 1);
      const float edge = load_edge(inst, inst_dtype, img_off, h, w, H, W) ? 1.f : 0.f;
      const size_t plane = static_cast<size_t>(H) * W;
      const size_t ip = static_cast<size_t>(b) * 3 * plane + static_cast<size_t>(h) * W + w;
      float r, g, bl;
      if (nm.u8) {
        const uint8_t* iu = static_cast<const uint8_t*>(image);
        r = s_lut[0][iu[ip]];
        g = s_lut[1][iu[ip + plane]];
        bl = s_lut[2][iu[ip + 2 * plane]];
      } else {
        r = load_px(image, nm, ip, 0), g = load_px(image, nm, ip + plane, 1), bl = load_px(image, nm, ip + 2 * plane, 2);
      }
      for (int v = 0; v < vpp; ++v) {
        uint32_t wds[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float e[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int ch = v * 8 + j * 2 + k;
            float x = 0.f;
            if (ch < num_labels) x = (ch == lab) ? 1.f : 0.f;
            else if (ch == num_labels) x = edge;
            else if (ch == num_labels + 1) x = r;
            else if (ch == num_labels + 2) x = g;
            else if (ch == num_labels + 3) x = bl;
            e[k] = x;
          }
          wds[j] = pack_bf16x2(e[0], e[1]);
        }
        s_rows[threadIdx.x * vpp + v] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
      }
    }
    __syncthreads();
    const size_t nvec_total = total * vpp;
    const size_t vbase = base * vpp;
    uint4* dst = reinterpret_cast<uint4*>(out);
    for (int i = threadIdx.x; i < kBuildThreads * vpp; i += kBuildThreads)
      if (vbase + i < nvec_total) dst[vbase + i] = s_rows[i];
    __syncthreads();
  }
}

// Reference layout: float32 NCHW `input_concat` (B, num_labels + 4, H, W); thread per pixel, plane-coalesced.
__global__ void __launch_bounds__(256)
build_input_nchw_kernel(const void* __restrict__ label, int label_dtype, const void* __restrict__ inst, int inst_dtype,
                        const void* __restrict__ image, ImgNorm nm, int B, int H, int W, int num_labels, float* __restrict__ out,
                        int* __restrict__ bad, int count_bad) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  const int C = num_labels + 4;
  for (size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(pix / plane);
    const size_t hw = pix % plane;
    const int h = static_cast<int>(hw / W), w = static_cast<int>(hw % W);
    const int lab = load_label(label, label_dtype, pix, num_labels);
    if (lab < 0 && bad != nullptr && count_bad) atomicAdd(bad, 1);
    const float edge = load_edge(inst, inst_dtype, static_cast<size_t>(b) * plane, h, w, H, W) ? 1.f : 0.f;
    float* o = out + static_cast<size_t>(b) * C * plane + hw;
    for (int ch = 0; ch < num_labels; ++ch) o[ch * plane] = (ch == lab) ? 1.f : 0.f;
    o[num_labels * plane] = edge;
    const size_t ip = static_cast<size_t>(b) * 3 * plane + hw;
    o[(num_labels + 1) * plane] = load_px(image, nm, ip, 0);
    o[(num_labels + 2) * plane] = load_px(image, nm, ip + plane, 1);
    o[(num_labels + 3) * plane] = load_px(image, nm, ip + 2 * plane, 2);
  }
}

// ------------------------------------------------------------------------------------------ discriminator input from ids
// cat(one-hot(label), edge(instance), image) [+ AvgPool2d(3, 2, 1, count_include_pad=False)] for the PatchGAN, straight
// from the class / instance ids (pix2pixHD_model.py:376-396 + :451-460 + networks.py:387) -- 20 B read per pixel instead
// of the 156 B of the reference's float32 (B,39,H,W) tensor. One thread per OUTPUT pixel builds the c_pad channels
// (8 x 16-byte stores); the one-hot / edge channels are shared by two images (fake and real): both outputs are written in
// one pass. The pooled one-hot is count / taps and the pooled image sum / taps in the tap order of ATen's kernel.
constexpr int kDIdsPixels = 128;  // output pixels (of one row) per block
constexpr int kDIdsMaxVec = 8;    // c_pad <= 64

__global__ void __launch_bounds__(kDIdsPixels)
d_input_ids_kernel(const void* __restrict__ label, int label_dtype, const void* __restrict__ inst, int inst_dtype,
                   const float* __restrict__ img_a, __nv_bfloat16* __restrict__ out_a, const float* __restrict__ img_b,
                   __nv_bfloat16* __restrict__ out_b, int B, int H, int W, int Ho, int Wo, int L, int c_pad, int pool, int out_pad) {
  // One thread per output pixel builds its c_pad channels (for both images) as rows of shared memory; the block then
  // writes its 128 pixels -- one contiguous run of 128 * c_pad * 2 bytes per image -- with coalesced 16-byte stores, like
  // build_input_nhwc_kernel. (Per-thread global stores of 16 bytes at a 128-byte stride ran at 1.4 TB/s; a thread per
  // (pixel, vector) without the staging at 0.8 TB/s.) grid = (row pieces, output rows, images).
  __shared__ uint4 s_a[kDIdsPixels * (kDIdsMaxVec + 1)];
  __shared__ uint4 s_b[kDIdsPixels * (kDIdsMaxVec + 1)];
  (void)B;
  const int vpp = c_pad >> 3;
  const int pitch = vpp + 1;  // odd pitch in 16-byte units: conflict-free row writes
  const size_t plane = static_cast<size_t>(H) * W;
  const int Wst = Wo + 2 * out_pad;
  const size_t ost = static_cast<size_t>(Ho + 2 * out_pad) * Wst;
  const int oy = blockIdx.y, b = blockIdx.z;
  const int ox0 = blockIdx.x * kDIdsPixels;
  const int ox = ox0 + threadIdx.x;
  if (ox < Wo) {
    int labs[9];
    float edge = 0.f, ia[3] = {0.f, 0.f, 0.f}, ib[3] = {0.f, 0.f, 0.f};
    int cnt = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) labs[t] = -2;  // -2: tap not present
    const size_t ibase = static_cast<size_t>(b) * 3 * plane;
    if (!pool) {
      const size_t hw = static_cast<size_t>(oy) * W + ox;
      labs[0] = load_label(label, label_dtype, static_cast<size_t>(b) * plane + hw, L);
      edge = load_edge(inst, inst_dtype, static_cast<size_t>(b) * plane, oy, ox, H, W) ? 1.f : 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ia[c] = __ldg(img_a + ibase + c * plane + hw);
        if (img_b != nullptr) ib[c] = __ldg(img_b + ibase + c * plane + hw);
      }
      cnt = 1;
    } else {
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int y = 2 * oy + dy;
        if (y < 0 || y >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int x = 2 * ox + dx;
          if (x < 0 || x >= W) continue;
          const size_t hw = static_cast<size_t>(y) * W + x;
          labs[(dy + 1) * 3 + dx + 1] = load_label(label, label_dtype, static_cast<size_t>(b) * plane + hw, L);
          edge += load_edge(inst, inst_dtype, static_cast<size_t>(b) * plane, y, x, H, W) ? 1.f : 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            ia[c] += __ldg(img_a + ibase + c * plane + hw);
            if (img_b != nullptr) ib[c] += __ldg(img_b + ibase + c * plane + hw);
          }
          ++cnt;
        }
      }
    }
    const float fc = static_cast<float>(cnt);
    edge = edge / fc;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ia[c] = ia[c] / fc;
      ib[c] = ib[c] / fc;
    }
    for (int k = 0; k < vpp; ++k) {
      float va[8], vb[8];
      bool differs = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 8 * k + j;
        float v = 0.f, w2 = 0.f;
        if (c < L) {
          if (!pool) {
            v = w2 = (labs[0] == c) ? 1.f : 0.f;
          } else {
            int n = 0;
#pragma unroll
            for (int t = 0; t < 9; ++t) n += (labs[t] == c) ? 1 : 0;
            v = w2 = n ? static_cast<float>(n) / fc : 0.f;  // at most 9 of the L channels take the division
          }
        } else if (c == L) {
          v = w2 = edge;
        } else if (c < L + 4) {
          v = ia[c - L - 1];
          w2 = ib[c - L - 1];
          differs = true;
        }
        va[j] = v;
        vb[j] = w2;
      }
      const uint4 pa = make_uint4(pack_bf16x2(va[0], va[1]), pack_bf16x2(va[2], va[3]), pack_bf16x2(va[4], va[5]), pack_bf16x2(va[6], va[7]));
      s_a[threadIdx.x * pitch + k] = pa;
      s_b[threadIdx.x * pitch + k] =
          differs ? make_uint4(pack_bf16x2(vb[0], vb[1]), pack_bf16x2(vb[2], vb[3]), pack_bf16x2(vb[4], vb[5]), pack_bf16x2(vb[6], vb[7])) : pa;
    }
  }
  __syncthreads();
  const int npx = Wo - ox0 < kDIdsPixels ? Wo - ox0 : kDIdsPixels;
  const size_t o = (static_cast<size_t>(b) * ost + static_cast<size_t>(oy + out_pad) * Wst + ox0 + out_pad) * c_pad;
  uint4* da = reinterpret_cast<uint4*>(out_a + o);
  uint4* db = out_b != nullptr ? reinterpret_cast<uint4*>(out_b + o) : nullptr;
  for (int i = threadIdx.x; i < npx * vpp; i += kDIdsPixels) {
    const int px = i / vpp, k = i - px * vpp;
    da[i] = s_a[px * pitch + k];
    if (db != nullptr) db[i] = s_b[px * pitch + k];
  }
}

// ------------------------------------------------------------------------------------------ InstanceNorm apply
constexpr int kNormThreads = 256;
constexpr int kNormIters = 16;

template <bool kRelu, bool kResidual, int kIters = kNormIters>
__global__ void __launch_bounds__(kNormThreads)
instnorm_apply_kernel(const __nv_bfloat16* __restrict__ raw, const double* __restrict__ stats,
                      const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out, int H, int W, int C,
                      int pad, float eps) {
  const int vpp = C >> 3;                 // 16-byte vectors per pixel
  const int ppi = kNormThreads / vpp;     // pixels per block iteration
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int npix = Hp * Wp;
  // PDL (JPDSE_PDL=1): this CTA may have been placed while the conv that produces `raw` / `stats` was still running
  grid_dep_launch_dependents();
  grid_dep_wait();

  float mean[8], rstd[8];
  {
    const double inv_n = 1.0 / (static_cast<double>(H) * W);
    const double* st = stats + (static_cast<size_t>(b) * C + vec * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double m = st[2 * j] * inv_n;
      double var = st[2 * j + 1] * inv_n - m * m;  // biased variance, fp64 so E[x^2]-E[x]^2 does not cancel
      if (var < 0.0) var = 0.0;
      mean[j] = static_cast<float>(m);
      rstd[j] = rsqrtf(static_cast<float>(var) + eps);  // fp32 like the reference's invstd
    }
  }
  const uint4* raw4 = reinterpret_cast<const uint4*>(raw) + static_cast<size_t>(b) * H * W * vpp;
  const uint4* res4 = reinterpret_cast<const uint4*>(residual) + static_cast<size_t>(b) * npix * vpp;
  uint4* out4 = reinterpret_cast<uint4*>(out) + static_cast<size_t>(b) * npix * vpp;
  if (psub >= ppi) return;  // only when vpp does not divide the block (never for power-of-two C)
  auto src_of = [&](int pp) -> size_t {
    if (pad == 0) return static_cast<size_t>(pp);  // output pixel == raw pixel, no index arithmetic
    const int ph = pp / Wp, pw = pp - ph * Wp;
    return static_cast<size_t>(reflect_index(ph - pad, H)) * W + reflect_index(pw - pad, W);
  };
  auto finish = [&](const uint4& x, const uint4& rs, int pp) {
    const uint32_t xw[4] = {x.x, x.y, x.z, x.w};
    const uint32_t rw[4] = {rs.x, rs.y, rs.z, rs.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&xw[j]);
      float lo = (__low2float(xv) - mean[2 * j]) * rstd[2 * j];
      float hi = (__high2float(xv) - mean[2 * j + 1]) * rstd[2 * j + 1];
      if (kRelu) {
        lo = fmaxf(lo, 0.f);
        hi = fmaxf(hi, 0.f);
      }
      if (kResidual) {
        const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
        lo += __low2float(rv);
        hi += __high2float(rv);
      }
      ow[j] = pack_bf16x2(lo, hi);
    }
    out4[static_cast<size_t>(pp) * vpp + vec] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  };
  // grid-stride over pixel blocks: the statistics prologue above (16 dependent fp64 loads per thread) runs once per
  // CTA and the grid is one resident wave, instead of once per 16 pixel groups.
  // Loads are issued FOUR pixels at a time before anything is consumed: with a bounds check (a possible `break`) between
  // the iterations the compiler cannot hoist the next pixel's loads above the current pixel's arithmetic, and every
  // thread then has a single 16-byte load in flight (the kernel sat at 0.46-0.68 of the HBM peak).
  for (int pix0 = blockIdx.x * (ppi * kIters); pix0 < npix; pix0 += gridDim.x * (ppi * kIters)) {
    if (pix0 + ppi * kIters <= npix) {
#pragma unroll
      for (int it0 = 0; it0 < kIters; it0 += 4) {
        uint4 x[4], rs[4];
        int pp[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pp[u] = pix0 + (it0 + u) * ppi + psub;
          x[u] = __ldg(raw4 + src_of(pp[u]) * vpp + vec);
          rs[u] = make_uint4(0, 0, 0, 0);
          if (kResidual) rs[u] = __ldg(res4 + static_cast<size_t>(pp[u]) * vpp + vec);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) finish(x[u], rs[u], pp[u]);
      }
    } else {
      for (int it = 0; it < kIters; ++it) {
        const int pp = pix0 + it * ppi + psub;
        if (pp >= npix) break;
        const uint4 x = __ldg(raw4 + src_of(pp) * vpp + vec);
        uint4 rs = make_uint4(0, 0, 0, 0);
        if (kResidual) rs = __ldg(res4 + static_cast<size_t>(pp) * vpp + vec);
        finish(x, rs, pp);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ layout converts
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C,
                                             int H, int W, int pad, int c_pad) {
  // 32 channels x 32 pixels transposed through shared memory: reads coalesced along W, writes along C
  __shared__ float tile[32][33];
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const size_t npix = static_cast<size_t>(Hp) * Wp;
  const size_t p0 = static_cast<size_t>(blockIdx.x) * 32;
  {
    const size_t pp = p0 + threadIdx.x;
    const int c = c0 + threadIdx.y;
    float v = 0.f;
    if (pp < npix && c < C) {
      const int ph = static_cast<int>(pp / Wp), pw = static_cast<int>(pp % Wp);
      const int h = reflect_index(ph - pad, H), w = reflect_index(pw - pad, W);
      v = x[((static_cast<size_t>(b) * C + c) * H + h) * W + w];
    }
    tile[threadIdx.y][threadIdx.x] = v;
  }
  __syncthreads();
  {
    const size_t pp = p0 + threadIdx.y;
    const int c = c0 + threadIdx.x;
    if (pp < npix && c < c_pad)
      y[(static_cast<size_t>(b) * npix + pp) * c_pad + c] = __float2bfloat16_rn(tile[threadIdx.x][threadIdx.y]);
  }
}

__global__ void nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int B, int C,
                                             int H, int W) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const size_t npix = static_cast<size_t>(H) * W;
  const size_t p0 = static_cast<size_t>(blockIdx.x) * 32;
  {
    const size_t pp = p0 + threadIdx.y;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (pp < npix && c < C) v = __bfloat162float(x[(static_cast<size_t>(b) * npix + pp) * C + c]);
    tile[threadIdx.y][threadIdx.x] = v;
  }
  __syncthreads();
  {
    const size_t pp = p0 + threadIdx.x;
    const int c = c0 + threadIdx.y;
    if (pp < npix && c < C) y[(static_cast<size_t>(b) * C + c) * npix + pp] = tile[threadIdx.x][threadIdx.y];
  }
}

// ------------------------------------------------------------------------------------------ quantisers
enum { kOpRound = 0, kOpSign = 1, kOpSoftSign = 2 };

__device__ __forceinline__ float quant_op(int op, float x, float u) {
  if (op == kOpRound) return rintf(x);  // round-half-to-even == torch.round
  if (op == kOpSign) return static_cast<float>((x > 0.f) - (x < 0.f));  // torch.sign: NaN -> 0, -0 -> 0
  // SoftSignFunction.forward: masks are evaluated on the input, x stays where both fail (NaN)
  const float t = (1.f - x) / 2.f;
  float y = x;
  if (t <= u) y = 1.f;
  if (t > u) y = -1.f;
  return y;
}

template <int kOp>
__global__ void __launch_bounds__(256)
quant_elementwise_kernel(const float* __restrict__ x, const float* __restrict__ u, float* __restrict__ y, size_t n) {
  const size_t n4 = n / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                         (kOp == kOpSoftSign ? reinterpret_cast<uintptr_t>(u) : 0)) & 15) == 0;
  if (aligned) {
    for (size_t i = tid; i < n4; i += stride) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kOp == kOpSoftSign) r = __ldg(reinterpret_cast<const float4*>(u) + i);
      float4 o;
      o.x = quant_op(kOp, a.x, r.x);
      o.y = quant_op(kOp, a.y, r.y);
      o.z = quant_op(kOp, a.z, r.z);
      o.w = quant_op(kOp, a.w, r.w);
      reinterpret_cast<float4*>(y)[i] = o;
    }
    for (size_t i = n4 * 4 + tid; i < n; i += stride) y[i] = quant_op(kOp, x[i], kOp == kOpSoftSign ? u[i] : 0.f);
  } else {
    for (size_t i = tid; i < n; i += stride) y[i] = quant_op(kOp, x[i], kOp == kOpSoftSign ? u[i] : 0.f);
  }
}

__device__ __forceinline__ uint32_t sign_bit_u8(float x) {
  // ((x + 1) / 2).astype(uint8): float -> uint8 truncation of {0, 0.5, 1}
  return static_cast<uint32_t>(static_cast<uint8_t>(static_cast<int>((x + 1.f) / 2.f)));
}

// 8 elements per thread: two 16-byte loads, one 8-byte store (one byte per element otherwise: 0.40 of the HBM peak)
__global__ void __launch_bounds__(256)
sign_to_bits_kernel(const float* __restrict__ x, uint8_t* __restrict__ y, size_t n) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(y) & 7)) == 0;
  const size_t n8 = aligned ? n / 8 : 0;
  for (size_t i = tid; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    uint2 o;
    o.x = sign_bit_u8(a.x) | (sign_bit_u8(a.y) << 8) | (sign_bit_u8(a.z) << 16) | (sign_bit_u8(a.w) << 24);
    o.y = sign_bit_u8(b.x) | (sign_bit_u8(b.y) << 8) | (sign_bit_u8(b.z) << 16) | (sign_bit_u8(b.w) << 24);
    reinterpret_cast<uint2*>(y)[i] = o;
  }
  for (size_t i = n8 * 8 + tid; i < n; i += stride) y[i] = static_cast<uint8_t>(sign_bit_u8(x[i]));
}

// S2HVQ encode: one thread per row of x; the code book lives in shared memory. kD > 0: the center size is a compile-time
// constant and the row of x sits in registers (vector loads); kD == 0: any size, the row is re-read from L1 per center
// (128 loads per row at 16 centers x 8 -- what held the generic form at 0.05 of the HBM peak). Same summation order.
template <int kD>
__global__ void __launch_bounds__(128)
s2hvq_encode_kernel(const float* __restrict__ x, const float* __restrict__ code_book, size_t rows, int d_rt, int L,
                    float sigma, float* __restrict__ scores, long long* __restrict__ index, float* __restrict__ one_hot,
                    float* __restrict__ soft) {
  extern __shared__ float s_cb[];  // [L][d]
  const int d = kD > 0 ? kD : d_rt;
  for (int i = threadIdx.x; i < L * d; i += blockDim.x) s_cb[i] = code_book[i];
  __syncthreads();
  for (size_t row = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows;
       row += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float xreg[kD > 0 ? kD : 1];
    if (kD > 0) {
      if (kD % 4 == 0) {
#pragma unroll
        for (int j = 0; j < kD / 4; ++j) {
          const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + row * kD) + j);
          xreg[4 * j] = v4.x; xreg[4 * j + 1] = v4.y; xreg[4 * j + 2] = v4.z; xreg[4 * j + 3] = v4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < kD; ++j) xreg[j] = __ldg(x + row * kD + j);
      }
    }
    const float* xr = kD > 0 ? xreg : x + row * d;
    float best = 0.f;
    int best_k = 0;
    bool best_nan = false;
    float smax = -INFINITY;  // max of -sigma*score for the soft path
    for (int k = 0; k < L; ++k) {
      float acc = 0.f;
      if (kD > 0 && kD % 4 == 0) {  // 16-byte shared-memory reads of the center (same summation order)
#pragma unroll
        for (int j4 = 0; j4 < kD / 4; ++j4) {
          const float4 c4 = reinterpret_cast<const float4*>(s_cb)[(k * kD) / 4 + j4];
          float df = xr[4 * j4] - c4.x;
          acc += df * df;
          df = xr[4 * j4 + 1] - c4.y;
          acc += df * df;
          df = xr[4 * j4 + 2] - c4.z;
          acc += df * df;
          df = xr[4 * j4 + 3] - c4.w;
          acc += df * df;
        }
      } else {
#pragma unroll
        for (int j = 0; j < d; ++j) {
          const float df = xr[j] - s_cb[k * d + j];
          acc += df * df;
        }
      }
      if (scores) scores[row * L + k] = acc;
      // torch.min: first minimal index; a NaN wins and the first NaN sticks
      if (k == 0) {
        best = acc;
        best_nan = isnan(acc);
      } else if (!best_nan && (isnan(acc) || acc < best)) {
        best = acc;
        best_k = k;
        best_nan = isnan(acc);
      }
      smax = fmaxf(smax, -sigma * acc);
    }
    if (index) index[row] = best_k;
    if (one_hot) {
      if ((L & 3) == 0 && (reinterpret_cast<uintptr_t>(one_hot) & 15) == 0) {
        float4* oh = reinterpret_cast<float4*>(one_hot + row * L);
        for (int k4 = 0; k4 < L / 4; ++k4)
          oh[k4] = make_float4(4 * k4 == best_k ? 1.f : 0.f, 4 * k4 + 1 == best_k ? 1.f : 0.f, 4 * k4 + 2 == best_k ? 1.f : 0.f,
                               4 * k4 + 3 == best_k ? 1.f : 0.f);
      } else {
        for (int k = 0; k < L; ++k) one_hot[row * L + k] = (k == best_k) ? 1.f : 0.f;
      }
    }
    if (soft) {
      float sum = 0.f;
      for (int k = 0; k < L; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < d; ++j) {
          const float df = xr[j] - s_cb[k * d + j];
          acc += df * df;
        }
        const float e = expf(-sigma * acc - smax);
        soft[row * L + k] = e;
        sum += e;
      }
      const float inv = 1.f / sum;
      for (int k = 0; k < L; ++k) soft[row * L + k] *= inv;
    }
  }
}

__global__ void __launch_bounds__(128)
s2hvq_decode_kernel(const float* __restrict__ code_raw, const float* __restrict__ code_book, size_t rows, int d, int L,
                    float* __restrict__ out, long long* __restrict__ index) {
  for (size_t row = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; row < rows;
       row += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float* cr = code_raw + row * L;
    float best = cr[0];
    int best_k = 0;
    bool best_nan = isnan(best);
    for (int k = 1; k < L; ++k) {
      const float v = cr[k];
      if (!best_nan && (isnan(v) || v > best)) {  // torch.max: first maximal index, NaN wins
        best = v;
        best_k = k;
        best_nan = isnan(v);
      }
    }
    if (index) index[row] = best_k;
    for (int j = 0; j < d; ++j) out[row * d + j] = code_book[best_k * d + j];
  }
}

static int grid_for(size_t work_items, int threads, int per_thread) {
  size_t blocks = (work_items + static_cast<size_t>(threads) * per_thread - 1) / (static_cast<size_t>(threads) * per_thread);
  const size_t cap = static_cast<size_t>(num_sms()) * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace jpdse

using namespace jpdse;

static int build_input_impl(const void* label, int label_dtype, const void* instance, int inst_dtype, const void* image,
                            const ImgNorm& nm, int batch, int height, int width, int num_labels, void* out_nhwc, int pad,
                            int c_pad, float* out_nchw, int* bad_label_count, void* stream_v) {
  if (label == nullptr || instance == nullptr || image == nullptr) return fail(JPDSE_ERR_INVALID, "build_input: NULL input");
  if (batch <= 0 || height <= 1 || width <= 1 || num_labels <= 0) return fail(JPDSE_ERR_INVALID, "build_input: bad sizes");
  if (label_dtype < 0 || label_dtype > 2 || inst_dtype < 0 || inst_dtype > 3) return fail(JPDSE_ERR_INVALID, "build_input: bad dtype code");
  if (out_nhwc == nullptr && out_nchw == nullptr) return fail(JPDSE_ERR_INVALID, "build_input: no output requested");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (out_nhwc != nullptr) {
    if (c_pad % 8 || c_pad < num_labels + 4) return fail(JPDSE_ERR_INVALID, "build_input: c_pad must be a multiple of 8 and >= num_labels+4");
    if (pad < 0 || pad >= height || pad >= width) return fail(JPDSE_ERR_INVALID, "build_input: bad pad");
    if (reinterpret_cast<uintptr_t>(out_nhwc) & 15) return fail(JPDSE_ERR_INVALID, "build_input: out_nhwc must be 16-byte aligned");
    const size_t total = static_cast<size_t>(batch) * (height + 2 * pad) * (width + 2 * pad);
    const size_t smem = static_cast<size_t>(kBuildThreads) * c_pad * 2;
    if (smem > 48 * 1024) return fail(JPDSE_ERR_UNSUPPORTED, "build_input: c_pad too large");
    const int grid = grid_for(total, kBuildThreads, 1);
    build_input_nhwc_kernel<<<grid, kBuildThreads, smem, stream>>>(label, label_dtype, instance, inst_dtype, image, nm, batch,
                                                                   height, width, num_labels, pad, c_pad,
                                                                   static_cast<__nv_bfloat16*>(out_nhwc), bad_label_count);
    int rc = check_launch("build_input_nhwc_kernel");
    if (rc != JPDSE_OK) return rc;
  }
  if (out_nchw != nullptr) {
    const size_t total = static_cast<size_t>(batch) * height * width;
    const int grid = grid_for(total, 256, 1);
    build_input_nchw_kernel<<<grid, 256, 0, stream>>>(label, label_dtype, instance, inst_dtype, image, nm, batch, height, width,
                                                      num_labels, out_nchw, bad_label_count, out_nhwc == nullptr ? 1 : 0);
    int rc = check_launch("build_input_nchw_kernel");
    if (rc != JPDSE_OK) return rc;
  }
  return JPDSE_OK;
}

extern "C" int jpdse_build_input(const void* label, int label_dtype, const void* instance, int inst_dtype,
                                 const float* image, int batch, int height, int width, int num_labels, void* out_nhwc,
                                 int pad, int c_pad, float* out_nchw, int* bad_label_count, void* stream_v) {
  ImgNorm nm;
  nm.u8 = 0;
  for (int c = 0; c < 3; ++c) nm.mean[c] = 0.f, nm.std[c] = 1.f;
  return build_input_impl(label, label_dtype, instance, inst_dtype, image, nm, batch, height, width, num_labels, out_nhwc, pad,
                          c_pad, out_nchw, bad_label_count, stream_v);
}

extern "C" int jpdse_build_input_u8(const void* label, int label_dtype, const void* instance, int inst_dtype,
                                    const uint8_t* image_u8, const float* mean, const float* std, int batch, int height,
                                    int width, int num_labels, void* out_nhwc, int pad, int c_pad, float* out_nchw,
                                    int* bad_label_count, void* stream_v) {
  if (mean == nullptr || std == nullptr) return fail(JPDSE_ERR_INVALID, "build_input_u8: NULL mean / std");
  ImgNorm nm;
  nm.u8 = 1;
  for (int c = 0; c < 3; ++c) {
    nm.mean[c] = mean[c];
    nm.std[c] = std[c];
    if (!(std[c] != 0.f)) return fail(JPDSE_ERR_INVALID, "build_input_u8: std must be non-zero");
  }
  return build_input_impl(label, label_dtype, instance, inst_dtype, image_u8, nm, batch, height, width, num_labels, out_nhwc,
                          pad, c_pad, out_nchw, bad_label_count, stream_v);
}

// plain launch, or (JPDSE_PDL=1) with the programmatic-stream-serialization attribute
template <typename K, typename... A>
static void launch_norm(K kernel, dim3 grid, cudaStream_t stream, A... args) {
  if (jpdse::pdl_enabled()) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(jpdse::kNormThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
  } else {
    kernel<<<grid, jpdse::kNormThreads, 0, stream>>>(args...);
  }
}

extern "C" int jpdse_instnorm_apply(const void* raw, const double* stats, const void* residual, void* out, int batch,
                                    int height, int width, int channels, int pad, int relu, float eps, void* stream_v) {
  if (raw == nullptr || stats == nullptr || out == nullptr) return fail(JPDSE_ERR_INVALID, "instnorm_apply: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0) return fail(JPDSE_ERR_INVALID, "instnorm_apply: bad sizes");
  if (channels % 8 || channels > 8 * kNormThreads || (kNormThreads % (channels / 8)))
    return fail(JPDSE_ERR_UNSUPPORTED, "instnorm_apply: channels must be 8*2^k <= %d (got %d)", 8 * kNormThreads, channels);
  if (pad < 0 || pad >= height || pad >= width) return fail(JPDSE_ERR_INVALID, "instnorm_apply: bad pad");
  if ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(residual)) & 15)
    return fail(JPDSE_ERR_INVALID, "instnorm_apply: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int vpp = channels / 8;
  const int ppi = kNormThreads / vpp;
  const int npix = (height + 2 * pad) * (width + 2 * pad);
  const int per_block = ppi * kNormIters;
  int gx = (npix + per_block - 1) / per_block;
  // one resident wave: CTAs per SM from the occupancy calculator, split over the images (grid.y)
  static int per_sm = 0;
  if (per_sm == 0) {
    int a = 0, b2 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, instnorm_apply_kernel<true, true>, kNormThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b2, instnorm_apply_kernel<true, false>, kNormThreads, 0);
    per_sm = a < b2 ? a : b2;
    if (per_sm < 1) per_sm = 1;
  }
  const int wave = (num_sms() * per_sm) / batch > 0 ? (num_sms() * per_sm) / batch : 1;
  // measured on B200 (batch 16): the single-wave grid wins for C >= 128 (0.275 -> 0.241 ms at C = 128, 0.063 -> 0.051
  // at C = 1024) and loses for the two full-resolution C = 64 layers (0.43 -> 0.455 ms), which keep one CTA per chunk
  int iters = kNormIters;
  if (channels >= 128 && gx > wave) {
    // One resident wave, and every CTA ONE chunk where a chunk of 20 ... 32 pixel groups makes that possible: with the
    // fixed 16 the 2244 padded pixels of a 1024-channel map are 71 chunks on 55 CTAs per image -- a quarter of the CTAs
    // run two chunks while the rest wait (JPDSE_NORM_BALANCE=0 keeps that form).
    static int balance = -1;
    if (balance < 0) {
      const char* e = getenv("JPDSE_NORM_BALANCE");
      balance = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    const int need = (npix + wave * ppi - 1) / (wave * ppi);  // pixel groups per CTA for one chunk each
    if (balance && need > 16 && need <= 32) {
      iters = (need + 3) / 4 * 4;
      gx = (npix + ppi * iters - 1) / (ppi * iters);
    } else {
      gx = wave;
    }
  }
  dim3 grid(gx, batch);
  const __nv_bfloat16* r = static_cast<const __nv_bfloat16*>(raw);
  const __nv_bfloat16* rs = static_cast<const __nv_bfloat16*>(residual);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
  if (iters != kNormIters) {
#define JPDSE_NORM_LAUNCH(IT)                                                                                                   \
  do {                                                                                                                          \
    if (relu && residual)                                                                                                       \
      launch_norm(instnorm_apply_kernel<true, true, IT>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps);  \
    else if (relu)                                                                                                              \
      launch_norm(instnorm_apply_kernel<true, false, IT>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps); \
    else if (residual)                                                                                                          \
      launch_norm(instnorm_apply_kernel<false, true, IT>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps); \
    else                                                                                                                        \
      launch_norm(instnorm_apply_kernel<false, false, IT>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps); \
  } while (0)
    switch (iters) {
      case 20: JPDSE_NORM_LAUNCH(20); break;
      case 24: JPDSE_NORM_LAUNCH(24); break;
      case 28: JPDSE_NORM_LAUNCH(28); break;
      default: JPDSE_NORM_LAUNCH(32); break;
    }
#undef JPDSE_NORM_LAUNCH
    return check_launch("instnorm_apply_kernel");
  }
  if (relu && residual)
    launch_norm(instnorm_apply_kernel<true, true>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps);
  else if (relu)
    launch_norm(instnorm_apply_kernel<true, false>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps);
  else if (residual)
    launch_norm(instnorm_apply_kernel<false, true>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps);
  else
    launch_norm(instnorm_apply_kernel<false, false>, grid, stream, r, stats, rs, o, height, width, channels, pad, eps);
  return check_launch("instnorm_apply_kernel");
}

extern "C" int jpdse_nchw_f32_to_nhwc_bf16(const float* x, void* y, int batch, int channels, int height, int width,
                                           int pad_reflect, int c_pad, void* stream) {
  if (x == nullptr || y == nullptr) return fail(JPDSE_ERR_INVALID, "nchw_to_nhwc: NULL pointer");
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || pad_reflect < 0 || pad_reflect >= height || pad_reflect >= width)
    return fail(JPDSE_ERR_INVALID, "nchw_to_nhwc: bad sizes");
  if (c_pad < channels) return fail(JPDSE_ERR_INVALID, "nchw_to_nhwc: c_pad < channels");
  const size_t npix = static_cast<size_t>(height + 2 * pad_reflect) * (width + 2 * pad_reflect);
  dim3 grid(static_cast<unsigned>((npix + 31) / 32), (c_pad + 31) / 32, batch), block(32, 32);
  nchw_f32_to_nhwc_bf16_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<__nv_bfloat16*>(y), batch, channels, height, width, pad_reflect, c_pad);
  return check_launch("nchw_f32_to_nhwc_bf16_kernel");
}

extern "C" int jpdse_nhwc_bf16_to_nchw_f32(const void* x, float* y, int batch, int channels, int height, int width,
                                           void* stream) {
  if (x == nullptr || y == nullptr) return fail(JPDSE_ERR_INVALID, "nhwc_to_nchw: NULL pointer");
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return fail(JPDSE_ERR_INVALID, "nhwc_to_nchw: bad sizes");
  const size_t npix = static_cast<size_t>(height) * width;
  dim3 grid(static_cast<unsigned>((npix + 31) / 32), (channels + 31) / 32, batch), block(32, 32);
  nhwc_bf16_to_nchw_f32_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, batch, channels, height, width);
  return check_launch("nhwc_bf16_to_nchw_f32_kernel");
}

template <int kOp>
static int launch_quant(const float* x, const float* u, float* y, size_t n, void* stream, const char* name) {
  if (n == 0) return JPDSE_OK;
  if (x == nullptr || y == nullptr || (kOp == kOpSoftSign && u == nullptr)) return fail(JPDSE_ERR_INVALID, "%s: NULL pointer", name);
  const int grid = grid_for(n, 256, 16);
  quant_elementwise_kernel<kOp><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, u, y, n);
  return check_launch(name);
}

extern "C" int jpdse_round_f32(const float* x, float* y, size_t n, void* stream) {
  return launch_quant<kOpRound>(x, nullptr, y, n, stream, "round_f32");
}
extern "C" int jpdse_sign_f32(const float* x, float* y, size_t n, void* stream) {
  return launch_quant<kOpSign>(x, nullptr, y, n, stream, "sign_f32");
}
extern "C" int jpdse_softsign_f32(const float* x, const float* u, float* y, size_t n, void* stream) {
  return launch_quant<kOpSoftSign>(x, u, y, n, stream, "softsign_f32");
}
extern "C" int jpdse_sign_to_bits_u8(const float* x, uint8_t* y, size_t n, void* stream) {
  if (n == 0) return JPDSE_OK;
  if (x == nullptr || y == nullptr) return fail(JPDSE_ERR_INVALID, "sign_to_bits: NULL pointer");
  sign_to_bits_kernel<<<grid_for(n, 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y, n);
  return check_launch("sign_to_bits_kernel");
}

extern "C" int jpdse_s2hvq_encode(const float* x, const float* code_book, size_t rows, int center_size, int n_center,
                                  float sigma, float* scores, int64_t* index, float* one_hot, float* soft, void* stream) {
  if (rows == 0) return JPDSE_OK;
  if (x == nullptr || code_book == nullptr) return fail(JPDSE_ERR_INVALID, "s2hvq_encode: NULL pointer");
  if (center_size <= 0 || n_center <= 0) return fail(JPDSE_ERR_INVALID, "s2hvq_encode: bad sizes");
  if (!(sigma > 0.f)) return fail(JPDSE_ERR_INVALID, "s2hvq_encode: sigma must be greater than 0");
  const size_t smem = static_cast<size_t>(center_size) * n_center * sizeof(float);
  if (smem > 200 * 1024) return fail(JPDSE_ERR_UNSUPPORTED, "s2hvq_encode: code book larger than shared memory");
  const bool vec_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  auto kern = s2hvq_encode_kernel<0>;
  if (center_size == 8 && vec_ok) kern = s2hvq_encode_kernel<8>;
  else if (center_size == 4 && vec_ok) kern = s2hvq_encode_kernel<4>;
  else if (center_size == 16 && vec_ok) kern = s2hvq_encode_kernel<16>;
  else if (center_size == 2) kern = s2hvq_encode_kernel<2>;
  else if (center_size == 1) kern = s2hvq_encode_kernel<1>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "s2hvq_encode: %s", cudaGetErrorString(e));
  }
  kern<<<grid_for(rows, 128, 1), 128, smem, static_cast<cudaStream_t>(stream)>>>(
      x, code_book, rows, center_size, n_center, sigma, scores, reinterpret_cast<long long*>(index), one_hot, soft);
  return check_launch("s2hvq_encode_kernel");
}

extern "C" int jpdse_s2hvq_decode(const float* code_raw, const float* code_book, size_t rows, int center_size,
                                  int n_center, float* out, int64_t* index, void* stream) {
  if (rows == 0) return JPDSE_OK;
  if (code_raw == nullptr || code_book == nullptr || out == nullptr) return fail(JPDSE_ERR_INVALID, "s2hvq_decode: NULL pointer");
  if (center_size <= 0 || n_center <= 0) return fail(JPDSE_ERR_INVALID, "s2hvq_decode: bad sizes");
  s2hvq_decode_kernel<<<grid_for(rows, 128, 1), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      code_raw, code_book, rows, center_size, n_center, out, reinterpret_cast<long long*>(index));
  return check_launch("s2hvq_decode_kernel");
}

// ------------------------------------------------------------------------------------------ eval-metric path
// tensor2im (ctu/utils/misc.py:64-95): uint8( clip( (x * std + mean) * 255, 0, 255 ) ) with numpy's float64 arithmetic and
// C truncation, and the distortion the reference then takes between two such images (pix2pixHD_model.py:636-641,
// test.py:115-123): L1Loss / MSELoss (mean over all elements) of the uint8 values.
namespace jpdse {

__device__ __forceinline__ uint8_t to_u8(float x, double mean, double std) {
  double v = (static_cast<double>(x) * std + mean) * 255.0;
  v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);  // np.clip; NaN falls through both compares like numpy's minimum/maximum do not,
  return static_cast<uint8_t>(v);                // but a NaN pixel is outside the reference's domain (tanh output / normalised image)
}

// 4 consecutive pixels per thread (3 channels): three 16-byte loads, 12 contiguous output bytes as three 4-byte stores
__global__ void __launch_bounds__(256)
tensor2im_u8x4_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int H, int W, double m0, double m1, double m2,
                      double s0, double s1, double s2) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t quads = plane / 4;
  const size_t total = static_cast<size_t>(B) * quads;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / quads, q = i % quads;
    const float4* src = reinterpret_cast<const float4*>(x + b * 3 * plane) + q;
    const float4 r = __ldg(src), g = __ldg(src + quads), bl = __ldg(src + 2 * quads);
    const uint32_t r0 = to_u8(r.x, m0, s0), r1 = to_u8(r.y, m0, s0), r2 = to_u8(r.z, m0, s0), r3 = to_u8(r.w, m0, s0);
    const uint32_t g0 = to_u8(g.x, m1, s1), g1 = to_u8(g.y, m1, s1), g2 = to_u8(g.z, m1, s1), g3 = to_u8(g.w, m1, s1);
    const uint32_t b0 = to_u8(bl.x, m2, s2), b1 = to_u8(bl.y, m2, s2), b2 = to_u8(bl.z, m2, s2), b3 = to_u8(bl.w, m2, s2);
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + (b * plane + q * 4) * 3);  // (B, H, W, 3): 12 bytes per quad
    dst[0] = r0 | (g0 << 8) | (b0 << 16) | (r1 << 24);
    dst[1] = g1 | (b1 << 8) | (r2 << 16) | (g2 << 24);
    dst[2] = b2 | (r3 << 8) | (g3 << 16) | (b3 << 24);
  }
}

__global__ void __launch_bounds__(256)
tensor2im_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int C, int H, int W, double m0, double m1,
                    double m2, double s0, double s1, double s2) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / plane, hw = i % plane;
    const float* src = x + b * C * plane + hw;
    uint8_t* dst = out + i * C;  // (B, H, W, C) like tensor2im's HWC images
    for (int c = 0; c < C; ++c) {
      const double mean = c == 0 ? m0 : (c == 1 ? m1 : m2), std = c == 0 ? s0 : (c == 1 ? s1 : s2);
      dst[c] = to_u8(__ldg(src + c * plane), mean, std);
    }
  }
}

// sum over all elements of |a - b| (mode 0) or (a - b)^2 (mode 1) of the de-normalised uint8 images, exact in int64
__global__ void __launch_bounds__(256)
distortion_u8_kernel(const float* __restrict__ a, const float* __restrict__ b, unsigned long long* __restrict__ acc, int B, int C,
                     int H, int W, int mode, double m0, double m1, double m2, double s0, double s1, double s2) {
  __shared__ unsigned long long s_part[8];
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * C * plane;
  unsigned long long sum = 0;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i / plane) % C);
    const double mean = c == 0 ? m0 : (c == 1 ? m1 : m2), std = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const int d = static_cast<int>(to_u8(__ldg(a + i), mean, std)) - static_cast<int>(to_u8(__ldg(b + i), mean, std));
    sum += static_cast<unsigned long long>(mode == 0 ? (d < 0 ? -d : d) : d * d);
  }
  for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    atomicAdd(acc, t);
  }
}

}  // namespace jpdse

extern "C" int jpdse_tensor2im_u8(const float* x, uint8_t* out, int batch, int channels, int height, int width,
                                  const double* mean, const double* std, void* stream) {
  if (x == nullptr || out == nullptr || mean == nullptr || std == nullptr) return fail(JPDSE_ERR_INVALID, "tensor2im_u8: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || channels < 1 || channels > 3)
    return fail(JPDSE_ERR_INVALID, "tensor2im_u8: bad sizes (channels must be 1..3)");
  const size_t total = static_cast<size_t>(batch) * height * width;
  const int grid = grid_for(total, 256, 4);
  const double m[3] = {mean[0], channels > 1 ? mean[1] : 0.0, channels > 2 ? mean[2] : 0.0};
  const double s[3] = {std[0], channels > 1 ? std[1] : 1.0, channels > 2 ? std[2] : 1.0};
  if (channels == 3 && (static_cast<size_t>(height) * width) % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(out) & 3)) == 0) {
    const int grid4 = grid_for(static_cast<size_t>(batch) * height * width / 4, 256, 2);
    tensor2im_u8x4_kernel<<<grid4, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, batch, height, width, m[0], m[1], m[2], s[0],
                                                                               s[1], s[2]);
    return check_launch("tensor2im_u8x4_kernel");
  }
  tensor2im_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, batch, channels, height, width, m[0], m[1], m[2],
                                                                            s[0], s[1], s[2]);
  return check_launch("tensor2im_u8_kernel");
}

extern "C" int jpdse_distortion_u8(const float* a, const float* b, unsigned long long* sum, int batch, int channels,
                                   int height, int width, int mode, const double* mean, const double* std, void* stream) {
  if (a == nullptr || b == nullptr || sum == nullptr || mean == nullptr || std == nullptr)
    return fail(JPDSE_ERR_INVALID, "distortion_u8: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || channels < 1 || channels > 3 || mode < 0 || mode > 1)
    return fail(JPDSE_ERR_INVALID, "distortion_u8: bad sizes / mode");
  const size_t total = static_cast<size_t>(batch) * channels * height * width;
  const int grid = grid_for(total, 256, 8);
  const double m[3] = {mean[0], channels > 1 ? mean[1] : 0.0, channels > 2 ? mean[2] : 0.0};
  const double s[3] = {std[0], channels > 1 ? std[1] : 1.0, channels > 2 ? std[2] : 1.0};
  distortion_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, sum, batch, channels, height, width, mode, m[0],
                                                                             m[1], m[2], s[0], s[1], s[2]);
  return check_launch("distortion_u8_kernel");
}

extern "C" int jpdse_d_input_ids(const void* label, int label_dtype, const void* instance, int inst_dtype, const float* image_a,
                                 void* out_a, const float* image_b, void* out_b, int batch, int height, int width, int num_labels,
                                 int c_pad, int pool, int out_pad, void* stream) {
  if (label == nullptr || instance == nullptr || image_a == nullptr || out_a == nullptr || ((image_b == nullptr) != (out_b == nullptr)))
    return fail(JPDSE_ERR_INVALID, "d_input_ids: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || num_labels <= 0 || num_labels + 4 > c_pad || c_pad % 8 || out_pad < 0)
    return fail(JPDSE_ERR_INVALID, "d_input_ids: bad sizes");
  if (label_dtype < 0 || label_dtype > 2 || inst_dtype < 0 || inst_dtype > 3) return fail(JPDSE_ERR_INVALID, "d_input_ids: bad dtype code");
  if ((reinterpret_cast<uintptr_t>(out_a) | reinterpret_cast<uintptr_t>(out_b)) & 15)
    return fail(JPDSE_ERR_INVALID, "d_input_ids: outputs must be 16-byte aligned");
  const int Ho = pool ? (height - 1) / 2 + 1 : height, Wo = pool ? (width - 1) / 2 + 1 : width;
  if (Ho > 65535 || batch > 65535) return fail(JPDSE_ERR_UNSUPPORTED, "d_input_ids: more than 65535 rows / images");
  if (c_pad > 8 * kDIdsMaxVec) return fail(JPDSE_ERR_UNSUPPORTED, "d_input_ids: c_pad > %d", 8 * kDIdsMaxVec);
  const dim3 blocks((Wo + kDIdsPixels - 1) / kDIdsPixels, Ho, batch);
  d_input_ids_kernel<<<blocks, kDIdsPixels, 0, static_cast<cudaStream_t>(stream)>>>(
      label, label_dtype, instance, inst_dtype, image_a, static_cast<__nv_bfloat16*>(out_a), image_b,
      static_cast<__nv_bfloat16*>(out_b), batch, height, width, Ho, Wo, num_labels, c_pad, pool ? 1 : 0, out_pad);
  return check_launch("d_input_ids_kernel");
}
