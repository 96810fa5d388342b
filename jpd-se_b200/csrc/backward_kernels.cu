// HBM-bound kernels of the generator BACKWARD pass (what `loss_G.backward()` runs through autograd in the
// reference: ctu/trainers/pix2pixHD_trainer.py:69 over ctu/models/pix2pixHD_networks/networks.py:198-305):
//
//   instnorm_backward_reduce  reflect-pad fold-back (+ skip-connection gradient) + ReLU mask -> dy (bf16), and the
//                             two per-(image, channel) sums  S1 = sum dy,  S2 = sum dy * xhat  of the
//                             InstanceNorm2d(affine=False) backward (networks.py:27-36)
//   instnorm_backward_apply   dx = rstd * (dy - S1/n - xhat * S2/n), written with the ZERO border the data-gradient
//                             conv (JPDSE_CONV3X3_FULL / CONV7X7_FULL) reads
//   tanh_backward_nchw        head: d_pre = dout * (1 - out^2) (nn.Tanh, networks.py:246) -> 8-channel NHWC bf16
//                             with a zero border of 6, plus the bias gradient
// Same conventions as bandwidth_kernels.cu: 16-byte vectors over the channel dimension, thread = (pixel, 8 channels).
#include <cuda_bf16.h>

#include <cstdint>

#include "common.cuh"

namespace jpdse {

constexpr int kBwdThreads = 256;
constexpr int kBwdItersMax = 32;  // pixel groups per CTA pass; the host lowers it for small tensors so the grid fills the SMs

__device__ __forceinline__ uint32_t bwd_pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

__device__ __forceinline__ void load_mean_rstd(const double* stats, int b, int C, int vec, double inv_n, float eps,
                                               float (&mean)[8], float (&rstd)[8]) {
  const double* st = stats + (static_cast<size_t>(b) * C + vec * 8) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const double m = st[2 * j] * inv_n;
    double var = st[2 * j + 1] * inv_n - m * m;
    if (var < 0.0) var = 0.0;
    mean[j] = static_cast<float>(m);
    rstd[j] = rsqrtf(static_cast<float>(var) + eps);  // same expression as the forward apply kernel
  }
}

// sources of the reflect-pad fold for interior index i: padded indices q with reflect(q - pad) == i
__device__ __forceinline__ int fold_sources(int i, int n, int pad, int (&q)[3]) {
  int c = 0;
  q[c++] = i + pad;
  if (pad > 0) {
    if (i >= 1 && i <= pad) q[c++] = pad - i;
    if (i <= n - 2 && i >= n - 1 - pad) q[c++] = pad + 2 * (n - 1) - i;
  }
  return c;
}

template <bool kRelu, bool kSkip>
__global__ void __launch_bounds__(kBwdThreads, 2)
instnorm_backward_reduce_kernel(const __nv_bfloat16* __restrict__ g, int gpad, const __nv_bfloat16* __restrict__ skip,
                                const __nv_bfloat16* __restrict__ raw, const double* __restrict__ stats,
                                __nv_bfloat16* __restrict__ dy, double* __restrict__ sums, int H, int W, int C, float eps, int iters,
                                float slope) {
  __shared__ float s_red[kBwdThreads][17];
  const int vpp = C >> 3;
  const int ppi = kBwdThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int npix = H * W;
  const int Wg = W + 2 * gpad, Hg = H + 2 * gpad;
  float mean[8], rstd[8];
  load_mean_rstd(stats, b, C, vec, 1.0 / (static_cast<double>(H) * W), eps, mean, rstd);
  const uint4* g4 = reinterpret_cast<const uint4*>(g) + static_cast<size_t>(b) * Hg * Wg * vpp;
  const uint4* skip4 = reinterpret_cast<const uint4*>(skip) + static_cast<size_t>(b) * npix * vpp;
  const uint4* raw4 = reinterpret_cast<const uint4*>(raw) + static_cast<size_t>(b) * npix * vpp;
  uint4* dy4 = reinterpret_cast<uint4*>(dy) + static_cast<size_t>(b) * npix * vpp;
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  // gradient of pixel pp's output: the pixel itself plus (border pixels only) the reflect-pad sources folded onto it
  auto load_g = [&](int h, int w, float (&acc)[8]) {
    unpack8(__ldg(g4 + (static_cast<size_t>(h + gpad) * Wg + (w + gpad)) * vpp + vec), acc);
  };
  auto fold_g = [&](int h, int w, float (&acc)[8]) {
    int qh[3], qw[3];
    const int nh = fold_sources(h, H, gpad, qh), nw = fold_sources(w, W, gpad, qw);
    if (nh * nw > 1) {
      for (int a = 0; a < nh; ++a)
        for (int c = (a == 0 ? 1 : 0); c < nw; ++c) {
          float f[8];
          unpack8(__ldg(g4 + (static_cast<size_t>(qh[a]) * Wg + qw[c]) * vpp + vec), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
    }
  };
  auto finish = [&](int pp, const uint4& xr, const uint4& sk, float (&acc)[8]) {
    if (kSkip) {
      float f[8];
      unpack8(sk, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += f[j];
    }
    float x[8];
    unpack8(xr, x);
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float xh0 = (x[j] - mean[j]) * rstd[j], xh1 = (x[j + 1] - mean[j + 1]) * rstd[j + 1];
      float d0 = acc[j], d1 = acc[j + 1];
      if (kRelu) {  // ReLU (slope 0: exact zeros) or LeakyReLU(slope) mask
        d0 = xh0 > 0.f ? d0 : (slope == 0.f ? 0.f : slope * d0);
        d1 = xh1 > 0.f ? d1 : (slope == 0.f ? 0.f : slope * d1);
      }
      const uint32_t pk = bwd_pack_bf16x2(d0, d1);
      ow[j >> 1] = pk;
      d0 = __uint_as_float(pk << 16);  // sums over the stored (bf16-rounded) gradient: the apply pass re-reads it
      d1 = __uint_as_float(pk & 0xffff0000u);
      s1[j] += d0;
      s1[j + 1] += d1;
      s2[j] = fmaf(d0, xh0, s2[j]);
      s2[j + 1] = fmaf(d1, xh1, s2[j + 1]);
    }
    dy4[static_cast<size_t>(pp) * vpp + vec] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  };
  // The loads of TWO pixels (raw, gradient, skip: up to six 16-byte loads) are issued before anything is consumed; with a
  // bounds check between the pixels the compiler cannot hoist the second pixel's loads and a thread has one pixel in
  // flight (the kernel sat at ~1 TB/s at training batch sizes).
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters)) {
    if (pix0 + ppi * iters <= npix && (iters & 1) == 0) {
      for (int it = 0; it < iters; it += 2) {
        const int pa = pix0 + it * ppi + psub, pb = pa + ppi;
        const int ha = pa / W, wa = pa - ha * W, hb = pb / W, wb = pb - hb * W;
        const uint4 xa = __ldg(raw4 + static_cast<size_t>(pa) * vpp + vec), xb = __ldg(raw4 + static_cast<size_t>(pb) * vpp + vec);
        uint4 ska = make_uint4(0, 0, 0, 0), skb = make_uint4(0, 0, 0, 0);
        if (kSkip) {
          ska = __ldg(skip4 + static_cast<size_t>(pa) * vpp + vec);
          skb = __ldg(skip4 + static_cast<size_t>(pb) * vpp + vec);
        }
        float ga[8], gb[8];
        load_g(ha, wa, ga);
        load_g(hb, wb, gb);
        if (gpad > 0) {
          fold_g(ha, wa, ga);
          fold_g(hb, wb, gb);
        }
        finish(pa, xa, ska, ga);
        finish(pb, xb, skb, gb);
      }
    } else {
      for (int it = 0; it < iters; ++it) {
        const int pp = pix0 + it * ppi + psub;
        if (pp >= npix) break;
        const int h = pp / W, w = pp - h * W;
        const uint4 xr = __ldg(raw4 + static_cast<size_t>(pp) * vpp + vec);
        uint4 sk = make_uint4(0, 0, 0, 0);
        if (kSkip) sk = __ldg(skip4 + static_cast<size_t>(pp) * vpp + vec);
        float acc[8];
        load_g(h, w, acc);
        if (gpad > 0) fold_g(h, w, acc);
        finish(pp, xr, sk, acc);
      }
    }
  }
  // block reduction over the pixel sub-lanes of each channel vector
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    s_red[threadIdx.x][j] = s1[j];
    s_red[threadIdx.x][8 + j] = s2[j];
  }
  __syncthreads();
  for (int stride = ppi >> 1; stride >= 1; stride >>= 1) {
    if (psub < stride) {
#pragma unroll
      for (int j = 0; j < 16; ++j) s_red[threadIdx.x][j] += s_red[threadIdx.x + stride * vpp][j];
    }
    __syncthreads();
  }
  if (psub == 0) {
    double* out = sums + (static_cast<size_t>(b) * C + vec * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(out + 2 * j, static_cast<double>(s_red[threadIdx.x][j]));
      atomicAdd(out + 2 * j + 1, static_cast<double>(s_red[threadIdx.x][8 + j]));
    }
  }
}

__global__ void __launch_bounds__(kBwdThreads, 2)
instnorm_backward_apply_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ raw,
                               const double* __restrict__ stats, const double* __restrict__ sums,
                               __nv_bfloat16* __restrict__ dx, int zpad, int H, int W, int C, float eps, int iters) {
  const int vpp = C >> 3;
  const int ppi = kBwdThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  // JPDSE_PAD_SHARED: pitch W + pad, image stride (H + pad) rows; the last image also owns the trailing zero run
  const bool shared = (zpad & JPDSE_PAD_SHARED) != 0;
  zpad &= JPDSE_PAD_SHARED - 1;
  const int Wz = shared ? W + zpad : W + 2 * zpad, Hz = shared ? H + zpad : H + 2 * zpad;
  const int npix = Hz * Wz + ((shared && b == static_cast<int>(gridDim.y) - 1) ? zpad * Wz + zpad : 0);
  const double inv_n = 1.0 / (static_cast<double>(H) * W);
  float mean[8], rstd[8], m1[8], m2[8];
  load_mean_rstd(stats, b, C, vec, inv_n, eps, mean, rstd);
  {
    const double* sm = sums + (static_cast<size_t>(b) * C + vec * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m1[j] = static_cast<float>(sm[2 * j] * inv_n);
      m2[j] = static_cast<float>(sm[2 * j + 1] * inv_n);
    }
  }
  const uint4* dy4 = reinterpret_cast<const uint4*>(dy) + static_cast<size_t>(b) * H * W * vpp;
  const uint4* raw4 = reinterpret_cast<const uint4*>(raw) + static_cast<size_t>(b) * H * W * vpp;
  uint4* dx4 = reinterpret_cast<uint4*>(dx) + static_cast<size_t>(b) * Hz * Wz * vpp;
  auto finish = [&](int pp, bool inside, const uint4& dv, const uint4& xv) {
    uint4 o = make_uint4(0, 0, 0, 0);
    if (inside) {
      float d[8], x[8];
      unpack8(dv, d);
      unpack8(xv, x);
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float xh0 = (x[j] - mean[j]) * rstd[j], xh1 = (x[j + 1] - mean[j + 1]) * rstd[j + 1];
        ow[j >> 1] = bwd_pack_bf16x2(rstd[j] * (d[j] - m1[j] - xh0 * m2[j]), rstd[j + 1] * (d[j + 1] - m1[j + 1] - xh1 * m2[j + 1]));
      }
      o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    dx4[static_cast<size_t>(pp) * vpp + vec] = o;
  };
  // four pixels' loads in flight before anything is consumed (see instnorm_backward_reduce_kernel)
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters)) {
    if (pix0 + ppi * iters <= npix && (iters & 3) == 0) {
      for (int it0 = 0; it0 < iters; it0 += 4) {
        uint4 dv[4], xv[4];
        int pp[4];
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          pp[u] = pix0 + (it0 + u) * ppi + psub;
          const int ph = pp[u] / Wz, pw = pp[u] - ph * Wz;
          const int h = ph - zpad, w = pw - zpad;
          in[u] = h >= 0 && h < H && w >= 0 && w < W;
          dv[u] = xv[u] = make_uint4(0, 0, 0, 0);
          if (in[u]) {
            const size_t src = (static_cast<size_t>(h) * W + w) * vpp + vec;
            dv[u] = __ldg(dy4 + src);
            xv[u] = __ldg(raw4 + src);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) finish(pp[u], in[u], dv[u], xv[u]);
      }
    } else {
      for (int it = 0; it < iters; ++it) {
        const int pp = pix0 + it * ppi + psub;
        if (pp >= npix) break;
        const int ph = pp / Wz, pw = pp - ph * Wz;
        const int h = ph - zpad, w = pw - zpad;
        const bool inside = h >= 0 && h < H && w >= 0 && w < W;
        uint4 dv = make_uint4(0, 0, 0, 0), xv = make_uint4(0, 0, 0, 0);
        if (inside) {
          const size_t src = (static_cast<size_t>(h) * W + w) * vpp + vec;
          dv = __ldg(dy4 + src);
          xv = __ldg(raw4 + src);
        }
        finish(pp, inside, dv, xv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ fused reduce + apply
// Small feature maps (H*W <= 8 * kBwdThreads: the 1024-channel bottleneck at 1024x512 images is 32 x 64 = 2048 pixels):
// ONE CTA owns all pixels of (image, 8 channels), so the two sums of the InstanceNorm backward are a block reduction and
// the masked gradient dy never goes to memory between the two halves -- it stays in the thread's registers. One launch
// instead of reduce + apply, no atomics, no sums buffer: reads g (+ skip) and raw once, writes dx once (and dy only when
// the caller needs it: the ResnetBlock skip connection).
constexpr int kFusedMaxPix = 8;  // pixels per thread

template <bool kRelu, bool kSkip>
__global__ void __launch_bounds__(kBwdThreads, 2)
instnorm_backward_fused_kernel(const __nv_bfloat16* __restrict__ g, int gpad, const __nv_bfloat16* __restrict__ skip,
                               const __nv_bfloat16* __restrict__ raw, const double* __restrict__ stats,
                               __nv_bfloat16* __restrict__ dy_out, __nv_bfloat16* __restrict__ dx, int zpad, int H, int W, int C,
                               float eps, float slope) {
  __shared__ float s_red[kBwdThreads / 32][16];
  __shared__ float s_tot[16];
  const int vpp = C >> 3;
  const int vec = blockIdx.x;
  const int b = blockIdx.y;
  const int npix = H * W;
  const int Wg = W + 2 * gpad, Hg = H + 2 * gpad;
  const bool shared = (zpad & JPDSE_PAD_SHARED) != 0;
  zpad &= JPDSE_PAD_SHARED - 1;
  const int Wz = shared ? W + zpad : W + 2 * zpad, Hz = shared ? H + zpad : H + 2 * zpad;
  const double inv_n = 1.0 / (static_cast<double>(H) * W);
  float mean[8], rstd[8];
  load_mean_rstd(stats, b, C, vec, inv_n, eps, mean, rstd);
  const uint4* g4 = reinterpret_cast<const uint4*>(g) + static_cast<size_t>(b) * Hg * Wg * vpp;
  const uint4* skip4 = reinterpret_cast<const uint4*>(skip) + static_cast<size_t>(b) * npix * vpp;
  const uint4* raw4 = reinterpret_cast<const uint4*>(raw) + static_cast<size_t>(b) * npix * vpp;
  uint4* dy4 = reinterpret_cast<uint4*>(dy_out) + static_cast<size_t>(b) * npix * vpp;
  uint4* dx4 = reinterpret_cast<uint4*>(dx) + static_cast<size_t>(b) * Hz * Wz * vpp;
  // ---- phase 1: every load of this thread's pixels is issued before anything is consumed
  uint4 xr[kFusedMaxPix], gv[kFusedMaxPix], sk[kFusedMaxPix];
#pragma unroll
  for (int u = 0; u < kFusedMaxPix; ++u) {
    const int pp = threadIdx.x + u * kBwdThreads;
    xr[u] = gv[u] = sk[u] = make_uint4(0, 0, 0, 0);
    if (pp < npix) {
      const int h = pp / W, w = pp - h * W;
      xr[u] = __ldg(raw4 + static_cast<size_t>(pp) * vpp + vec);
      gv[u] = __ldg(g4 + (static_cast<size_t>(h + gpad) * Wg + (w + gpad)) * vpp + vec);
      if (kSkip) sk[u] = __ldg(skip4 + static_cast<size_t>(pp) * vpp + vec);
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  uint4 dyv[kFusedMaxPix];
#pragma unroll
  for (int u = 0; u < kFusedMaxPix; ++u) {
    const int pp = threadIdx.x + u * kBwdThreads;
    dyv[u] = make_uint4(0, 0, 0, 0);
    if (pp < npix) {
      float acc[8];
      unpack8(gv[u], acc);
      if (gpad > 0) {  // reflect-pad fold-back: border pixels only
        const int h = pp / W, w = pp - h * W;
        int qh[3], qw[3];
        const int nh = fold_sources(h, H, gpad, qh), nw = fold_sources(w, W, gpad, qw);
        if (nh * nw > 1) {
          for (int a = 0; a < nh; ++a)
            for (int c = (a == 0 ? 1 : 0); c < nw; ++c) {
              float f[8];
              unpack8(__ldg(g4 + (static_cast<size_t>(qh[a]) * Wg + qw[c]) * vpp + vec), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] += f[j];
            }
        }
      }
      if (kSkip) {
        float f[8];
        unpack8(sk[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
      float x[8];
      unpack8(xr[u], x);
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float xh0 = (x[j] - mean[j]) * rstd[j], xh1 = (x[j + 1] - mean[j + 1]) * rstd[j + 1];
        float d0 = acc[j], d1 = acc[j + 1];
        if (kRelu) {
          d0 = xh0 > 0.f ? d0 : (slope == 0.f ? 0.f : slope * d0);
          d1 = xh1 > 0.f ? d1 : (slope == 0.f ? 0.f : slope * d1);
        }
        const uint32_t pk = bwd_pack_bf16x2(d0, d1);
        ow[j >> 1] = pk;
        d0 = __uint_as_float(pk << 16);  // sums over the bf16-rounded gradient, like the two-kernel path
        d1 = __uint_as_float(pk & 0xffff0000u);
        s1[j] += d0;
        s1[j + 1] += d1;
        s2[j] = fmaf(d0, xh0, s2[j]);
        s2[j + 1] = fmaf(d1, xh1, s2[j + 1]);
      }
      dyv[u] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      if (dy_out != nullptr) dy4[static_cast<size_t>(pp) * vpp + vec] = dyv[u];
    }
  }
  // ---- block reduction of the 16 sums (fixed order: bit-reproducible from run to run)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = s1[j], c = s2[j];
    for (int o = 16; o >= 1; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
      s_red[warp][j] = a;
      s_red[warp][8 + j] = c;
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int wv = 0; wv < kBwdThreads / 32; ++wv) t += s_red[wv][threadIdx.x];
    s_tot[threadIdx.x] = t * static_cast<float>(inv_n);
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m1[j] = s_tot[j];
    m2[j] = s_tot[8 + j];
  }
  // ---- phase 2: dx = rstd * (dy - mean(dy) - xhat * mean(dy * xhat)), written with its zero border
#pragma unroll
  for (int u = 0; u < kFusedMaxPix; ++u) {
    const int pp = threadIdx.x + u * kBwdThreads;
    if (pp < npix) {
      const int h = pp / W, w = pp - h * W;
      float d[8], x[8];
      unpack8(dyv[u], d);
      unpack8(xr[u], x);
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float xh0 = (x[j] - mean[j]) * rstd[j], xh1 = (x[j + 1] - mean[j + 1]) * rstd[j + 1];
        ow[j >> 1] = bwd_pack_bf16x2(rstd[j] * (d[j] - m1[j] - xh0 * m2[j]), rstd[j + 1] * (d[j + 1] - m1[j + 1] - xh1 * m2[j + 1]));
      }
      dx4[(static_cast<size_t>(h + zpad) * Wz + (w + zpad)) * vpp + vec] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
  if (shared) {
    // every stored position of this image that is not an interior pixel (the last image: plus the trailing zero run)
    const int stored = Hz * Wz + (b == static_cast<int>(gridDim.y) - 1 ? zpad * Wz + zpad : 0);
    for (int i = threadIdx.x; i < stored; i += kBwdThreads) {
      const int ph = i / Wz, pw = i - ph * Wz;
      if (ph < zpad || ph >= H + zpad || pw < zpad) dx4[static_cast<size_t>(i) * vpp + vec] = make_uint4(0, 0, 0, 0);
    }
  } else if (zpad > 0) {
    const int nborder = Hz * Wz - npix;
    for (int i = threadIdx.x; i < nborder; i += kBwdThreads) {
      // border pixels in stored order: top rows, bottom rows, then the left / right columns of the interior rows
      int ph, pw;
      const int top = zpad * Wz;
      if (i < top) {
        ph = i / Wz;
        pw = i - ph * Wz;
      } else if (i < 2 * top) {
        const int k = i - top;
        ph = H + zpad + k / Wz;
        pw = k % Wz;
      } else {
        const int k = i - 2 * top;
        ph = zpad + k / (2 * zpad);
        const int c = k % (2 * zpad);
        pw = c < zpad ? c : W + c;
      }
      dx4[(static_cast<size_t>(ph) * Wz + pw) * vpp + vec] = make_uint4(0, 0, 0, 0);
    }
  }
}

// one thread per STORED pixel of d_pre (B, H+12, W+12, 8)
__global__ void __launch_bounds__(256)
tanh_backward_nchw_kernel(const float* __restrict__ gout, const float* __restrict__ out, __nv_bfloat16* __restrict__ dpre,
                          float* __restrict__ dbias, int B, int Cout, int H, int W) {
  __shared__ float s_b[8][8];
  const int Ws = W + 12, Hs = H + 12;
  const size_t total = static_cast<size_t>(B) * Hs * Ws;
  const size_t plane = static_cast<size_t>(H) * W;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int pw = static_cast<int>(i % Ws);
    const size_t r = i / Ws;
    const int ph = static_cast<int>(r % Hs);
    const int b = static_cast<int>(r / Hs);
    const int h = ph - 6, w = pw - 6;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const size_t src = static_cast<size_t>(b) * Cout * plane + static_cast<size_t>(h) * W + w;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < Cout) {
          const float y = __ldg(out + src + j * plane);
          v[j] = __ldg(gout + src + j * plane) * (1.f - y * y);
          bsum[j] += v[j];
        }
    }
    reinterpret_cast<uint4*>(dpre)[i] =
        make_uint4(bwd_pack_bf16x2(v[0], v[1]), bwd_pack_bf16x2(v[2], v[3]), bwd_pack_bf16x2(v[4], v[5]), bwd_pack_bf16x2(v[6], v[7]));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float s = bsum[j];
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) s_b[warp][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < Cout) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += s_b[wv][threadIdx.x];
    atomicAdd(dbias + threadIdx.x, s);
  }
}

// pixel groups per CTA: as many as kBwdItersMax, fewer while the grid would leave SMs idle
static int pick_iters(int npix, int ppi, int batch) {
  int iters = kBwdItersMax;
  while (iters > 2 && static_cast<long long>((npix + ppi * iters - 1) / (ppi * iters)) * batch < 1LL * num_sms()) iters >>= 1;
  return iters;
}

static int check_channels(int channels, const char* what) {
  if (channels % 8 || channels > 8 * kBwdThreads || (kBwdThreads % (channels / 8)))
    return fail(JPDSE_ERR_UNSUPPORTED, "%s: channels must be 8*2^k <= %d (got %d)", what, 8 * kBwdThreads, channels);
  return JPDSE_OK;
}

}  // namespace jpdse

using namespace jpdse;

extern "C" int jpdse_instnorm_backward_reduce_act(const void* g, int g_pad, const void* skip, const void* raw, const double* stats,
                                                  void* dy, double* sums, int batch, int height, int width, int channels, int relu,
                                                  float slope, float eps, void* stream_v);

extern "C" int jpdse_instnorm_backward_reduce(const void* g, int g_pad, const void* skip, const void* raw, const double* stats,
                                              void* dy, double* sums, int batch, int height, int width, int channels, int relu,
                                              float eps, void* stream_v) {
  return jpdse_instnorm_backward_reduce_act(g, g_pad, skip, raw, stats, dy, sums, batch, height, width, channels, relu, 0.f, eps,
                                            stream_v);
}

extern "C" int jpdse_instnorm_backward_reduce_act(const void* g, int g_pad, const void* skip, const void* raw, const double* stats,
                                                  void* dy, double* sums, int batch, int height, int width, int channels, int relu,
                                                  float slope, float eps, void* stream_v) {
  if (g == nullptr || raw == nullptr || stats == nullptr || dy == nullptr || sums == nullptr)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_reduce: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0) return fail(JPDSE_ERR_INVALID, "instnorm_backward_reduce: bad sizes");
  int rc = check_channels(channels, "instnorm_backward_reduce");
  if (rc != JPDSE_OK) return rc;
  if (g_pad < 0 || 2 * g_pad >= height || 2 * g_pad >= width) return fail(JPDSE_ERR_INVALID, "instnorm_backward_reduce: bad pad");
  if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(raw) |
       reinterpret_cast<uintptr_t>(dy)) & 15)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_reduce: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int vpp = channels / 8, ppi = kBwdThreads / vpp;
  const int iters = pick_iters(height * width, ppi, batch);
  const int per_block = ppi * iters;
  int gx = (height * width + per_block - 1) / per_block;
  static int per_sm_r = 0;
  if (per_sm_r == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_r, instnorm_backward_reduce_kernel<true, true>, kBwdThreads, 0);
    if (per_sm_r < 1) per_sm_r = 1;
  }
  const int wave = (num_sms() * per_sm_r) / batch > 0 ? (num_sms() * per_sm_r) / batch : 1;
  if (batch >= 8 && channels >= 128 && gx > wave) gx = wave;  // the single-wave grid only pays at inference-size batches
  dim3 grid(gx, batch);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(g);
  const __nv_bfloat16* sp = static_cast<const __nv_bfloat16*>(skip);
  const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(raw);
  __nv_bfloat16* dp = static_cast<__nv_bfloat16*>(dy);
  if (relu && skip)
    instnorm_backward_reduce_kernel<true, true><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, dp, sums, height, width, channels, eps, iters, slope);
  else if (relu)
    instnorm_backward_reduce_kernel<true, false><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, dp, sums, height, width, channels, eps, iters, slope);
  else if (skip)
    instnorm_backward_reduce_kernel<false, true><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, dp, sums, height, width, channels, eps, iters, slope);
  else
    instnorm_backward_reduce_kernel<false, false><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, dp, sums, height, width, channels, eps, iters, slope);
  return check_launch("instnorm_backward_reduce_kernel");
}

extern "C" int jpdse_instnorm_backward_fused(const void* g, int g_pad, const void* skip, const void* raw, const double* stats,
                                             void* dy, void* dx, int dx_pad, int batch, int height, int width, int channels,
                                             int relu, float slope, float eps, void* stream_v) {
  if (g == nullptr || raw == nullptr || stats == nullptr || dx == nullptr)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_fused: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || dx_pad < 0 || channels % 8)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_fused: bad sizes");
  if (height * width > kFusedMaxPix * kBwdThreads)
    return fail(JPDSE_ERR_UNSUPPORTED, "instnorm_backward_fused: %d x %d pixels exceed the %d one CTA holds (use the reduce + "
                "apply pair)", width, height, kFusedMaxPix * kBwdThreads);
  if (g_pad < 0 || 2 * g_pad >= height || 2 * g_pad >= width) return fail(JPDSE_ERR_INVALID, "instnorm_backward_fused: bad pad");
  if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(raw) |
       reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_fused: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  dim3 grid(channels / 8, batch);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(g);
  const __nv_bfloat16* sp = static_cast<const __nv_bfloat16*>(skip);
  const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(raw);
  __nv_bfloat16* yp = static_cast<__nv_bfloat16*>(dy);
  __nv_bfloat16* xp = static_cast<__nv_bfloat16*>(dx);
  if (relu && skip)
    instnorm_backward_fused_kernel<true, true><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, yp, xp, dx_pad, height, width, channels, eps, slope);
  else if (relu)
    instnorm_backward_fused_kernel<true, false><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, yp, xp, dx_pad, height, width, channels, eps, slope);
  else if (skip)
    instnorm_backward_fused_kernel<false, true><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, yp, xp, dx_pad, height, width, channels, eps, slope);
  else
    instnorm_backward_fused_kernel<false, false><<<grid, kBwdThreads, 0, stream>>>(gp, g_pad, sp, rp, stats, yp, xp, dx_pad, height, width, channels, eps, slope);
  return check_launch("instnorm_backward_fused_kernel");
}

extern "C" int jpdse_instnorm_backward_apply(const void* dy, const void* raw, const double* stats, const double* sums, void* dx,
                                             int dx_pad, int batch, int height, int width, int channels, float eps,
                                             void* stream_v) {
  if (dy == nullptr || raw == nullptr || stats == nullptr || sums == nullptr || dx == nullptr)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_apply: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || dx_pad < 0) return fail(JPDSE_ERR_INVALID, "instnorm_backward_apply: bad sizes");
  int rc = check_channels(channels, "instnorm_backward_apply");
  if (rc != JPDSE_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(dx)) & 15)
    return fail(JPDSE_ERR_INVALID, "instnorm_backward_apply: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int vpp = channels / 8, ppi = kBwdThreads / vpp;
  const int zp = dx_pad & (JPDSE_PAD_SHARED - 1);
  const int npix = (dx_pad & JPDSE_PAD_SHARED) ? (height + zp) * (width + zp) + zp * (width + zp) + zp
                                               : (height + 2 * zp) * (width + 2 * zp);
  const int iters = pick_iters(npix, ppi, batch);
  const int per_block = ppi * iters;
  int gx = (npix + per_block - 1) / per_block;
  static int per_sm_a = 0;
  if (per_sm_a == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_a, instnorm_backward_apply_kernel, kBwdThreads, 0);
    if (per_sm_a < 1) per_sm_a = 1;
  }
  const int wave = (num_sms() * per_sm_a) / batch > 0 ? (num_sms() * per_sm_a) / batch : 1;
  if (batch >= 8 && channels >= 128 && gx > wave) gx = wave;
  dim3 grid(gx, batch);
  instnorm_backward_apply_kernel<<<grid, kBwdThreads, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(raw), stats, sums, static_cast<__nv_bfloat16*>(dx),
      dx_pad, height, width, channels, eps, iters);
  return check_launch("instnorm_backward_apply_kernel");
}

extern "C" int jpdse_tanh_backward_nchw(const float* grad_out, const float* out, void* d_pre, float* dbias, int batch,
                                        int channels, int height, int width, void* stream_v) {
  if (grad_out == nullptr || out == nullptr || d_pre == nullptr || dbias == nullptr)
    return fail(JPDSE_ERR_INVALID, "tanh_backward_nchw: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || channels <= 0 || channels > 8)
    return fail(JPDSE_ERR_INVALID, "tanh_backward_nchw: bad sizes (channels must be 1..8)");
  if (reinterpret_cast<uintptr_t>(d_pre) & 15) return fail(JPDSE_ERR_INVALID, "tanh_backward_nchw: d_pre must be 16-byte aligned");
  const size_t total = static_cast<size_t>(batch) * (height + 12) * (width + 12);
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  tanh_backward_nchw_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      grad_out, out, static_cast<__nv_bfloat16*>(d_pre), dbias, batch, channels, height, width);
  return check_launch("tanh_backward_nchw_kernel");
}

// ------------------------------------------------------------------------------------------ Binarizer, training mode
// ctu/quantizers/binarize.py:44-65 in train(): y = SoftSign(tanh(conv1x1(x))). The 1x1 conv is the implicit-GEMM kernel
// (raw bf16 NHWC output); these two kernels are the rest: tanh + stochastic sign forward (noise supplied by the caller,
// drawn like the reference), and the straight-through backward  d_pre = dy * (1 - tanh^2)  (binarize.py:26-28).
namespace jpdse {

__global__ void __launch_bounds__(256)
binarizer_train_fwd_kernel(const __nv_bfloat16* __restrict__ pre, const float* __restrict__ u, float* __restrict__ y,
                           float* __restrict__ t_out, int B, int C, int H, int W) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * C * plane;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // i indexes NCHW (the module's output layout); the conv output is NHWC
    const size_t hw = i % plane;
    const size_t bc = i / plane;
    const int c = static_cast<int>(bc % C);
    const size_t b = bc / C;
    const float t = tanhf(__bfloat162float(pre[(b * plane + hw) * C + c]));
    t_out[i] = t;
    const float thr = (1.f - t) / 2.f;  // SoftSignFunction.forward
    y[i] = thr <= __ldg(u + i) ? 1.f : -1.f;
  }
}

__global__ void __launch_bounds__(256)
binarizer_train_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ t, __nv_bfloat16* __restrict__ dpre, int B,
                           int C, int H, int W) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * C * plane;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // i indexes NHWC (coalesced writes of the conv-gradient operand)
    const int c = static_cast<int>(i % C);
    const size_t p = i / C;
    const size_t hw = p % plane, b = p / plane;
    const size_t src = (b * C + c) * plane + hw;
    const float tv = __ldg(t + src);
    dpre[i] = __float2bfloat16_rn(__ldg(gy + src) * (1.f - tv * tv));
  }
}

}  // namespace jpdse

extern "C" int jpdse_binarizer_train_forward(const void* pre_nhwc, const float* noise, float* y, float* tanh_out, int batch,
                                             int channels, int height, int width, void* stream) {
  if (pre_nhwc == nullptr || noise == nullptr || y == nullptr || tanh_out == nullptr)
    return fail(JPDSE_ERR_INVALID, "binarizer_train_forward: NULL pointer");
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return fail(JPDSE_ERR_INVALID, "binarizer_train_forward: bad sizes");
  const size_t total = static_cast<size_t>(batch) * channels * height * width;
  size_t blocks = (total + 1023) / 1024;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  binarizer_train_fwd_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(pre_nhwc), noise, y, tanh_out, batch, channels, height, width);
  return check_launch("binarizer_train_fwd_kernel");
}

extern "C" int jpdse_binarizer_train_backward(const float* grad_y, const float* tanh_out, void* d_pre_nhwc, int batch,
                                              int channels, int height, int width, void* stream) {
  if (grad_y == nullptr || tanh_out == nullptr || d_pre_nhwc == nullptr)
    return fail(JPDSE_ERR_INVALID, "binarizer_train_backward: NULL pointer");
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0) return fail(JPDSE_ERR_INVALID, "binarizer_train_backward: bad sizes");
  const size_t total = static_cast<size_t>(batch) * channels * height * width;
  size_t blocks = (total + 1023) / 1024;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  binarizer_train_bwd_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      grad_y, tanh_out, static_cast<__nv_bfloat16*>(d_pre_nhwc), batch, channels, height, width);
  return check_launch("binarizer_train_bwd_kernel");
}
