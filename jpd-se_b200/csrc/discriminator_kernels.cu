// HBM-bound kernels around the PatchGAN discriminator and the feature losses of the training step
// (reference: ctu/models/pix2pixHD_networks/networks.py:371-471 MultiscaleDiscriminator / NLayerDiscriminator,
// :80-139 GANLoss / VGGLoss, ctu/models/pix2pixHD_model.py:451-460 discriminate, :746-756 feature matching):
//
//   d_input            cat(input_label, image) [+ AvgPool2d(3, 2, 1, count_include_pad=False), networks.py:387] from the
//                      reference's float32 NCHW tensors -> NHWC bf16 with the ZERO border of 2 the 4x4 convs read
//   d_input_backward   gradient of that (both scales) w.r.t. a channel range of the float32 NCHW input
//   instnorm_apply_act InstanceNorm2d apply + LeakyReLU(0.2) (networks.py:436-445) -> zero-bordered NHWC bf16
//   act_backward       LeakyReLU / ReLU backward through a layer WITHOUT a norm (PatchGAN layer 0, every VGG conv),
//                      sum of up to two incoming gradients, zero-bordered output, optional bias gradient
//   l1_pair / l1_pair_backward   nn.L1Loss between two feature maps of one shape (feature matching, VGG loss)
//   maxpool2x2 / maxpool2x2_backward   nn.MaxPool2d(2, 2) of torchvision's VGG19 on zero-bordered NHWC bf16
//   nhwc_pad_to_nchw_f32   feature maps back to the reference's float32 NCHW (the generic netD.forward API)
// Conventions as in bandwidth_kernels.cu: 16-byte vectors over the channel dimension, thread = (pixel, 8 channels).
#include <cuda_bf16.h>

#include <cstdint>

#include "common.cuh"

namespace jpdse {

constexpr int kDThreads = 256;

__device__ __forceinline__ uint32_t d_pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void d_unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// grid for a (stored pixels x channel vectors) sweep: `iters` pixel groups per CTA, a few waves at most
static dim3 sweep_grid(long long npix, int ppi, int batch, int* iters_out) {
  int iters = 16;
  while (iters > 1 && ((npix + static_cast<long long>(ppi) * iters - 1) / (static_cast<long long>(ppi) * iters)) * batch <
                          2LL * num_sms())
    iters >>= 1;
  *iters_out = iters;
  long long gx = (npix + static_cast<long long>(ppi) * iters - 1) / (static_cast<long long>(ppi) * iters);
  const long long cap = 8LL * num_sms();
  if (gx * batch > cap) gx = cap / batch > 0 ? cap / batch : 1;
  return dim3(static_cast<unsigned>(gx), static_cast<unsigned>(batch));
}

static int check_vec_channels(int channels, const char* what) {
  if (channels % 8 || channels > 8 * kDThreads || (kDThreads % (channels / 8)))
    return fail(JPDSE_ERR_UNSUPPORTED, "%s: channels must be 8*2^k <= %d (got %d)", what, 8 * kDThreads, channels);
  return JPDSE_OK;
}

// ------------------------------------------------------------------------------------------ discriminator input
// 32 output pixels x 32 channels per block through a shared-memory transpose: sources are read along W (coalesced,
// NCHW float32), the NHWC bf16 tile is written along C.
__global__ void __launch_bounds__(256)
d_input_kernel(const float* __restrict__ a, int ca, const float* __restrict__ bsrc, int cb, __nv_bfloat16* __restrict__ out, int H,
               int W, int Ho, int Wo, int c_pad, int pool, int out_pad) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const int npix = Ho * Wo;
  const int p0 = blockIdx.x * 32;
  const size_t plane = static_cast<size_t>(H) * W;
  for (int cy = threadIdx.y; cy < 32; cy += 8) {
    const int c = c0 + cy;
    const int pp = p0 + threadIdx.x;
    float v = 0.f;
    if (pp < npix && c < ca + cb) {
      const float* src = c < ca ? a + (static_cast<size_t>(b) * ca + c) * plane
                                : bsrc + (static_cast<size_t>(b) * cb + (c - ca)) * plane;
      const int oy = pp / Wo, ox = pp - oy * Wo;
      if (!pool) {
        v = __ldg(src + static_cast<size_t>(oy) * W + ox);
      } else {
        // AvgPool2d(3, stride 2, padding 1, count_include_pad=False): mean over the in-bounds taps only
        float s = 0.f;
        int cnt = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int y = 2 * oy + dy;
          if (y < 0 || y >= H) continue;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int x = 2 * ox + dx;
            if (x < 0 || x >= W) continue;
            s += __ldg(src + static_cast<size_t>(y) * W + x);
            ++cnt;
          }
        }
        v = s / static_cast<float>(cnt);
      }
    }
    tile[cy][threadIdx.x] = v;
  }
  __syncthreads();
  const int Wst = Wo + 2 * out_pad;
  const size_t img = static_cast<size_t>(Ho + 2 * out_pad) * Wst;
  for (int py = threadIdx.y; py < 32; py += 8) {
    const int pp = p0 + py;
    const int c = c0 + threadIdx.x;
    if (pp < npix && c < c_pad) {
      const int oy = pp / Wo, ox = pp - oy * Wo;
      out[(static_cast<size_t>(b) * img + static_cast<size_t>(oy + out_pad) * Wst + ox + out_pad) * c_pad + c] =
          __float2bfloat16_rn(tile[threadIdx.x][py]);
    }
  }
}

// c_pad == 8 without pooling (the VGG19's RGB operand): one thread per pixel, one 16-byte store
__global__ void __launch_bounds__(256)
d_input8_kernel(const float* __restrict__ a, int ca, const float* __restrict__ bsrc, int cb, __nv_bfloat16* __restrict__ out, int B,
                int H, int W, int out_pad) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  const int Wst = W + 2 * out_pad;
  const size_t img = static_cast<size_t>(H + 2 * out_pad) * Wst;
  for (size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(pix / plane);
    const size_t hw = pix % plane;
    const int y = static_cast<int>(hw / W), x = static_cast<int>(hw % W);
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = 0.f;
      if (c < ca) v[c] = __ldg(a + (static_cast<size_t>(b) * ca + c) * plane + hw);
      else if (c < ca + cb) v[c] = __ldg(bsrc + (static_cast<size_t>(b) * cb + (c - ca)) * plane + hw);
    }
    *reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * img + static_cast<size_t>(y + out_pad) * Wst + x + out_pad) * 8) =
        make_uint4(d_pack2(v[0], v[1]), d_pack2(v[2], v[3]), d_pack2(v[4], v[5]), d_pack2(v[6], v[7]));
  }
}

// out[b, c, y, x] (float32 NCHW) = g0[b, y, x, c0 + c] + sum over the pooled pixels whose 3x3 window holds (y, x) of
// g1[b, i, j, c0 + c] / (in-bounds taps of that window). One thread per (pixel, channel group of <= 4).
__global__ void __launch_bounds__(256)
d_input_backward_kernel(const __nv_bfloat16* __restrict__ g0, const __nv_bfloat16* __restrict__ g1, float* __restrict__ out, int B,
                        int H, int W, int Ho, int Wo, int c_stored, int c0, int c) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  for (size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(pix / plane);
    const size_t hw = pix % plane;
    const int y = static_cast<int>(hw / W), x = static_cast<int>(hw % W);
    // pooled windows containing (y, x): 2i-1 <= y <= 2i+1
    int wi[2], wj[2], ni = 0, nj = 0;
    float fi[2], fj[2];
    if (g1 != nullptr) {
      for (int i = (y) / 2; i <= (y + 1) / 2; ++i) {
        if (i < 0 || i >= Ho || 2 * i - 1 > y || 2 * i + 1 < y) continue;
        const int lo = 2 * i - 1 < 0 ? 0 : 2 * i - 1, hi = 2 * i + 1 >= H ? H - 1 : 2 * i + 1;
        wi[ni] = i;
        fi[ni++] = static_cast<float>(hi - lo + 1);
      }
      for (int j = (x) / 2; j <= (x + 1) / 2; ++j) {
        if (j < 0 || j >= Wo || 2 * j - 1 > x || 2 * j + 1 < x) continue;
        const int lo = 2 * j - 1 < 0 ? 0 : 2 * j - 1, hi = 2 * j + 1 >= W ? W - 1 : 2 * j + 1;
        wj[nj] = j;
        fj[nj++] = static_cast<float>(hi - lo + 1);
      }
    }
    for (int ch = 0; ch < c; ++ch) {
      float v = __bfloat162float(g0[pix * c_stored + c0 + ch]);
      for (int u = 0; u < ni; ++u)
        for (int t = 0; t < nj; ++t)
          v += __bfloat162float(g1[((static_cast<size_t>(b) * Ho + wi[u]) * Wo + wj[t]) * c_stored + c0 + ch]) / (fi[u] * fj[t]);
      out[(static_cast<size_t>(b) * c + ch) * plane + hw] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------ InstanceNorm + LeakyReLU
// out is (B, H + 2 pad, W + 2 pad, C); the border is written as zeros (the next 4x4 conv's padding)
__global__ void __launch_bounds__(kDThreads)
instnorm_apply_act_kernel(const __nv_bfloat16* __restrict__ raw, const double* __restrict__ stats, __nv_bfloat16* __restrict__ out,
                          int H, int W, int C, int pad, float slope, float eps, int iters) {
  const int vpp = C >> 3;
  const int ppi = kDThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  const int npix = Hp * Wp;
  float mean[8], rstd[8];
  {
    const double inv_n = 1.0 / (static_cast<double>(H) * W);
    const double* st = stats + (static_cast<size_t>(b) * C + vec * 8) * 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double m = st[2 * j] * inv_n;
      double var = st[2 * j + 1] * inv_n - m * m;
      if (var < 0.0) var = 0.0;
      mean[j] = static_cast<float>(m);
      rstd[j] = rsqrtf(static_cast<float>(var) + eps);
    }
  }
  const uint4* raw4 = reinterpret_cast<const uint4*>(raw) + static_cast<size_t>(b) * H * W * vpp;
  uint4* out4 = reinterpret_cast<uint4*>(out) + static_cast<size_t>(b) * npix * vpp;
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters))
    for (int it = 0; it < iters; ++it) {
      const int pp = pix0 + it * ppi + psub;
      if (pp >= npix) break;
      const int ph = pp / Wp, pw = pp - ph * Wp;
      const int h = ph - pad, w = pw - pad;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (h >= 0 && h < H && w >= 0 && w < W) {
        float x[8];
        d_unpack8(__ldg(raw4 + (static_cast<size_t>(h) * W + w) * vpp + vec), x);
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float lo = (x[j] - mean[j]) * rstd[j], hi = (x[j + 1] - mean[j + 1]) * rstd[j + 1];
          lo = lo > 0.f ? lo : slope * lo;
          hi = hi > 0.f ? hi : slope * hi;
          ow[j >> 1] = d_pack2(lo, hi);
        }
        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      out4[static_cast<size_t>(pp) * vpp + vec] = o;
    }
}

// d_pre (zero-bordered by out_pad) = (g + skip) * act'(f), act' = 1 where the stored activation f > 0, else slope.
// f is (B, H + 2 f_pad, W + 2 f_pad, C); g is (B, H + 2 g_pad, W + 2 g_pad, C) (its border -- the gradient w.r.t. a zero
// padding -- is skipped), skip is dense (B,H,W,C). dbias[c] += sum of d_pre (optional).
template <bool kSkip, bool kBias>
__global__ void __launch_bounds__(kDThreads)
act_backward_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ skip, const __nv_bfloat16* __restrict__ f,
                    __nv_bfloat16* __restrict__ dpre, float* __restrict__ dbias, int H, int W, int C, int f_pad, int out_pad,
                    float slope, int iters, int g_pad) {
  __shared__ float s_red[kBias ? kDThreads : 1][9];
  const int vpp = C >> 3;
  const int ppi = kDThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int Ho = H + 2 * out_pad, Wo = W + 2 * out_pad;
  const int Wf = W + 2 * f_pad;
  const int npix = Ho * Wo;
  const int Wg = W + 2 * g_pad;
  const uint4* g4 = reinterpret_cast<const uint4*>(g) + static_cast<size_t>(b) * (H + 2 * g_pad) * Wg * vpp;
  const uint4* s4 = reinterpret_cast<const uint4*>(skip) + static_cast<size_t>(b) * H * W * vpp;
  const uint4* f4 = reinterpret_cast<const uint4*>(f) + static_cast<size_t>(b) * (H + 2 * f_pad) * Wf * vpp;
  uint4* o4 = reinterpret_cast<uint4*>(dpre) + static_cast<size_t>(b) * npix * vpp;
  float bs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bs[j] = 0.f;
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters))
    for (int it = 0; it < iters; ++it) {
      const int pp = pix0 + it * ppi + psub;
      if (pp >= npix) break;
      const int ph = pp / Wo, pw = pp - ph * Wo;
      const int h = ph - out_pad, w = pw - out_pad;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (h >= 0 && h < H && w >= 0 && w < W) {
        const size_t src = (static_cast<size_t>(h) * W + w) * vpp + vec;
        float d[8], a[8];
        d_unpack8(__ldg(g4 + (static_cast<size_t>(h + g_pad) * Wg + w + g_pad) * vpp + vec), d);
        if (kSkip) {
          float e[8];
          d_unpack8(__ldg(s4 + src), e);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] += e[j];
        }
        d_unpack8(__ldg(f4 + (static_cast<size_t>(h + f_pad) * Wf + w + f_pad) * vpp + vec), a);
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float d0 = a[j] > 0.f ? d[j] : slope * d[j], d1 = a[j + 1] > 0.f ? d[j + 1] : slope * d[j + 1];
          const uint32_t pk = d_pack2(d0, d1);
          ow[j >> 1] = pk;
          if (kBias) {
            bs[j] += __uint_as_float(pk << 16);
            bs[j + 1] += __uint_as_float(pk & 0xffff0000u);
          }
        }
        o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      }
      o4[static_cast<size_t>(pp) * vpp + vec] = o;
    }
  if (kBias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s_red[threadIdx.x][j] = bs[j];
    __syncthreads();
    for (int stride = ppi >> 1; stride >= 1; stride >>= 1) {
      if (psub < stride) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s_red[threadIdx.x][j] += s_red[threadIdx.x + stride * vpp][j];
      }
      __syncthreads();
    }
    if (psub == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dbias + vec * 8 + j, s_red[threadIdx.x][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------ L1 between two feature maps
// Both tensors have the same stored shape (zero borders and zero pad channels contribute nothing): flat sweep.
__global__ void __launch_bounds__(256)
l1_pair_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, size_t nvec, double* __restrict__ sum) {
  __shared__ float s_w[8];
  float acc = 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    d_unpack8(__ldg(a + i), x);
    d_unpack8(__ldg(b + i), y);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += fabsf(x[j] - y[j]);
    acc += s;
  }
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += static_cast<double>(s_w[w]);
    atomicAdd(sum, t);
  }
}

// out (dense B,H,W,C) = sign(a - b) * (*scale_dev) * scale_host, a / b stored with a border of `pad`
__global__ void __launch_bounds__(kDThreads)
l1_pair_backward_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, __nv_bfloat16* __restrict__ out,
                        const float* __restrict__ scale_dev, float scale_host, int H, int W, int C, int pad, int iters) {
  const int vpp = C >> 3;
  const int ppi = kDThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int bi = blockIdx.y;
  const int Wp = W + 2 * pad;
  const int npix = H * W;
  const float sc = (scale_dev != nullptr ? __ldg(scale_dev) : 1.f) * scale_host;
  const size_t img = static_cast<size_t>(H + 2 * pad) * Wp * vpp;
  const uint4* a4 = reinterpret_cast<const uint4*>(a) + bi * img;
  const uint4* b4 = reinterpret_cast<const uint4*>(b) + bi * img;
  uint4* o4 = reinterpret_cast<uint4*>(out) + static_cast<size_t>(bi) * npix * vpp;
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters))
    for (int it = 0; it < iters; ++it) {
      const int pp = pix0 + it * ppi + psub;
      if (pp >= npix) break;
      const int h = pp / W, w = pp - h * W;
      const size_t src = (static_cast<size_t>(h + pad) * Wp + w + pad) * vpp + vec;
      float x[8], y[8];
      d_unpack8(__ldg(a4 + src), x);
      d_unpack8(__ldg(b4 + src), y);
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float e0 = x[j] - y[j], e1 = x[j + 1] - y[j + 1];
        ow[j >> 1] = d_pack2(e0 > 0.f ? sc : (e0 < 0.f ? -sc : 0.f), e1 > 0.f ? sc : (e1 < 0.f ? -sc : 0.f));
      }
      o4[static_cast<size_t>(pp) * vpp + vec] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ------------------------------------------------------------------------------------------ MaxPool2d(2, 2)
// x (B, H + 2 in_pad, W + 2 in_pad, C) -> y (B, H/2 + 2 out_pad, W/2 + 2 out_pad, C), border written as zeros
__global__ void __launch_bounds__(kDThreads)
maxpool2x2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int H, int W, int C, int in_pad, int out_pad,
                  int iters) {
  const int vpp = C >> 3;
  const int ppi = kDThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int Ho = H / 2, Wo = W / 2;
  const int Hs = Ho + 2 * out_pad, Ws = Wo + 2 * out_pad;
  const int Wi = W + 2 * in_pad;
  const int npix = Hs * Ws;
  const uint4* x4 = reinterpret_cast<const uint4*>(x) + static_cast<size_t>(b) * (H + 2 * in_pad) * Wi * vpp;
  uint4* y4 = reinterpret_cast<uint4*>(y) + static_cast<size_t>(b) * npix * vpp;
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters))
    for (int it = 0; it < iters; ++it) {
      const int pp = pix0 + it * ppi + psub;
      if (pp >= npix) break;
      const int ph = pp / Ws, pw = pp - ph * Ws;
      const int oh = ph - out_pad, ow_ = pw - out_pad;
      uint4 o = make_uint4(0, 0, 0, 0);
      if (oh >= 0 && oh < Ho && ow_ >= 0 && ow_ < Wo) {
        float m[8];
        const size_t base = (static_cast<size_t>(2 * oh + in_pad) * Wi + 2 * ow_ + in_pad) * vpp + vec;
        d_unpack8(__ldg(x4 + base), m);
        const size_t offs[3] = {static_cast<size_t>(vpp), static_cast<size_t>(Wi) * vpp, static_cast<size_t>(Wi + 1) * vpp};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          float v[8];
          d_unpack8(__ldg(x4 + base + offs[k]), v);
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
        }
        o = make_uint4(d_pack2(m[0], m[1]), d_pack2(m[2], m[3]), d_pack2(m[4], m[5]), d_pack2(m[6], m[7]));
      }
      y4[static_cast<size_t>(pp) * vpp + vec] = o;
    }
}

// dx (dense B,H,W,C) = the pooled gradient g ((B, H/2 + 2 g_pad, W/2 + 2 g_pad, C), border skipped) routed to the FIRST position of each 2x2 window that
// holds the maximum (row-major scan order, like ATen's max_pool2d backward); other positions get zero.
__global__ void __launch_bounds__(kDThreads)
maxpool2x2_backward_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dx,
                           int H, int W, int C, int in_pad, int iters, int g_pad) {
  const int vpp = C >> 3;
  const int ppi = kDThreads / vpp;
  const int vec = threadIdx.x % vpp;
  const int psub = threadIdx.x / vpp;
  const int b = blockIdx.y;
  const int Ho = H / 2, Wo = W / 2;
  const int Wi = W + 2 * in_pad;
  const int npix = Ho * Wo;
  const uint4* x4 = reinterpret_cast<const uint4*>(x) + static_cast<size_t>(b) * (H + 2 * in_pad) * Wi * vpp;
  const int Wg = Wo + 2 * g_pad;
  const uint4* g4 = reinterpret_cast<const uint4*>(g) + static_cast<size_t>(b) * (Ho + 2 * g_pad) * Wg * vpp;
  uint4* d4 = reinterpret_cast<uint4*>(dx) + static_cast<size_t>(b) * H * W * vpp;
  for (int pix0 = blockIdx.x * (ppi * iters); pix0 < npix; pix0 += gridDim.x * (ppi * iters))
    for (int it = 0; it < iters; ++it) {
      const int pp = pix0 + it * ppi + psub;
      if (pp >= npix) break;
      const int oh = pp / Wo, ow_ = pp - oh * Wo;
      const size_t base = (static_cast<size_t>(2 * oh + in_pad) * Wi + 2 * ow_ + in_pad) * vpp + vec;
      float v[4][8], gg[8];
      d_unpack8(__ldg(x4 + base), v[0]);
      d_unpack8(__ldg(x4 + base + vpp), v[1]);
      d_unpack8(__ldg(x4 + base + static_cast<size_t>(Wi) * vpp), v[2]);
      d_unpack8(__ldg(x4 + base + static_cast<size_t>(Wi + 1) * vpp), v[3]);
      d_unpack8(__ldg(g4 + (static_cast<size_t>(oh + g_pad) * Wg + ow_ + g_pad) * vpp + vec), gg);
      float o[4][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int arg = 0;
        float m = v[0][j];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (v[k][j] > m) {
            m = v[k][j];
            arg = k;
          }
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][j] = k == arg ? gg[j] : 0.f;
      }
      const size_t d0 = (static_cast<size_t>(2 * oh) * W + 2 * ow_) * vpp + vec;
      const size_t doff[4] = {0, static_cast<size_t>(vpp), static_cast<size_t>(W) * vpp, static_cast<size_t>(W + 1) * vpp};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        d4[d0 + doff[k]] = make_uint4(d_pack2(o[k][0], o[k][1]), d_pack2(o[k][2], o[k][3]), d_pack2(o[k][4], o[k][5]),
                                      d_pack2(o[k][6], o[k][7]));
    }
}

// ------------------------------------------------------------------------------------------ stored NHWC -> float32 NCHW
__global__ void nhwc_pad_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int C, int H, int W, int pad,
                                            int c_stored) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32;
  const size_t npix = static_cast<size_t>(H) * W;
  const size_t p0 = static_cast<size_t>(blockIdx.x) * 32;
  const int Wp = W + 2 * pad;
  const size_t img = static_cast<size_t>(H + 2 * pad) * Wp;
  {
    const size_t pp = p0 + threadIdx.y;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (pp < npix && c < C) {
      const int h = static_cast<int>(pp / W), w = static_cast<int>(pp % W);
      v = __bfloat162float(x[(static_cast<size_t>(b) * img + static_cast<size_t>(h + pad) * Wp + w + pad) * c_stored + c]);
    }
    tile[threadIdx.y][threadIdx.x] = v;
  }
  __syncthreads();
  {
    const size_t pp = p0 + threadIdx.x;
    const int c = c0 + threadIdx.y;
    if (pp < npix && c < C) y[(static_cast<size_t>(b) * C + c) * npix + pp] = tile[threadIdx.x][threadIdx.y];
  }
}

// few channels (the 3-channel image gradient out of a 64-channel data-gradient tensor): one thread per pixel reads the
// first 16 bytes of the pixel and writes up to 8 planes, coalesced along W
__global__ void __launch_bounds__(256)
nhwc_pad_to_nchw_f32_few_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int B, int C, int H, int W, int pad,
                                int c_stored) {
  const size_t plane = static_cast<size_t>(H) * W;
  const size_t total = static_cast<size_t>(B) * plane;
  const int Wp = W + 2 * pad;
  const size_t img = static_cast<size_t>(H + 2 * pad) * Wp;
  for (size_t pix = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
       pix += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(pix / plane);
    const size_t hw = pix % plane;
    const int h = static_cast<int>(hw / W), w = static_cast<int>(hw % W);
    float f[8];
    d_unpack8(__ldg(reinterpret_cast<const uint4*>(x + (static_cast<size_t>(b) * img + static_cast<size_t>(h + pad) * Wp + w + pad) * c_stored)), f);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c < C) y[(static_cast<size_t>(b) * C + c) * plane + hw] = f[c];
  }
}

}  // namespace jpdse

using namespace jpdse;

extern "C" int jpdse_d_input(const float* a, int ca, const float* b, int cb, void* out, int batch, int height, int width, int c_pad,
                             int pool, int out_pad, void* stream) {
  if (a == nullptr || out == nullptr || (cb > 0 && b == nullptr)) return fail(JPDSE_ERR_INVALID, "d_input: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || ca <= 0 || cb < 0 || ca + cb > c_pad || c_pad % 8 || out_pad < 0)
    return fail(JPDSE_ERR_INVALID, "d_input: bad sizes");
  const int Ho = pool ? (height - 1) / 2 + 1 : height, Wo = pool ? (width - 1) / 2 + 1 : width;
  if (c_pad == 8 && !pool && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const size_t total = static_cast<size_t>(batch) * height * width;
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(num_sms()) * 32;
    if (blocks > cap) blocks = cap;
    d_input8_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        a, ca, b, cb, static_cast<__nv_bfloat16*>(out), batch, height, width, out_pad);
    return check_launch("d_input8_kernel");
  }
  dim3 grid((Ho * Wo + 31) / 32, (c_pad + 31) / 32, batch);
  d_input_kernel<<<grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(a, ca, b, cb, static_cast<__nv_bfloat16*>(out), height,
                                                                            width, Ho, Wo, c_pad, pool ? 1 : 0, out_pad);
  return check_launch("d_input_kernel");
}

extern "C" int jpdse_d_input_backward(const void* g0, const void* g1, float* out, int batch, int height, int width, int c_stored,
                                      int c0, int c, void* stream) {
  if (g0 == nullptr || out == nullptr) return fail(JPDSE_ERR_INVALID, "d_input_backward: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || c <= 0 || c0 < 0 || c0 + c > c_stored)
    return fail(JPDSE_ERR_INVALID, "d_input_backward: bad sizes");
  const size_t total = static_cast<size_t>(batch) * height * width;
  size_t blocks = (total + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  d_input_backward_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(g0), static_cast<const __nv_bfloat16*>(g1), out, batch, height, width, (height - 1) / 2 + 1,
      (width - 1) / 2 + 1, c_stored, c0, c);
  return check_launch("d_input_backward_kernel");
}

extern "C" int jpdse_instnorm_apply_act(const void* raw, const double* stats, void* out, int batch, int height, int width,
                                        int channels, int out_pad, float slope, float eps, void* stream) {
  if (raw == nullptr || stats == nullptr || out == nullptr) return fail(JPDSE_ERR_INVALID, "instnorm_apply_act: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || out_pad < 0) return fail(JPDSE_ERR_INVALID, "instnorm_apply_act: bad sizes");
  int rc = check_vec_channels(channels, "instnorm_apply_act");
  if (rc != JPDSE_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(raw) | reinterpret_cast<uintptr_t>(out)) & 15)
    return fail(JPDSE_ERR_INVALID, "instnorm_apply_act: pointers must be 16-byte aligned");
  const int vpp = channels / 8, ppi = kDThreads / vpp;
  int iters;
  dim3 grid = sweep_grid(static_cast<long long>(height + 2 * out_pad) * (width + 2 * out_pad), ppi, batch, &iters);
  instnorm_apply_act_kernel<<<grid, kDThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(raw), stats, static_cast<__nv_bfloat16*>(out), height, width, channels, out_pad, slope, eps,
      iters);
  return check_launch("instnorm_apply_act_kernel");
}

extern "C" int jpdse_act_backward(const void* g, int g_pad, const void* skip, const void* f, void* d_pre, float* dbias, int batch,
                                  int height, int width, int channels, int f_pad, int out_pad, float slope, void* stream_v) {
  if (g == nullptr || f == nullptr || d_pre == nullptr) return fail(JPDSE_ERR_INVALID, "act_backward: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || f_pad < 0 || out_pad < 0 || g_pad < 0)
    return fail(JPDSE_ERR_INVALID, "act_backward: bad sizes");
  int rc = check_vec_channels(channels, "act_backward");
  if (rc != JPDSE_OK) return rc;
  if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(f) |
       reinterpret_cast<uintptr_t>(d_pre)) & 15)
    return fail(JPDSE_ERR_INVALID, "act_backward: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const int vpp = channels / 8, ppi = kDThreads / vpp;
  int iters;
  dim3 grid = sweep_grid(static_cast<long long>(height + 2 * out_pad) * (width + 2 * out_pad), ppi, batch, &iters);
  const __nv_bfloat16* gp = static_cast<const __nv_bfloat16*>(g);
  const __nv_bfloat16* sp = static_cast<const __nv_bfloat16*>(skip);
  const __nv_bfloat16* fp = static_cast<const __nv_bfloat16*>(f);
  __nv_bfloat16* dp = static_cast<__nv_bfloat16*>(d_pre);
  if (skip && dbias)
    act_backward_kernel<true, true><<<grid, kDThreads, 0, stream>>>(gp, sp, fp, dp, dbias, height, width, channels, f_pad, out_pad, slope, iters, g_pad);
  else if (skip)
    act_backward_kernel<true, false><<<grid, kDThreads, 0, stream>>>(gp, sp, fp, dp, dbias, height, width, channels, f_pad, out_pad, slope, iters, g_pad);
  else if (dbias)
    act_backward_kernel<false, true><<<grid, kDThreads, 0, stream>>>(gp, sp, fp, dp, dbias, height, width, channels, f_pad, out_pad, slope, iters, g_pad);
  else
    act_backward_kernel<false, false><<<grid, kDThreads, 0, stream>>>(gp, sp, fp, dp, dbias, height, width, channels, f_pad, out_pad, slope, iters, g_pad);
  return check_launch("act_backward_kernel");
}

extern "C" int jpdse_l1_pair(const void* a, const void* b, size_t n_elements, double* sum, void* stream) {
  if (a == nullptr || b == nullptr || sum == nullptr) return fail(JPDSE_ERR_INVALID, "l1_pair: NULL pointer");
  if (n_elements % 8) return fail(JPDSE_ERR_INVALID, "l1_pair: element count must be a multiple of 8");
  if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15)
    return fail(JPDSE_ERR_INVALID, "l1_pair: pointers must be 16-byte aligned");
  const size_t nvec = n_elements / 8;
  size_t blocks = (nvec + 256 * 8 - 1) / (256 * 8);
  const size_t cap = static_cast<size_t>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  l1_pair_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), nvec, sum);
  return check_launch("l1_pair_kernel");
}

extern "C" int jpdse_l1_pair_backward(const void* a, const void* b, void* out, const float* scale_dev, float scale_host, int batch,
                                      int height, int width, int channels, int pad, void* stream) {
  if (a == nullptr || b == nullptr || out == nullptr) return fail(JPDSE_ERR_INVALID, "l1_pair_backward: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || pad < 0) return fail(JPDSE_ERR_INVALID, "l1_pair_backward: bad sizes");
  int rc = check_vec_channels(channels, "l1_pair_backward");
  if (rc != JPDSE_OK) return rc;
  const int vpp = channels / 8, ppi = kDThreads / vpp;
  int iters;
  dim3 grid = sweep_grid(static_cast<long long>(height) * width, ppi, batch, &iters);
  l1_pair_backward_kernel<<<grid, kDThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), static_cast<__nv_bfloat16*>(out), scale_dev,
      scale_host, height, width, channels, pad, iters);
  return check_launch("l1_pair_backward_kernel");
}

extern "C" int jpdse_maxpool2x2(const void* x, void* y, int batch, int height, int width, int channels, int in_pad, int out_pad,
                                void* stream) {
  if (x == nullptr || y == nullptr) return fail(JPDSE_ERR_INVALID, "maxpool2x2: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || (height & 1) || (width & 1) || in_pad < 0 || out_pad < 0)
    return fail(JPDSE_ERR_INVALID, "maxpool2x2: bad sizes (even height / width required)");
  int rc = check_vec_channels(channels, "maxpool2x2");
  if (rc != JPDSE_OK) return rc;
  const int vpp = channels / 8, ppi = kDThreads / vpp;
  int iters;
  dim3 grid = sweep_grid(static_cast<long long>(height / 2 + 2 * out_pad) * (width / 2 + 2 * out_pad), ppi, batch, &iters);
  maxpool2x2_kernel<<<grid, kDThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), height, width, channels, in_pad, out_pad, iters);
  return check_launch("maxpool2x2_kernel");
}

extern "C" int jpdse_maxpool2x2_backward(const void* x, const void* g, int g_pad, void* dx, int batch, int height, int width,
                                         int channels, int in_pad, void* stream) {
  if (x == nullptr || g == nullptr || dx == nullptr) return fail(JPDSE_ERR_INVALID, "maxpool2x2_backward: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || (height & 1) || (width & 1) || in_pad < 0)
    return fail(JPDSE_ERR_INVALID, "maxpool2x2_backward: bad sizes");
  int rc = check_vec_channels(channels, "maxpool2x2_backward");
  if (rc != JPDSE_OK) return rc;
  const int vpp = channels / 8, ppi = kDThreads / vpp;
  int iters;
  dim3 grid = sweep_grid(static_cast<long long>(height / 2) * (width / 2), ppi, batch, &iters);
  maxpool2x2_backward_kernel<<<grid, kDThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(g), static_cast<__nv_bfloat16*>(dx), height, width,
      channels, in_pad, iters, g_pad);
  return check_launch("maxpool2x2_backward_kernel");
}

// ------------------------------------------------------------------------------------------ 1-channel output conv
namespace jpdse {
// out[b,y,x] = bias + sum_t z[b][t][(y + t/4) * Ws + x + t%4]: one thread per output pixel, 16 plane-coalesced loads
__global__ void __launch_bounds__(256)
patch_out_gather_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ out, int B, int H, int W) {
  const int Ho = H + 1, Wo = W + 1, Ws = W + 4;
  const size_t plane = static_cast<size_t>(H + 4) * Ws;
  const size_t total = static_cast<size_t>(B) * Ho * Wo;
  const float b0 = bias != nullptr ? __ldg(bias) : 0.f;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % Wo);
    const int y = static_cast<int>((i / Wo) % Ho);
    const int b = static_cast<int>(i / (static_cast<size_t>(Wo) * Ho));
    const float* zb = z + static_cast<size_t>(b) * 16 * plane + static_cast<size_t>(y) * Ws + x;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) acc += __ldg(zb + t * plane + (t >> 2) * Ws + (t & 3));
    out[i] = acc + b0;
  }
}

// dz[b][py][px][t] = dout[b][py - t/4][px - t%4] (0 outside, and for the channels t >= 16): thread = (stored pixel, 8-ch vector)
__global__ void __launch_bounds__(256)
patch_out_scatter_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dz, int B, int H, int W, int c_pad) {
  const int Ho = H + 1, Wo = W + 1, Hs = H + 4, Ws = W + 4;
  const int vpp = c_pad >> 3;
  const size_t total = static_cast<size_t>(B) * Hs * Ws * vpp;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % vpp);
    const size_t pix = i / vpp;
    const int px = static_cast<int>(pix % Ws);
    const int py = static_cast<int>((pix / Ws) % Hs);
    const int b = static_cast<int>(pix / (static_cast<size_t>(Ws) * Hs));
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = 8 * k + j;
      const int y = py - (t >> 2), x = px - (t & 3);
      v[j] = (t < 16 && y >= 0 && y < Ho && x >= 0 && x < Wo) ? __ldg(dout + (static_cast<size_t>(b) * Ho + y) * Wo + x) : 0.f;
    }
    reinterpret_cast<uint4*>(dz)[i] = make_uint4(d_pack2(v[0], v[1]), d_pack2(v[2], v[3]), d_pack2(v[4], v[5]), d_pack2(v[6], v[7]));
  }
}

}  // namespace jpdse
using namespace jpdse;

extern "C" int jpdse_patch_out_gather(const float* z, const float* bias, float* out, int batch, int height, int width, void* stream) {
  if (z == nullptr || out == nullptr) return fail(JPDSE_ERR_INVALID, "patch_out_gather: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0) return fail(JPDSE_ERR_INVALID, "patch_out_gather: bad sizes");
  const size_t total = static_cast<size_t>(batch) * (height + 1) * (width + 1);
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
  patch_out_gather_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(z, bias, out, batch, height, width);
  return check_launch("patch_out_gather_kernel");
}

extern "C" int jpdse_patch_out_scatter(const float* dout, void* dz, int batch, int height, int width, int c_pad, void* stream) {
  if (dout == nullptr || dz == nullptr) return fail(JPDSE_ERR_INVALID, "patch_out_scatter: NULL pointer");
  if (batch <= 0 || height <= 0 || width <= 0 || c_pad < 16 || c_pad % 8) return fail(JPDSE_ERR_INVALID, "patch_out_scatter: bad sizes");
  if (reinterpret_cast<uintptr_t>(dz) & 15) return fail(JPDSE_ERR_INVALID, "patch_out_scatter: dz must be 16-byte aligned");
  const size_t total = static_cast<size_t>(batch) * (height + 4) * (width + 4) * (c_pad / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > static_cast<size_t>(num_sms()) * 16) blocks = static_cast<size_t>(num_sms()) * 16;
  patch_out_scatter_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      dout, static_cast<__nv_bfloat16*>(dz), batch, height, width, c_pad);
  return check_launch("patch_out_scatter_kernel");
}

extern "C" int jpdse_nhwc_pad_to_nchw_f32(const void* x, float* y, int batch, int channels, int height, int width, int pad,
                                          int c_stored, void* stream) {
  if (x == nullptr || y == nullptr) return fail(JPDSE_ERR_INVALID, "nhwc_pad_to_nchw_f32: NULL pointer");
  if (batch <= 0 || channels <= 0 || height <= 0 || width <= 0 || pad < 0 || channels > c_stored)
    return fail(JPDSE_ERR_INVALID, "nhwc_pad_to_nchw_f32: bad sizes");
  if (channels <= 8 && c_stored % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const size_t total = static_cast<size_t>(batch) * height * width;
    size_t blocks = (total + 255) / 256;
    const size_t cap = static_cast<size_t>(num_sms()) * 32;
    if (blocks > cap) blocks = cap;
    nhwc_pad_to_nchw_f32_few_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), y, batch, channels, height, width, pad, c_stored);
    return check_launch("nhwc_pad_to_nchw_f32_few_kernel");
  }
  dim3 grid((height * width + 31) / 32, (channels + 31) / 32, batch);
  nhwc_pad_to_nchw_f32_kernel<<<grid, dim3(32, 32), 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), y, channels, height, width, pad, c_stored);
  return check_launch("nhwc_pad_to_nchw_f32_kernel");
}
