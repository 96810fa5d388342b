// Weight gradients of the generator's convolutions on sm_100a (the wgrad third of the backward pass that
// `loss_G.backward()` triggers in the reference: ctu/trainers/pix2pixHD_trainer.py:69 through
// ctu/models/pix2pixHD_networks/networks.py:210,215,244,246,283-299).
//
//   dW[m][n][tap] = sum over pixels  P[pixel][m] * Q[pixel @ tap][n]
//
// P is the tensor indexed on the GEMM's own pixel grid (the output gradient for Conv2d, the forward input for
// ConvTranspose2d), Q the one addressed through the filter taps. Both are NHWC bf16, i.e. *pixel-major*: a TMA box
// {64 channels, 64 pixels} lands in shared memory as 64 rows of 128 bytes, which is exactly the MN-major SWIZZLE_128B
// operand layout of tcgen05.mma (tools/umma_mnmajor_test.cu pins the descriptor: LBO = pitch between 64-channel
// atoms, SBO = 1024, K advances 16 pixel rows = 2048 bytes). The reduction dimension K is the pixel index, so no
// transposed copy of either activation is ever made.
//
// One work item = (k-split, tap group, m tile, n tile): M = 128 rows (two A boxes), N = 64 * nb columns (nb <= 4
// B boxes). A box carries its own TMA coordinate offsets and output (tap, channel) mapping, which lets the host
// express, with one kernel:
//   * 3x3 stride-1 (reflect-padded input), 3x3 stride-2 and ConvTranspose (through the (col-pair, row-parity)
//     5-D view the forward kernel uses), 1x1;
//   * narrow layers: when Q has only 64 / 128 channels, the B boxes of one item are DIFFERENT TAPS, so P is
//     fetched once per 3-4 taps instead of once per tap;
//   * the 7x7 stem: Q = the overlapping-window view (7*40 contiguous elements under a filter row), the two A
//     boxes are the output gradient shifted by two consecutive filter rows;
//   * the 7x7 head: P = padded input rows shifted by two filter rows, Q = a window over 8-channel gradient pixels.
// Partial sums leave TMEM through vectorised fp32 reductions (red.global.add.v4.f32) into a [tap][m][n] workspace
// (split-K across CTAs for the full-resolution layers whose M x N is tiny); a finalize kernel permutes the
// workspace into the torch weight layout.
//
// Warp roles as in conv_igemm.cu: warps 0..3 / 4..7 epilogue warpgroups (one per TMEM accumulator), warp 8 TMA
// producer, warp 9 TMEM owner + MMA issuer.
#include <cuda_bf16.h>

#include <cstring>

#include "common.cuh"
#include "conv_shared.cuh"
#include "ptx.cuh"

namespace jpdse {

constexpr int kWgThreads = 320;
constexpr int kWgProducerWarp = 8;
constexpr int kWgMmaWarp = 9;
constexpr int kWgMaxGroups = 16;  // 4x4 PatchGAN convs: one tap per group when Q has >= 192 channels
constexpr int kWgBox = 64 * 64 * 2;          // one {64 channels, 64 pixels} box, bytes
constexpr int kWgStageBytes = 6 * kWgBox;    // 2 A boxes + up to 4 B boxes
constexpr int kWgStages = 4;
// "tall" items (WgParams::tall): M = 256 = both TMEM accumulators of one item, 4 A boxes + 4 B boxes a stage, three
// stages in the same shared memory. 64 KB of operands feed 8 MMAs instead of 48 KB feeding 4: the kernel is bound by
// the L2 -> SM fabric (profiles/r2_ncu_full_wgrad_b2_raw.csv: 11.6 TB/s, tensor pipe 48 %), so bytes per MMA is the lever.
constexpr int kWgTallStageBytes = 8 * kWgBox;
constexpr int kWgTallStages = 3;
constexpr int kWgSmemBytes = 1024 + kWgStages * kWgStageBytes + 256;
static_assert(kWgTallStages * kWgTallStageBytes <= kWgStages * kWgStageBytes, "tall stages must fit the same allocation");

struct WgView {
  int rank, dim_w, dim_h, dim_b;
};

struct WgParams {
  int batch, nbw, nbh;   // k-blocks (64 pixels) per grid row / per image column of blocks
  int kw_cols, kh_rows;  // pixel shape of a k-block (kw_cols * kh_rows == 64)
  int n_groups, m_tiles, n_tiles, splits;
  int nb;                // B boxes per item
  int a_tile_stride, b_tile_stride;      // TMA channel-coordinate step per m / n tile
  int m_out_stride, n_out_stride;        // workspace row / column step per m / n tile
  WgView a, b;
  int a_off[kWgMaxGroups][4][5];   // boxes 2, 3: tall items only
  int b_off[kWgMaxGroups][4][5];
  short a_tap[kWgMaxGroups][4], b_tap[kWgMaxGroups][4];  // output tap contribution, < 0 = box unused
  short a_ch[kWgMaxGroups][4], b_ch[kWgMaxGroups][4];    // workspace row / column offset of the box
  int m_tot, n_tot;      // workspace extents: ws[tap][m_tot][n_tot]
  int direct;            // 1: no split-K -> every workspace element has exactly one writer: plain stores, no memset
  int tall;              // 1: M = 256 per item = four A boxes (generic: 256 P channels; stem: four filter rows)
  float* ws;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(kWgBox >> 4) << 16;  // LBO: pitch between 64-channel atoms
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: pitch between 8-pixel row groups
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// Both views keep the pixel column in dim 1; rank 4 = {c, w, h, b}, rank 5 = {c, w, parity, h, b}. Coordinates are
// built with straight-line code: the single producer thread must never touch local memory (a dynamically indexed
// coordinate array costs ~300 cycles of dependent local loads per box and serialises the whole pipeline).
__device__ __forceinline__ void wg_load_box(const CUtensorMap* tm, int rank, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4, int pw, int ph, int b) {
  if (rank == 4)
    tma_load_4d(tm, bar, dst, c0, c1 + pw, c2 + ph, c3 + b);
  else
    tma_load_5d(tm, bar, dst, c0, c1 + pw, c2, c3 + ph, c4 + b);
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kWgStages;
  uint64_t* tfull_bar = bars + 2 * kWgStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kWgProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == kWgMmaWarp) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kb_total = p.batch * p.nbh * p.nbw;
  const int per_split = p.n_groups * p.m_tiles * p.n_tiles;
  const int items = per_split * p.splits;
  const int bn = 64 * p.nb;
  const int n_a = p.tall ? 4 : 2;  // A boxes per stage
  const int n_stages = p.tall ? kWgTallStages : kWgStages;
  const int stage_bytes = p.tall ? kWgTallStageBytes : kWgStageBytes;
  const uint32_t stage_tx = static_cast<uint32_t>((n_a + p.nb) * kWgBox);

  if (warp == kWgProducerWarp) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int s = item / per_split;
        int rem = item - s * per_split;
        const int nt = rem % p.n_tiles;
        rem /= p.n_tiles;
        const int mt = rem % p.m_tiles;
        const int g = rem / p.m_tiles;
        const int kb0 = static_cast<int>(static_cast<long long>(kb_total) * s / p.splits);
        const int kb1 = static_cast<int>(static_cast<long long>(kb_total) * (s + 1) / p.splits);
        // per-item box coordinates in registers
        int ao[4][5], bo[4][5];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int q = 0; q < 5; ++q) ao[j][q] = p.a_off[g][j][q];
          ao[j][0] += mt * p.a_tile_stride;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int q = 0; q < 5; ++q) bo[j][q] = p.b_off[g][j][q];
          bo[j][0] += nt * p.b_tile_stride;
        }
        const int per_img = p.nbh * p.nbw;
        int b = kb0 / per_img;
        int r2 = kb0 - b * per_img;
        int rh = r2 / p.nbw;
        int rw = r2 - rh * p.nbw;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int pw = rw * p.kw_cols;
          const int ph = rh * p.kh_rows;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
#pragma unroll
          for (int j = 0; j < 2; ++j)
            wg_load_box(&tm_a, p.a.rank, &full_bar[stage], st + j * kWgBox, ao[j][0], ao[j][1], ao[j][2], ao[j][3], ao[j][4], pw, ph, b);
          if (p.tall) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              wg_load_box(&tm_a, p.a.rank, &full_bar[stage], st + (2 + j) * kWgBox, ao[2 + j][0], ao[2 + j][1], ao[2 + j][2], ao[2 + j][3], ao[2 + j][4], pw, ph, b);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < p.nb)
              wg_load_box(&tm_b, p.b.rank, &full_bar[stage], st + (n_a + j) * kWgBox, bo[j][0], bo[j][1], bo[j][2], bo[j][3], bo[j][4], pw, ph, b);
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++rw == p.nbw) {
            rw = 0;
            if (++rh == p.nbh) {
              rh = 0;
              ++b;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWgMmaWarp) {
    if (elect_one()) {
      // A and B both MN-major (bits 15 / 16)
      const uint32_t idesc = umma_idesc_bf16(128, bn) | (1u << 15) | (1u << 16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int s = item / per_split;
        const int kb0 = static_cast<int>(static_cast<long long>(kb_total) * s / p.splits);
        const int kb1 = static_cast<int>(static_cast<long long>(kb_total) * (s + 1) / p.splits);
        if (p.tall) {
          // both accumulators belong to this item: rows 0..127 in columns 0..255, rows 128..255 in columns 256..511
          mbar_wait(&tempty_bar[0], acc_phase ^ 1);
          mbar_wait(&tempty_bar[1], acc_phase ^ 1);
          tc_fence_after();
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * stage_bytes);
            const uint64_t adesc0 = umma_desc_mn_sw128(sa);
            const uint64_t adesc1 = umma_desc_mn_sw128(sa + 2 * kWgBox);
            const uint64_t bdesc = umma_desc_mn_sw128(sa + 4 * kWgBox);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t accum = (kb != kb0 || k != 0) ? 1u : 0u;
              umma_bf16<1>(tmem_base, adesc0 + static_cast<uint64_t>(k * 128), bdesc + static_cast<uint64_t>(k * 128), idesc, accum);
              umma_bf16<1>(tmem_base + 256u, adesc1 + static_cast<uint64_t>(k * 128), bdesc + static_cast<uint64_t>(k * 128), idesc, accum);
            }
            umma_commit(&empty_bar[stage]);
            if (kb == kb1 - 1) {
              umma_commit(&tfull_bar[0]);
              umma_commit(&tfull_bar[1]);
            }
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (kb1 <= kb0) {
            umma_commit(&tfull_bar[0]);
            umma_commit(&tfull_bar[1]);
          }
          acc_phase ^= 1;
          continue;
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kWgStageBytes);
          const uint64_t adesc = umma_desc_mn_sw128(sa);
          const uint64_t bdesc = umma_desc_mn_sw128(sa + 2 * kWgBox);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 pixel rows = 2048 bytes per MMA
            umma_bf16<1>(d_tmem, adesc + static_cast<uint64_t>(k * 128), bdesc + static_cast<uint64_t>(k * 128), idesc,
                         (kb != kb0 || k != 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kWgStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (kb1 <= kb0) umma_commit(&tfull_bar[acc]);  // empty split: nothing to add, still hand the buffer over
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int group = warp >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int ja = (row >> 6) + (p.tall ? 2 * group : 0);  // tall items: group g drains A boxes 2g, 2g + 1
    const int acc = group;
    uint32_t acc_phase = 0;
    int item_i = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x, ++item_i) {
      if (!p.tall && (item_i & 1) != group) continue;  // tall items: group g drains rows 128 g .. 128 g + 127 of every item
      const int s = item / per_split;
      int rem = item - s * per_split;
      const int nt = rem % p.n_tiles;
      rem /= p.n_tiles;
      const int mt = rem % p.m_tiles;
      const int g = rem / p.m_tiles;
      const int kb0 = static_cast<int>(static_cast<long long>(kb_total) * s / p.splits);
      const int kb1 = static_cast<int>(static_cast<long long>(kb_total) * (s + 1) / p.splits);
      mbar_wait_parked(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int a_tap = p.a_tap[g][ja];
      const int m = mt * p.m_out_stride + p.a_ch[g][ja] + (row & 63);
      if (kb1 > kb0) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256);
        for (int ch = 0; ch < 2 * p.nb; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + ch * 32, v);
          tmem_ld_wait();
          const int jb = ch >> 1;
          const int b_tap = p.b_tap[g][jb];
          if (a_tap < 0 || b_tap < 0 || m >= p.m_tot) continue;
          const int n0 = nt * p.n_out_stride + p.b_ch[g][jb] + (ch & 1) * 32;
          float* dst = p.ws + (static_cast<size_t>(a_tap + b_tap) * p.m_tot + m) * p.n_tot + n0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (n0 + j < p.n_tot) {  // n_tot is a multiple of 4
              if (p.direct)
                *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
              else
                red_add_v4(dst + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                           __uint_as_float(v[j + 3]));
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWgMmaWarp) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// workspace [tap][m_tot][n_tot] fp32 -> torch weight layout, mode:
//   0: dw[m][n][tap]                    (Conv2d (Cout,Cin,k,k): m = co, n = ci; ConvTranspose2d (Cin,Cout,k,k): m = ci, n = co)
//   1: 7x7 stem   ws[kh][co][e = kw*cin_st + ci]  -> dw[co][ci][kh][kw]
//   2: 7x7 head   ws[kh][ci][e = j*8 + co], kw = 6 - j -> dw[co][ci][kh][kw]
struct WgFinalize {
  int mode, taps, m_tot, n_tot, m_real, n_real, cin_st, accumulate;
};

__global__ void __launch_bounds__(256) wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw, WgFinalize f) {
  if (f.mode == 0) {
    // a block owns 256 consecutive (m, n) pairs: coalesced reads along n in every tap plane, a transpose through shared
    // memory (row pitch `taps` is odd or 16: 9 is conflict-free), then the block's 256 * taps output floats -- one
    // contiguous run of dw -- leave as float4 rows (scalar 36-byte-strided stores ran this kernel at half the HBM rate)
    __shared__ __align__(16) float s_t[256 * 16];
    const size_t mn_total = static_cast<size_t>(f.m_real) * f.n_real;
    const size_t plane = static_cast<size_t>(f.m_tot) * f.n_tot;
    const int taps = f.taps;
    for (size_t i0 = static_cast<size_t>(blockIdx.x) * 256; i0 < mn_total; i0 += static_cast<size_t>(gridDim.x) * 256) {
      const size_t i = i0 + threadIdx.x;
      if (i < mn_total) {
        const int n = static_cast<int>(i % f.n_real), m = static_cast<int>(i / f.n_real);
        const float* src = ws + static_cast<size_t>(m) * f.n_tot + n;
        for (int t = 0; t < taps; ++t) s_t[threadIdx.x * taps + t] = src[t * plane];
      }
      __syncthreads();
      const size_t left = mn_total - i0;
      const int count = static_cast<int>(left < 256 ? left : 256) * taps;  // floats of this block's run
      float* dst = dw + i0 * taps;  // i0 * taps * 4 bytes is a multiple of 1024
      if ((count & 3) == 0) {
        for (int j = threadIdx.x * 4; j < count; j += 256 * 4) {
          float4 v = *reinterpret_cast<const float4*>(s_t + j);
          if (f.accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(dst + j);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
          }
          *reinterpret_cast<float4*>(dst + j) = v;
        }
      } else {
        for (int j = threadIdx.x; j < count; j += 256) dst[j] = f.accumulate ? dst[j] + s_t[j] : s_t[j];
      }
      __syncthreads();
    }
    return;
  }
  const size_t total = static_cast<size_t>(f.m_real) * f.n_real * 49;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    // dw index i = ((co * cin + ci) * 7 + kh) * 7 + kw ; m_real = cout, n_real = cin
    const int kw = static_cast<int>(i % 7), kh = static_cast<int>((i / 7) % 7);
    const size_t cc = i / 49;
    const int ci = static_cast<int>(cc % f.n_real), co = static_cast<int>(cc / f.n_real);
    float v;
    if (f.mode == 1)
      v = ws[(static_cast<size_t>(kh) * f.m_tot + co) * f.n_tot + kw * f.cin_st + ci];
    else
      v = ws[(static_cast<size_t>(kh) * f.m_tot + ci) * f.n_tot + (6 - kw) * 8 + co];
    dw[i] = f.accumulate ? dw[i] + v : v;
  }
}

struct WgPlan {
  WgParams p;
  WgFinalize f;
  int a_is_dy;                 // which tensor feeds the A (P) operand
  uint64_t a_dims[5], a_strides[4], b_dims[5], b_strides[4];
  uint32_t a_box[5], b_box[5];
  size_t a_base_off, b_base_off;  // byte offsets of the view origins inside x / dy
  size_t ws_bytes;
};

static int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// 4-D {C, W, H, B} view of a (B, H + 2 pad, W + 2 pad, C) tensor; origin = logical pixel (oy, ox) (may be negative
// only if the border exists)
static void plain_view(uint64_t* dims, uint64_t* strides, uint32_t* box, size_t* base_off, int C, int H, int W, int B,
                       int pad, int oy, int ox, int view_h, int view_w, int kw_cols, int kh_rows) {
  // pad | JPDSE_PAD_SHARED: pitch W + pad, image stride (H + pad) rows (include/jpdse_b200.h)
  const bool shared = (pad & JPDSE_PAD_SHARED) != 0;
  pad &= JPDSE_PAD_SHARED - 1;
  const uint64_t Wp = shared ? W + pad : W + 2 * pad, Hp = shared ? H + pad : H + 2 * pad;
  dims[0] = C; dims[1] = view_w; dims[2] = view_h; dims[3] = B;
  strides[0] = static_cast<uint64_t>(C) * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
  box[0] = 64; box[1] = kw_cols; box[2] = kh_rows; box[3] = 1;
  *base_off = (static_cast<size_t>(pad + oy) * Wp + (pad + ox)) * C * 2;
}

// 5-D {2C, W/2, 2, H/2, B} (column pair, row parity) view used for stride-2 addressing of a (B,H+2pad,W+2pad,C) tensor
static void s2_view(uint64_t* dims, uint64_t* strides, uint32_t* box, size_t* base_off, int C, int H, int W, int B, int pad,
                    int kw_cols, int kh_rows) {
  const uint64_t Wp = W + 2 * pad, Hp = H + 2 * pad;
  dims[0] = 2ull * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
  strides[0] = 2ull * C * 2; strides[1] = Wp * C * 2; strides[2] = 2 * Wp * C * 2; strides[3] = Hp * Wp * C * 2;
  box[0] = 64; box[1] = kw_cols; box[2] = 1; box[3] = kh_rows; box[4] = 1;
  *base_off = (static_cast<size_t>(pad) * Wp + pad) * C * 2;
}

static void s2_tap(int* off, int C, int kh, int kw) {
  off[0] += (kw == 1) ? 0 : C;
  off[1] = (kw == 0) ? -1 : 0;
  off[2] = (kh == 1) ? 0 : 1;
  off[3] = (kh == 0) ? -1 : 0;
}

static int wgrad_plan(const jpdse_conv_desc* d, int dy_pad, WgPlan* w) {
  ConvGeom g;
  int rc = conv_geom(d, &g);
  if (rc != JPDSE_OK) return rc;
  if (dy_pad < 0) return fail(JPDSE_ERR_INVALID, "conv_wgrad: dy_pad < 0");
  if ((dy_pad & JPDSE_PAD_SHARED) && d->kind != JPDSE_CONV3X3_PAD1)
    return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad: the shared-border gradient layout is only built for JPDSE_CONV3X3_PAD1");
  memset(w, 0, sizeof(*w));
  WgParams& p = w->p;
  WgFinalize& f = w->f;
  const int B = d->batch, H = d->in_h, W = d->in_w, Cin = d->cin, Cout = d->cout;
  p.batch = B;
  int grid_h, grid_w;  // the K (pixel) grid
  const bool is7 = d->kind == JPDSE_CONV7X7_PAD3;
  const bool head = is7 && d->epilogue == JPDSE_EPI_BIAS_TANH_NCHW;
  if (is7 && head) {
    if (Cin != 64 || Cout > 8 || dy_pad != 6) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad head: needs cin == 64, cout <= 8, dy_pad == 6");
    grid_h = H;
    grid_w = W + 6;
  } else if (is7) {
    if (Cin != 40 || Cout != 64 || dy_pad != 0) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad stem: needs cin == 40 (stored), cout == 64, dy_pad == 0");
    grid_h = H + 6;
    grid_w = W;
  } else if (d->kind == JPDSE_CONV3X3_PAD1 || d->kind == JPDSE_CONV1X1 || d->kind == JPDSE_CONVT3X3_S2) {
    grid_h = H;
    grid_w = W;
  } else if (d->kind == JPDSE_CONV3X3_S2) {
    grid_h = H / 2;
    grid_w = W / 2;
  } else if (d->kind == JPDSE_CONV4X4_S2 || d->kind == JPDSE_CONV4X4_S1) {
    grid_h = g.out_h;
    grid_w = g.out_w;
  } else {
    return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad: kind %d has no weight gradient", d->kind);
  }
  p.kw_cols = grid_w >= 64 ? 64 : next_pow2(grid_w);
  p.kh_rows = 64 / p.kw_cols;
  if (d->kind == JPDSE_CONV4X4_S2 || d->kind == JPDSE_CONV4X4_S1 || d->kind == JPDSE_CONV1X1) {
    // the PatchGAN's grids (513, 257, 131, 67 ... wide) are never a whole number of 64-pixel row pieces: take the
    // 64-pixel box shape (64x1 ... 8x8) that covers the grid with the least overhang (67 wide: 48 % -> 22 % dead pixels)
    const char* e = getenv("JPDSE_WGRAD_BOX");  // "0": always 64 x 1
    long long best = -1;
    for (int kw = 64; kw >= 8 && !(e != nullptr && e[0] == '0'); kw >>= 1) {
      const int kh = 64 / kw;
      const long long area = static_cast<long long>((grid_w + kw - 1) / kw) * kw * ((grid_h + kh - 1) / kh) * kh;
      if (best < 0 || area < best) {
        best = area;
        p.kw_cols = kw;
        p.kh_rows = kh;
      }
    }
  }
  p.nbw = (grid_w + p.kw_cols - 1) / p.kw_cols;
  p.nbh = (grid_h + p.kh_rows - 1) / p.kh_rows;

  if (is7 && !head) {
    // stem: A = dy rows shifted by -kh (two filter rows per item), B = window view of the padded input
    w->a_is_dy = 1;
    plain_view(w->a_dims, w->a_strides, w->a_box, &w->a_base_off, Cout, H, W, B, 0, 0, 0, H, W, p.kw_cols, p.kh_rows);
    p.a = WgView{4, 1, 2, 3};
    const uint64_t Hp = H + 6, Wp = W + 6;
    w->b_dims[0] = 320; w->b_dims[1] = W; w->b_dims[2] = Hp; w->b_dims[3] = B;
    w->b_strides[0] = static_cast<uint64_t>(Cin) * 2; w->b_strides[1] = Wp * Cin * 2; w->b_strides[2] = Hp * Wp * Cin * 2;
    w->b_box[0] = 64; w->b_box[1] = p.kw_cols; w->b_box[2] = p.kh_rows; w->b_box[3] = 1;
    w->b_base_off = 0;
    p.b = WgView{4, 1, 2, 3};
    // four filter rows per item (tall: 7 boxes feed 8 MMAs) unless switched off: two rows per item, 5 boxes per 4 MMAs
    const char* e_tall = getenv("JPDSE_WGRAD_TALL");
    p.tall = (e_tall && e_tall[0] == '0') ? 0 : 1;
    const int rows_per_item = p.tall ? 4 : 2;
    p.n_groups = p.tall ? 2 : 4;
    p.m_tiles = 1;
    p.n_tiles = 2;
    p.nb = 3;
    p.b_tile_stride = 192;
    p.n_out_stride = 192;
    for (int gi = 0; gi < p.n_groups; ++gi) {
      for (int j = 0; j < rows_per_item; ++j) {
        const int kh = rows_per_item * gi + j;
        p.a_off[gi][j][2] = -(kh < 7 ? kh : 0);
        p.a_tap[gi][j] = kh < 7 ? kh : -1;
        p.a_ch[gi][j] = 0;
      }
      for (int j = 0; j < 3; ++j) {
        p.b_off[gi][j][0] = 64 * j;
        p.b_tap[gi][j] = 0;
        p.b_ch[gi][j] = 64 * j;
      }
    }
    p.m_tot = 64;
    p.n_tot = 320;
    f = WgFinalize{1, 7, p.m_tot, p.n_tot, Cout, d->cin_real, Cin, 0};
  } else if (head) {
    // head: A = padded input rows shifted by +kh, B = window over the 8-channel gradient pixels (zero border 6)
    w->a_is_dy = 0;
    plain_view(w->a_dims, w->a_strides, w->a_box, &w->a_base_off, Cin, H, W, B, 3, -3, -3, H + 6, W + 6, p.kw_cols, p.kh_rows);
    p.a = WgView{4, 1, 2, 3};
    const uint64_t Hs = H + 12, Ws = W + 12;
    w->b_dims[0] = 64; w->b_dims[1] = W + 6; w->b_dims[2] = H; w->b_dims[3] = B;
    w->b_strides[0] = 16; w->b_strides[1] = Ws * 16; w->b_strides[2] = Hs * Ws * 16;
    w->b_box[0] = 64; w->b_box[1] = p.kw_cols; w->b_box[2] = p.kh_rows; w->b_box[3] = 1;
    w->b_base_off = static_cast<size_t>(6) * Ws * 16;
    p.b = WgView{4, 1, 2, 3};
    p.n_groups = 4;
    p.m_tiles = 1;
    p.n_tiles = 1;
    p.nb = 1;
    for (int gi = 0; gi < 4; ++gi) {
      for (int j = 0; j < 2; ++j) {
        const int kh = 2 * gi + j;
        p.a_off[gi][j][2] = kh < 7 ? kh : 0;
        p.a_tap[gi][j] = kh < 7 ? kh : -1;
        p.a_ch[gi][j] = 0;
      }
      p.b_tap[gi][0] = 0;
      p.b_ch[gi][0] = 0;
    }
    p.m_tot = 64;
    p.n_tot = 64;
    f = WgFinalize{2, 7, p.m_tot, p.n_tot, Cout, d->cin_real, 0, 0};
  } else {
    // generic 3x3 / 1x1: P channels on M, (tap, Q channel block) on N
    const bool convt = d->kind == JPDSE_CONVT3X3_S2;
    const bool k4 = d->kind == JPDSE_CONV4X4_S2 || d->kind == JPDSE_CONV4X4_S1;
    const int taps = d->kind == JPDSE_CONV1X1 ? 1 : (k4 ? 16 : 9);
    const int kdim = k4 ? 4 : 3;
    // the PatchGAN's 1-channel output conv: its gradient tensor is stored with 64 channels (zeros beyond cout)
    const int pc = convt ? Cin : (k4 ? (Cout + 63) / 64 * 64 : Cout);   // P channels (as stored)
    const int qc = convt ? Cout : Cin;   // Q channels
    if (pc % 64 || qc % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad: channels must be multiples of 64 (got %d, %d)", pc, qc);
    w->a_is_dy = convt ? 0 : 1;
    // ---- P view
    if (convt) {
      plain_view(w->a_dims, w->a_strides, w->a_box, &w->a_base_off, Cin, H, W, B, d->in_pad, 0, 0, H, W, p.kw_cols, p.kh_rows);
    } else {
      plain_view(w->a_dims, w->a_strides, w->a_box, &w->a_base_off, pc, g.out_h, g.out_w, B, dy_pad, 0, 0, g.out_h, g.out_w,
                 p.kw_cols, p.kh_rows);
    }
    p.a = WgView{4, 1, 2, 3};
    // ---- Q view
    const bool strided = convt || d->kind == JPDSE_CONV3X3_S2;
    if (k4 && d->in_pad != 2) return fail(JPDSE_ERR_INVALID, "conv_wgrad 4x4: in_pad must be 2");
    if (d->kind == JPDSE_CONV3X3_PAD1) {
      plain_view(w->b_dims, w->b_strides, w->b_box, &w->b_base_off, Cin, H, W, B, 1, -1, -1, H + 2, W + 2, p.kw_cols, p.kh_rows);
      p.b = WgView{4, 1, 2, 3};
    } else if (d->kind == JPDSE_CONV1X1) {
      plain_view(w->b_dims, w->b_strides, w->b_box, &w->b_base_off, Cin, H, W, B, d->in_pad, 0, 0, H, W, p.kw_cols, p.kh_rows);
      p.b = WgView{4, 1, 2, 3};
    } else if (d->kind == JPDSE_CONV3X3_S2) {
      if ((H & 1) || (W & 1)) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad s2: odd input size");
      s2_view(w->b_dims, w->b_strides, w->b_box, &w->b_base_off, Cin, H, W, B, d->in_pad, p.kw_cols, p.kh_rows);
      p.b = WgView{5, 1, 3, 4};
    } else if (d->kind == JPDSE_CONV4X4_S1) {
      plain_view(w->b_dims, w->b_strides, w->b_box, &w->b_base_off, Cin, H, W, B, 2, -2, -2, H + 4, W + 4, p.kw_cols, p.kh_rows);
      p.b = WgView{4, 1, 2, 3};
    } else if (d->kind == JPDSE_CONV4X4_S2) {
      // the stored zero-bordered (B,H+4,W+4,C) tensor as {2C, (W+4)/2, 2, (H+4)/2, B}, origin at the stored corner
      const uint64_t Wst = W + 4, Hst = H + 4;
      w->b_dims[0] = 2ull * Cin; w->b_dims[1] = Wst / 2; w->b_dims[2] = 2; w->b_dims[3] = Hst / 2; w->b_dims[4] = B;
      w->b_strides[0] = 2ull * Cin * 2; w->b_strides[1] = Wst * Cin * 2; w->b_strides[2] = 2 * Wst * Cin * 2;
      w->b_strides[3] = Hst * Wst * Cin * 2;
      w->b_box[0] = 64; w->b_box[1] = p.kw_cols; w->b_box[2] = 1; w->b_box[3] = p.kh_rows; w->b_box[4] = 1;
      w->b_base_off = 0;
      p.b = WgView{5, 1, 3, 4};
    } else {
      s2_view(w->b_dims, w->b_strides, w->b_box, &w->b_base_off, Cout, 2 * H, 2 * W, B, dy_pad, p.kw_cols, p.kh_rows);
      p.b = WgView{5, 1, 3, 4};
    }
    // ---- N side: list of (tap, 64-channel block) columns, chopped into groups of nb boxes
    const int qblocks = qc / 64;
    int per_tap_tiles = 1;
    if (qblocks >= 4) {
      if (qblocks % 4) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad: %d channels not a multiple of 256", qc);
      p.nb = 4;
      per_tap_tiles = qblocks / 4;
    } else if (qblocks == 3) {
      p.nb = 3;
    } else {
      p.nb = (taps == 1) ? qblocks : (qblocks == 2 ? 4 : 3);  // 2 taps x 2 blocks, or 3 taps x 1 block
    }
    const int taps_per_group = qblocks >= 3 ? 1 : p.nb / qblocks;
    p.n_groups = (taps + taps_per_group - 1) / taps_per_group;
    if (p.n_groups > kWgMaxGroups) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad: too many tap groups");
    p.n_tiles = per_tap_tiles;
    p.b_tile_stride = 256;
    p.n_out_stride = 256;
    for (int gi = 0; gi < p.n_groups; ++gi)
      for (int j = 0; j < p.nb; ++j) {
        const int t = gi * taps_per_group + (qblocks >= 3 ? 0 : j / qblocks);
        const int blk = qblocks >= 3 ? j : j % qblocks;
        int* off = p.b_off[gi][j];
        off[0] = 64 * blk;
        if (t < taps) {
          const int kh = taps == 1 ? 0 : t / kdim, kw = taps == 1 ? 0 : t % kdim;
          if (strided) {
            s2_tap(off, qc, kh, kw);
          } else if (d->kind == JPDSE_CONV4X4_S2) {
            off[0] += (kw & 1) * qc;
            off[1] = kw >> 1;
            off[2] = kh & 1;
            off[3] = kh >> 1;
          } else if (taps == 9 || taps == 16) {
            off[1] = kw;
            off[2] = kh;
          }
          p.b_tap[gi][j] = static_cast<short>(t);
        } else {
          p.b_tap[gi][j] = -1;
        }
        p.b_ch[gi][j] = static_cast<short>(64 * blk);
      }
    // ---- M side
    p.m_tiles = (pc + 127) / 128;
    p.a_tile_stride = 128;
    p.m_out_stride = 128;
    {
      // tall items (M = 256) wherever P has the channels for them and the K loop is long enough to amortise the
      // un-overlapped drain of both accumulators (>= 24 k-blocks per item even after a split across the machine)
      const char* e = getenv("JPDSE_WGRAD_TALL");  // "0": the M = 128 form everywhere (read per call: tests toggle it)
      const int kb_all = B * p.nbh * p.nbw;
      const int tall_items = p.n_groups * (pc / 256) * p.n_tiles;  // 0 when P has fewer than 256 channels
      const int tall_splits = (tall_items == 0 || tall_items >= num_sms()) ? 1 : num_sms() / tall_items;
      if (!(e && e[0] == '0') && pc >= 256 && pc % 256 == 0 && kb_all / tall_splits >= 24) {
        p.tall = 1;
        p.m_tiles = pc / 256;
        p.a_tile_stride = 256;
        p.m_out_stride = 256;
      }
    }
    for (int gi = 0; gi < p.n_groups; ++gi)
      for (int j = 0; j < 4; ++j) {
        const bool valid = j < 2 ? (pc >= 128 || j == 0) : p.tall != 0;
        p.a_off[gi][j][0] = valid ? 64 * j : 0;
        p.a_tap[gi][j] = valid ? 0 : -1;
        p.a_ch[gi][j] = static_cast<short>(64 * j);
      }
    p.m_tot = pc;
    p.n_tot = qc;
    const int n_real = convt ? Cout : d->cin_real;
    f = WgFinalize{0, taps, p.m_tot, p.n_tot, k4 ? Cout : pc, n_real, 0, 0};
  }
  // split K so that the launch fills the machine
  const int per_split = p.n_groups * p.m_tiles * p.n_tiles;
  const int kb_total = B * p.nbh * p.nbw;
  // Never more items than SMs (rounded DOWN): 8 items per split and 19 splits were 152 items -- a second wave of four
  // items as long as the first (the 7x7 stem: 762 us, the longest kernel of the backward pass); 18 splits are one wave.
  int splits = 1;
  if (per_split < num_sms()) {
    splits = num_sms() / per_split;
    if (splits > kb_total / 8) splits = kb_total / 8;  // keep >= 8 k-blocks per item
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  // with one split every (tap, m, n) of the workspace that finalize reads is written by exactly one item -- true for
  // the generic forms; the 7x7 window forms keep the reduction path (their unused boxes leave holes finalize skips,
  // but rows of two filter taps share an item only through distinct workspace planes, so they qualify too)
  p.direct = splits == 1 ? 1 : 0;
  w->ws_bytes = static_cast<size_t>(f.taps) * p.m_tot * p.n_tot * sizeof(float);
  return JPDSE_OK;
}

}  // namespace jpdse

using namespace jpdse;

extern "C" size_t jpdse_conv_wgrad_workspace_bytes(const jpdse_conv_desc* d, int dy_pad) {
  WgPlan w;
  if (wgrad_plan(d, dy_pad, &w) != JPDSE_OK) return 0;
  return w.ws_bytes;
}

extern "C" int jpdse_conv_wgrad(const jpdse_conv_desc* d, const void* x, const void* dy, int dy_pad, float* dw,
                                int accumulate, void* workspace, size_t workspace_bytes, void* stream_v) {
  WgPlan w;
  int rc = wgrad_plan(d, dy_pad, &w);
  if (rc != JPDSE_OK) return rc;
  if (x == nullptr || dy == nullptr || dw == nullptr || workspace == nullptr) return fail(JPDSE_ERR_INVALID, "conv_wgrad: NULL pointer");
  if (workspace_bytes < w.ws_bytes) return fail(JPDSE_ERR_INVALID, "conv_wgrad: workspace too small (%zu < %zu)", workspace_bytes, w.ws_bytes);
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(workspace)) & 15)
    return fail(JPDSE_ERR_INVALID, "conv_wgrad: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const uint8_t* a_ptr = static_cast<const uint8_t*>(w.a_is_dy ? dy : x) + w.a_base_off;
  const uint8_t* b_ptr = static_cast<const uint8_t*>(w.a_is_dy ? x : dy) + w.b_base_off;
  CUtensorMap ta, tb;
  rc = make_tmap_bf16(&ta, a_ptr, w.p.a.rank, w.a_dims, w.a_strides, w.a_box);
  if (rc != JPDSE_OK) return rc;
  rc = make_tmap_bf16(&tb, b_ptr, w.p.b.rank, w.b_dims, w.b_strides, w.b_box);
  if (rc != JPDSE_OK) return rc;
  cudaError_t e = cudaSuccess;
  if (!w.p.direct) {
    e = cudaMemsetAsync(workspace, 0, w.ws_bytes, stream);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "conv_wgrad: memset failed: %s", cudaGetErrorString(e));
  }
  w.p.ws = static_cast<float*>(workspace);
  static DeviceOnce configured;  // the attribute is per device: set it on each device this process uses
  if (configured.first_use()) {
    e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "cudaFuncSetAttribute(wgrad smem=%d): %s", kWgSmemBytes, cudaGetErrorString(e));
    configured.done();
  }
  const int items = w.p.n_groups * w.p.m_tiles * w.p.n_tiles * w.p.splits;
  int grid = num_sms();
  if (grid > items) grid = items;
  wgrad_kernel<<<grid, kWgThreads, kWgSmemBytes, stream>>>(ta, tb, w.p);
  rc = check_launch("wgrad_kernel");
  if (rc != JPDSE_OK) return rc;
  w.f.accumulate = accumulate ? 1 : 0;
  if (w.f.mode == 0 && w.f.taps > 16) return fail(JPDSE_ERR_UNSUPPORTED, "conv_wgrad: finalize stages at most 16 taps");
  if (w.f.mode == 0 && (reinterpret_cast<uintptr_t>(dw) & 15)) return fail(JPDSE_ERR_INVALID, "conv_wgrad: dw must be 16-byte aligned");
  const size_t total = static_cast<size_t>(w.f.m_real) * w.f.n_real * (w.f.mode == 0 ? 1 : 49);
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  wgrad_finalize_kernel<<<blocks, 256, 0, stream>>>(static_cast<const float*>(workspace), dw, w.f);
  return check_launch("wgrad_finalize_kernel");
}
