// Fused-phase ConvTranspose2d(3x3, stride 2, pad 1, output_padding 1) for the narrow, full-resolution upsampling
// layers of GlobalGenerator (reference ctu/models/pix2pixHD_networks/networks.py:241-245) and, with the roles swapped,
// the data gradient of the stride-2 downsampling convs (:213-216).
//
// Why a second kernel: the generic implicit GEMM (conv_igemm.cu) runs the four output phases as four launches and
// fetches a 128-pixel A tile once per filter tap -- 18 boxes of 16 KiB per input tile at Cin = 128. Those tiles are
// unique to their CTA, and unique bytes are what the L2 -> SM fabric caps (~6300 B/clk chip-wide, ~43 B/clk/SM:
// measured with the MMAs and the epilogue switched off, every low-K conv sat at ~480 cycles per 16 KiB A tile). With
// N = Cout <= 128 there is far too little MMA work per A byte, so these layers ran at 180-310 TFLOP/s.
//
// Here one tile = one input row piece of 128 pixels (+1 halo pixel), ALL FOUR output phases:
//   * the A box {64 channels, 129 pixels} of input row i (and of row i+1) is loaded once per filter ROW; the kw = 0
//     tap reads the same shared-memory tile shifted by one 128-byte row (a SWIZZLE_128B K-major operand may start at
//     any 128-byte row, tools/umma_shift_test.cu) -- 6 boxes per 64 input channels instead of 18;
//   * per stage two MMAs: shift 0 x [W(kh,1) | W(kh,2)] (N = 128: both column phases at once) and
//     shift 1 x W(kh,0) (N = 64) accumulate into a 4 x 64-column TMEM accumulator [ph0pw0 | ph0pw1 | ph1pw0 | ph1pw1];
//   * the epilogue drains the four 128 x 64 sub-tiles through swizzled shared-memory staging and TMA stores into the
//     (column-parity, row-parity) view of the output, reducing the InstanceNorm statistics on the way.
//
// Warp roles: warps 0..3 / 4..7 epilogue warpgroups (one per TMEM accumulator buffer), warp 8 TMA producer, warp 9 TMEM
// owner + MMA issuer.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_shared.cuh"
#include "ptx.cuh"

namespace jpdse {

constexpr int kCtThreads = 320;
constexpr int kCtProducerWarp = 8;
constexpr int kCtMmaWarp = 9;
constexpr int kCtABytes = 17 * 1024;        // 129 rows x 128 B, padded to the 1024 B swizzle period
constexpr int kCtABox = 129 * 128;          // bytes the TMA actually writes
constexpr int kCtBBytes = 3 * 64 * 128;     // [kw = 1 | kw = 2 | kw = 0] x 64 output channels x 64 k
constexpr int kCtStageBytes = kCtABytes + kCtBBytes;
constexpr int kCtStages = 3;
constexpr int kCtOutBytes = 4 * 128 * 128;  // 2 groups x 2 staging buffers of 128 rows x 128 B
constexpr int kCtRedBytes = 8 * 32 * 17 * 4;
constexpr int kCtSmemBytes = 1024 + kCtStages * kCtStageBytes + kCtOutBytes + kCtRedBytes + 256;

struct ConvtParams {
  int batch, height, width;  // input (= GEMM) extent
  int tiles_w;               // width / 128
  int n_blocks;              // cout / 64
  int chunks;                // cin / 64
  int cout;
  int want_stats;
  double* stats;
  int dbg_flags;  // developer experiments (JPDSE_DEBUG_FLAGS, as in conv_igemm.cu): 1 skip statistics | 2 skip the TMA
                  // stores | 8 hand the accumulator straight back | 16 skip the MMAs | 32 skip the weight loads
};

__global__ void __launch_bounds__(kCtThreads, 1)
convt_fused_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ ConvtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_out = smem + kCtStages * kCtStageBytes;
  uint32_t* s_red = reinterpret_cast<uint32_t*>(s_out + kCtOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + kCtOutBytes + kCtRedBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kCtStages;
  uint64_t* tfull_bar = bars + 2 * kCtStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kCtProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    tma_prefetch_desc(&tm_c);
    for (int i = 0; i < kCtStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == kCtMmaWarp) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int per_img = p.height * p.tiles_w;
  const int total_tiles = p.batch * p.n_blocks * per_img;
  const int kstages = 3 * p.chunks;  // pipeline steps per tile: (chunk, filter row)

  if (warp == kCtProducerWarp) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int r = tile % per_img;
        const int q = tile / per_img;
        const int tw = r % p.tiles_w;
        const int i = r / p.tiles_w;
        const int nblk = q % p.n_blocks;
        const int b = q / p.n_blocks;
        for (int c = 0; c < p.chunks; ++c) {
#pragma unroll
          for (int khi = 0; khi < 3; ++khi) {  // filter rows in the order kh = 1, 2 (input row i), 0 (input row i+1)
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * kCtStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], (p.dbg_flags & 32) ? kCtABox : kCtABox + kCtBBytes);
            tma_load_4d(&tm_a, &full_bar[stage], sa, c * 64, tw * 128, i + (khi == 2 ? 1 : 0), b);
            if (!(p.dbg_flags & 32)) tma_load_2d(&tm_b, &full_bar[stage], sa + kCtABytes, c * 64, (nblk * 3 + khi) * 192);
            if (++stage == kCtStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kCtMmaWarp) {
    if (elect_one()) {
      constexpr uint32_t idesc_wide = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_narrow = umma_idesc_bf16(128, 64);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tile = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int s = 0; s < kstages; ++s) {
          const int khi = s % 3;
          const uint32_t d_ph = d_tile + (khi == 0 ? 0u : 128u);  // kh = 1 feeds row phase 0, kh = 2 / 0 row phase 1
          const bool first = s < 2;                               // first touch of this row phase's columns
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kCtStageBytes);
          const uint64_t a0 = umma_smem_desc_sw128(sa);                    // pixels j .. j+127
          const uint64_t a1 = umma_smem_desc_sw128(sa + 128);              // pixels j+1 .. j+128
          const uint64_t b01 = umma_smem_desc_sw128(sa + kCtABytes);       // [W(kh,1) | W(kh,2)]
          const uint64_t b2 = umma_smem_desc_sw128(sa + kCtABytes + 128 * 128);  // W(kh,0)
          if (!(p.dbg_flags & 16)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16<1>(d_ph, a0 + static_cast<uint64_t>(k * 2), b01 + static_cast<uint64_t>(k * 2), idesc_wide,
                           (first && k == 0) ? 0u : 1u);
              umma_bf16<1>(d_ph + 64, a1 + static_cast<uint64_t>(k * 2), b2 + static_cast<uint64_t>(k * 2), idesc_narrow, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (s == kstages - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kCtStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue
    const int group = warp >> 2;
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    uint32_t* s_t = s_red + warp * (32 * 17);
    const int acc = group;
    uint32_t acc_phase = 0;
    float run_s1a[2], run_s1b[2], run_s2a[2], run_s2b[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) run_s1a[ch] = run_s1b[ch] = run_s2a[ch] = run_s2b[ch] = 0.f;
    int cur_b = -1, cur_n0 = 0;
    auto flush = [&]() {
      if (cur_b >= 0 && lane < 16) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          double* st = p.stats + (static_cast<size_t>(cur_b) * p.cout + cur_n0 + ch * 32 + lane * 2) * 2;
          atomicAdd(st + 0, static_cast<double>(run_s1a[ch]));
          atomicAdd(st + 1, static_cast<double>(run_s2a[ch]));
          atomicAdd(st + 2, static_cast<double>(run_s1b[ch]));
          atomicAdd(st + 3, static_cast<double>(run_s2b[ch]));
          run_s1a[ch] = run_s1b[ch] = run_s2a[ch] = run_s2b[ch] = 0.f;
        }
      }
    };
    auto chunk_stats = [&](uint32_t (&pk)[16], int ch) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) s_t[lane * 17 + j] = pk[j];
      __syncwarp();
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
      const uint32_t* col = s_t + (lane >> 4) * (16 * 17) + (lane & 15);
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const uint32_t w2 = col[t * 17];
        const float lo = __uint_as_float(w2 << 16), hi = __uint_as_float(w2 & 0xffff0000u);
        s1a += lo;
        s1b += hi;
        s2a = fmaf(lo, lo, s2a);
        s2b = fmaf(hi, hi, s2b);
      }
      s1a += __shfl_xor_sync(0xffffffffu, s1a, 16);
      s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
      s2a += __shfl_xor_sync(0xffffffffu, s2a, 16);
      s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
      run_s1a[ch] += s1a;
      run_s1b[ch] += s1b;
      run_s2a[ch] += s2a;
      run_s2b[ch] += s2b;
    };
    const bool issuer = quarter == 0 && lane == 0;
    int out_buf = 0;
    int tile_i = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      if ((tile_i & 1) != group) continue;
      const int r = tile % per_img;
      const int q = tile / per_img;
      const int tw = r % p.tiles_w;
      const int i = r / p.tiles_w;
      const int nblk = q % p.n_blocks;
      const int b = q / p.n_blocks;
      const int n0 = nblk * 64;
      mbar_wait_parked(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (p.want_stats && (b != cur_b || n0 != cur_n0)) {
        flush();
        cur_b = b;
        cur_n0 = n0;
      }
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256);
      if (p.dbg_flags & 8) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc_phase ^= 1;
        continue;
      }
#pragma unroll
      for (int sub = 0; sub < 4; ++sub) {  // sub-tile = (row phase, column phase), 64 columns each
        uint8_t* buf = s_out + (group * 2 + (out_buf & 1)) * (128 * 128);
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1 + group, 128);
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + sub * 64 + c2 * 32, v);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            pk[j] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* rowp = reinterpret_cast<uint4*>(buf + m * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            rowp[(c2 * 4 + j) ^ (m & 7)] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          if (p.want_stats && !(p.dbg_flags & 1)) chunk_stats(pk, c2);
        }
        if (sub == 3) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + group, 128);
        if (issuer && !(p.dbg_flags & 2)) {
          // output view {2*Cout (column phase major), W, 2 (row phase), H, B}
          tma_store_5d(&tm_c, buf, (sub & 1) * p.cout + n0, tw * 128, sub >> 1, i, b);
          tma_store_commit();
        }
        ++out_buf;
      }
      acc_phase ^= 1;
    }
    if (p.want_stats) flush();
    if (issuer) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kCtMmaWarp) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// weights (Cin, Cout, 3, 3) fp32 -> rows [nblk][kh in (1,2,0)][kw in (1,2,0)][64 channels of the block] x K = Cin, bf16
__global__ void convt_fused_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin, int cout) {
  const size_t total = static_cast<size_t>(9) * cout * cin;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % cin);
    const int row = static_cast<int>(idx / cin);
    const int nblk = row / 576, r2 = row % 576;
    const int khi = r2 / 192, r3 = r2 % 192;
    const int kwi = r3 / 64, co = nblk * 64 + r3 % 64;
    const int kh = khi == 0 ? 1 : (khi == 1 ? 2 : 0);
    const int kw = kwi == 0 ? 1 : (kwi == 1 ? 2 : 0);
    out[idx] = __float2bfloat16_rn(w[((static_cast<size_t>(k) * cout + co) * 3 + kh) * 3 + kw]);
  }
}

bool convt_fused_applicable(const jpdse_conv_desc* d) {
  static int disabled = -1;
  if (disabled < 0) {
    const char* e = getenv("JPDSE_NO_CONVT_FUSED");
    disabled = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return !disabled && d->kind == JPDSE_CONVT3X3_S2 && (d->epilogue == JPDSE_EPI_RAW_STATS || d->epilogue == JPDSE_EPI_RAW) &&
         d->cin % 64 == 0 && d->cout % 64 == 0 && d->cout <= 128 && d->in_w % 128 == 0 && d->cin_real == d->cin;
}

int convt_fused_pack(const jpdse_conv_desc* d, const float* w, void* w_packed, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(9) * d->cout * d->cin;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  convt_fused_pack_kernel<<<blocks, 256, 0, stream>>>(w, static_cast<__nv_bfloat16*>(w_packed), d->cin, d->cout);
  return check_launch("convt_fused_pack_kernel");
}

int convt_fused_forward(const jpdse_conv_desc* d, const void* x, const void* w_packed, void* y, double* stats,
                        cudaStream_t stream) {
  ConvtParams p;
  memset(&p, 0, sizeof(p));
  p.batch = d->batch;
  p.height = d->in_h;
  p.width = d->in_w;
  p.tiles_w = d->in_w / 128;
  p.n_blocks = d->cout / 64;
  p.chunks = d->cin / 64;
  p.cout = d->cout;
  p.want_stats = d->epilogue == JPDSE_EPI_RAW_STATS ? 1 : 0;
  p.stats = stats;
  {
    static int flags = -1;
    if (flags < 0) {
      const char* e = getenv("JPDSE_DEBUG_FLAGS");
      flags = e ? atoi(e) : 0;
    }
    p.dbg_flags = flags;
  }
  const uint64_t C = d->cin, Co = d->cout, H = d->in_h, W = d->in_w, B = d->batch;
  const uint64_t Hp = H + 2 * static_cast<uint64_t>(d->in_pad), Wp = W + 2 * static_cast<uint64_t>(d->in_pad);
  const uint8_t* xin = static_cast<const uint8_t*>(x) + (static_cast<uint64_t>(d->in_pad) * Wp + d->in_pad) * C * 2;
  CUtensorMap ta, tb, tc;
  {
    // the +1 halo pixel / row past the edge is TMA zero fill (a stored border is skipped over, never read)
    uint64_t dims[4] = {C, W, H, B};
    uint64_t strides[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    uint32_t box[4] = {64, 129, 1, 1};
    int rc = make_tmap_bf16(&ta, xin, 4, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
  }
  {
    uint64_t dims[2] = {C, 9 * Co};
    uint64_t strides[1] = {C * 2};
    uint32_t box[2] = {64, 192};
    int rc = make_tmap_bf16(&tb, w_packed, 2, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
  }
  {
    uint64_t dims[5] = {2 * Co, W, 2, H, B};
    uint64_t strides[4] = {2 * Co * 2, 2 * W * Co * 2, 2 * 2 * W * Co * 2, 2 * H * 2 * W * Co * 2};
    uint32_t box[5] = {64, 128, 1, 1, 1};
    int rc = make_tmap_bf16(&tc, y, 5, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
  }
  static DeviceOnce configured;  // the attribute is per device: set it on each device this process uses
  if (configured.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(convt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemBytes);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "cudaFuncSetAttribute(convt smem=%d): %s", kCtSmemBytes, cudaGetErrorString(e));
    configured.done();
  }
  const long long total = static_cast<long long>(p.batch) * p.n_blocks * p.height * p.tiles_w;
  int grid = num_sms();
  if (grid > total) grid = static_cast<int>(total);
  convt_fused_kernel<<<grid, kCtThreads, kCtSmemBytes, stream>>>(ta, tb, tc, p);
  return check_launch("convt_fused_kernel");
}

}  // namespace jpdse
