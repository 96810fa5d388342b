// CTA-pair (tcgen05 cta_group::2) implicit GEMM for the residual-block convolutions -- 3x3, stride 1, input already
// reflect-padded by 1, Cout a multiple of 256 (reference ctu/models/pix2pixHD_networks/networks.py:283-299), 45-50 % of a
// forward. Also serves the stride-1 "valid" form of any other wide 3x3 layer.
//
// The single-CTA kernel (conv_igemm.cu) keeps the tensor pipe 84 % busy: a 128 x 256 tile needs a 16 KiB A tile and a
// 32 KiB B tile per 64-element k-block, four 48 KiB stages are all that fits, and the L2 -> SM fabric runs at 65 % of
// its peak just for this kernel. Here two CTAs on the two SMs of a TPC share ONE 256 x 256 tile:
//   * each CTA loads its own 128 pixels of A and HALF of the weight tile (128 of the 256 output channels): 32 KiB per
//     stage instead of 48 -> six stages, and a third less L2 -> SM traffic per FLOP;
//   * the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16): the hardware reads A
//     rows 0..127 / 128..255 and B rows 0..127 / 128..255 from the two CTAs' shared memory at the same offsets and leaves
//     each CTA's 128 x 256 half of the accumulator in that CTA's own TMEM;
//   * both CTAs' TMA loads signal the LEADER's "full" barrier (cta_group::2 form); the leader's commits are multicast to
//     both CTAs' "empty" / "accumulator full" barriers; the follower's epilogue warps release the accumulator on the
//     leader's barrier through the cluster.
// Epilogue as in conv_igemm.cu: bf16 NHWC stores + InstanceNorm statistics.
//
// Warp roles per CTA (320 threads): warps 0..3 / 4..7 epilogue warpgroups (one per TMEM accumulator buffer), warp 8 TMA
// producer, warp 9 TMEM allocation + (leader only) MMA issue.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_shared.cuh"
#include "ptx.cuh"

namespace jpdse {

constexpr int kPrThreads = 320;
constexpr int kPrProducerWarp = 8;
constexpr int kPrMmaWarp = 9;
constexpr int kPrABytes = 128 * 64 * 2;      // this CTA's 128 pixels x 64 k
constexpr int kPrBBytes = 128 * 64 * 2;      // this CTA's 128 output channels x 64 k
constexpr int kPrStageBytes = kPrABytes + kPrBBytes;
constexpr int kPrStages = 6;
constexpr int kPrRedBytes = 8 * 32 * 17 * 4;
constexpr int kPrSmemBytes = 1024 + kPrStages * kPrStageBytes + kPrRedBytes + 256;

struct PairParams {
  int batch, tiles_h, tiles_w, n_tiles;  // m tiles = batch * tiles_h * tiles_w (even), n tiles of 256
  int tile_h, tile_w;
  int chunks;                            // cin / 64
  int out_h, out_w, ldc;
  int want_stats;
  void* out;
  double* stats;
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// idesc for cta_group::2: M = 256 spans both CTAs
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(256 >> 4) << 24);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPrThreads, 1)
pair_conv3x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const __grid_constant__ PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t* s_red = reinterpret_cast<uint32_t*>(smem + kPrStages * kPrStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPrStages * kPrStageBytes + kPrRedBytes);
  uint64_t* full_bar = bars;                    // [kPrStages]  leader's copy is the one in use: TMA (both CTAs) -> MMA
  uint64_t* empty_bar = bars + kPrStages;       // [kPrStages]  per CTA: MMA commit (multicast) -> this CTA's producer
  uint64_t* tfull_bar = bars + 2 * kPrStages;   // [2]          per CTA: MMA commit (multicast) -> this CTA's epilogue
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]          leader's copy: 8 epilogue warps (both CTAs) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == kPrProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kPrStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 8);  // 4 epilogue warps of each CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == kPrMmaWarp) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits and both TMEM allocations visible to the peer before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL (JPDSE_PDL=1): everything above touched no global memory and ran while the previous kernel drained; from here on
  // the kernel reads what that kernel wrote. The InstanceNorm kernel behind this conv may place its CTAs now.
  grid_dep_launch_dependents();
  grid_dep_wait();

  const int kblocks = 9 * p.chunks;
  const int m_pairs = (p.batch * p.tiles_h * p.tiles_w) >> 1;
  const int total = m_pairs * p.n_tiles;

  if (warp == kPrProducerWarp) {
    // ================================================================== TMA producer (both CTAs)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int q = cluster_id; q < total; q += n_clusters) {
        const int nt = q % p.n_tiles;
        int mt = 2 * (q / p.n_tiles) + static_cast<int>(rank);
        const int tw = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int th = mt % p.tiles_h;
        const int b = mt / p.tiles_h;
        const int w0 = tw * p.tile_w, h0 = th * p.tile_h;
        int kb = 0;
        for (int t = 0; t < 9; ++t) {
          const int c1 = w0 + t % 3, c2 = h0 + t / 3;
          for (int ch = 0; ch < p.chunks; ++ch, ++kb) {
            mbar_wait_cluster(&empty_bar[stage], phase ^ 1);  // released by the leader's multicast commit
            uint8_t* sa = smem + stage * kPrStageBytes;
            const uint32_t lead_full = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kPrStageBytes);  // both CTAs' bytes land here
            tma_load_4d_cg2(&tm_a, lead_full, sa, ch * 64, c1, c2, b);
            tma_load_2d_cg2(&tm_b, lead_full, sa + kPrABytes, kb * 64, nt * 256 + static_cast<int>(rank) * 128);
            if (++stage == kPrStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kPrMmaWarp) {
    // ================================================================== MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_m256(256);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int q = cluster_id; q < total; q += n_clusters) {
        mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait_cluster(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kPrStageBytes);
          const uint64_t adesc = umma_smem_desc_sw128(sa);
          const uint64_t bdesc = umma_smem_desc_sw128(sa + kPrABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16<2>(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                         (kb | k) != 0 ? 1u : 0u);
          umma_commit_cg2_mc(&empty_bar[stage], 0b11);  // both producers may refill once these MMAs retire
          if (kb == kblocks - 1) umma_commit_cg2_mc(&tfull_bar[acc], 0b11);
          if (++stage == kPrStages) {
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
    // ================================================================== epilogue (both CTAs, own 128 pixels)
    const int group = warp >> 2;
    const int quarter = warp & 3;
    const int m = quarter * 32 + lane;
    uint32_t* s_t = s_red + warp * (32 * 17);
    const int acc = group;
    uint32_t acc_phase = 0;
    float run_s1a[8], run_s1b[8], run_s2a[8], run_s2b[8];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) run_s1a[ch] = run_s1b[ch] = run_s2a[ch] = run_s2b[ch] = 0.f;
    int cur_b = -1, cur_n0 = 0;
    auto flush = [&]() {
      if (cur_b >= 0 && lane < 16) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          double* st = p.stats + (static_cast<size_t>(cur_b) * p.ldc + cur_n0 + ch * 32 + lane * 2) * 2;
          atomicAdd(st + 0, static_cast<double>(run_s1a[ch]));
          atomicAdd(st + 1, static_cast<double>(run_s2a[ch]));
          atomicAdd(st + 2, static_cast<double>(run_s1b[ch]));
          atomicAdd(st + 3, static_cast<double>(run_s2b[ch]));
          run_s1a[ch] = run_s1b[ch] = run_s2a[ch] = run_s2b[ch] = 0.f;
        }
      }
    };
    int it = 0;
    for (int q = cluster_id; q < total; q += n_clusters, ++it) {
      if ((it & 1) != group) continue;
      const int nt = q % p.n_tiles;
      int mt = 2 * (q / p.n_tiles) + static_cast<int>(rank);
      const int tw = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int b = mt / p.tiles_h;
      const int r = m / p.tile_w;
      const int c = m - r * p.tile_w;
      const int oh = th * p.tile_h + r, ow = tw * p.tile_w + c;
      const int n0 = nt * 256;
      mbar_wait_cluster(&tfull_bar[acc], acc_phase);  // (a parked wait measured 3-4 % slower here: wake-up latency)
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 256);
      if (p.want_stats && (b != cur_b || n0 != cur_n0)) {
        flush();
        cur_b = b;
        cur_n0 = n0;
      }
      __nv_bfloat16* orow =
          reinterpret_cast<__nv_bfloat16*>(p.out) + ((static_cast<size_t>(b) * p.out_h + oh) * p.out_w + ow) * p.ldc + n0;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + ch * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
        uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        if (!p.want_stats) continue;
        // column sums through a 32x16-word shared transpose (see conv_igemm.cu)
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);  // the leader's barrier, from either CTA
      acc_phase ^= 1;
    }
    if (p.want_stats) flush();
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the peer still reads its shared memory or signals it
  if (warp == kPrMmaWarp) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

bool pair_conv_applicable(const jpdse_conv_desc* d) {
  const char* e = getenv("JPDSE_PAIR_CONV");  // "0" falls back to the single-CTA kernel (read per call: tests toggle it)
  if (e != nullptr && e[0] == '0') return false;
  if (d->kind != JPDSE_CONV3X3_PAD1 || d->in_pad != 1) return false;
  if (d->epilogue != JPDSE_EPI_RAW_STATS && d->epilogue != JPDSE_EPI_RAW) return false;
  if (d->cin % 64 || d->cout % 256 || d->cin_real != d->cin) return false;
  const int w = d->in_w < 128 ? d->in_w : 128;
  if (w <= 0 || (128 % w) || (d->in_w % w) || (d->in_h % (128 / w))) return false;
  const long long m_tiles = static_cast<long long>(d->batch) * (d->in_h / (128 / w)) * (d->in_w / w);
  if (m_tiles & 1) return false;
  // worth it only when the pairs of SMs get enough tiles and the k loop hides the epilogue
  static int min_tiles = -1;
  if (min_tiles < 0) {
    const char* t = getenv("JPDSE_PAIR_MIN_TILES");
    min_tiles = t ? atoi(t) : 96;  // measured: even with batch 3 (96 pair tiles on 74 SM pairs) against the single-CTA kernel, ahead from 4
  }
  return m_tiles / 2 * (d->cout / 256) >= min_tiles && 9 * (d->cin / 64) >= 72;
}

int pair_conv_forward(const jpdse_conv_desc* d, const void* x, const void* w_packed, void* y, double* stats,
                      cudaStream_t stream) {
  PairParams p;
  memset(&p, 0, sizeof(p));
  p.batch = d->batch;
  p.tile_w = d->in_w < 128 ? d->in_w : 128;
  p.tile_h = 128 / p.tile_w;
  p.tiles_w = d->in_w / p.tile_w;
  p.tiles_h = d->in_h / p.tile_h;
  p.n_tiles = d->cout / 256;
  p.chunks = d->cin / 64;
  p.out_h = d->in_h;
  p.out_w = d->in_w;
  p.ldc = d->cout;
  p.want_stats = d->epilogue == JPDSE_EPI_RAW_STATS ? 1 : 0;
  p.out = y;
  p.stats = stats;
  const uint64_t C = d->cin, B = d->batch, Hp = d->in_h + 2, Wp = d->in_w + 2;
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {C, Wp, Hp, B};
    uint64_t strides[3] = {C * 2, Wp * C * 2, Hp * Wp * C * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
    int rc = make_tmap_bf16(&ta, x, 4, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
  }
  {
    const uint64_t ktot = 9 * C;
    uint64_t dims[2] = {ktot, static_cast<uint64_t>(d->cout)};
    uint64_t strides[1] = {ktot * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16(&tb, w_packed, 2, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
  }
  static DeviceOnce configured;  // the attribute is per device: set it on each device this process uses
  if (configured.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(pair_conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPrSmemBytes);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "cudaFuncSetAttribute(pair smem=%d): %s", kPrSmemBytes, cudaGetErrorString(e));
    configured.done();
  }
  int grid = num_sms() & ~1;
  const long long total = static_cast<long long>(p.batch) * p.tiles_h * p.tiles_w / 2 * p.n_tiles;
  if (grid > 2 * total) grid = static_cast<int>(2 * total);
  if (pdl_enabled()) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kPrThreads);
    cfg.dynamicSmemBytes = kPrSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pair_conv3x3_kernel, ta, tb, p);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "pair_conv3x3_kernel (PDL launch): %s", cudaGetErrorString(e));
    return check_launch("pair_conv3x3_kernel");
  }
  pair_conv3x3_kernel<<<grid, kPrThreads, kPrSmemBytes, stream>>>(ta, tb, p);
  return check_launch("pair_conv3x3_kernel");
}

}  // namespace jpdse
