// Implicit-GEMM convolutions for the JPD-SE generator on sm_100a.
//
//   D[pixel, cout] = sum_{tap, cin} A[pixel @ tap, cin] * W[cout, tap, cin]
//
// One persistent, warp-specialised kernel covers every conv of GlobalGenerator
// (reference: ctu/models/pix2pixHD_networks/networks.py:210,215,244,246,283-299):
//   * A tiles (128 output pixels x 64 input channels, bf16) are fetched by TMA straight out of the
//     NHWC activation tensor: the tensor map is a 4-D/5-D *view* of the activations chosen per conv
//     kind so that "the 128 pixels under filter tap t" is one box (no im2col buffer, zero padding
//     comes from TMA out-of-bounds fill, reflect padding was materialised by the producer kernel).
//   * B tiles (BN output channels x 64, bf16) come from a pre-packed K-major weight matrix.
//   * tcgen05.mma (M=128, N=BN, K=16) accumulates in TMEM (fp32); two accumulator buffers let the
//     epilogue of tile i overlap the main loop of tile i+1.
//   * The epilogue reads TMEM with tcgen05.ld and reduces the InstanceNorm statistics (sum, sum of squares per
//     (image, channel)) through a 32x16-word shared-memory transpose, or applies bias+tanh / sign for the two terminal
//     convs. bf16 NHWC tiles of the N <= 128 instantiations (and of N = 256 where a tile has <= 48 k-blocks) are written
//     into swizzled shared-memory staging and leave through ONE TMA store per 64-channel half: per-lane global stores
//     (16 B at a pixel stride) cost the epilogue warps 2-3k cycles per 32-column chunk and made the low-K convs
//     epilogue-bound. The res-block convs (144 k-blocks a tile) keep the fourth pipeline stage and per-lane stores.
//   * The same kernel is the data-gradient conv of the backward pass: JPDSE_CONV3X3_FULL / CONV7X7_FULL run over flat
//     positions of the zero-bordered gradient (any H, W), and stride-2 <-> ConvTranspose swap roles on one weight tensor.
//
// Warp roles (320 threads): warps 0..3 and 4..7 = two epilogue warpgroups, one per TMEM accumulator buffer (TMEM
// lane quarter = warp_idx % 4); warp 8 = TMA producer, warp 9 = TMEM owner + MMA issuer.
// What bounds it: the res-block convs the tensor pipe (81-85 % active); every conv whose A tiles are unique to a CTA the
// L2 -> SM fabric (~43 B/clk/SM of unique bytes), see conv_convt.cu for the layers where that was worth a new kernel.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_shared.cuh"
#include "ptx.cuh"

namespace jpdse {

constexpr int kTileM = 128;         // output pixels per tile (UMMA M)
constexpr int kBlockK = 64;         // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;          // K of one tcgen05.mma for 16-bit inputs
constexpr int kABytes = kTileM * kBlockK * 2;
constexpr int kMaxTaps = 16;        // 4x4 PatchGAN convs
constexpr int kThreads = 320;       // two epilogue warpgroups (one per TMEM accumulator) + 2 control warps
// The control warps get the HIGHEST warp ids: the sub-partition arbiter favours high ids, and the single TMA /
// MMA threads are the critical path of the low-K convs.
// kNumProducers > 1 makes that many warps fill alternate pipeline stages. Measured: a second producer changes nothing
// (0.496 ms either way with MMAs and epilogue off) -- the ~480-cycle floor per 16 KiB A box is the L2 -> SM fabric's
// unique-byte cap, not the issuing thread -- so one producer it is; the code path stays for the experiment.
constexpr int kProducerWarp = 8;
constexpr int kNumProducers = 1;
constexpr int kMmaWarp = kProducerWarp + kNumProducers;
static_assert(kThreads == 32 * (kMmaWarp + 1), "warp roles and block size out of step");

// Row mode (igemm_kernel<64, 1> / <64, 0>; the VGG19's 64 -> 64 conv at full resolution and the data gradients of its
// first block). The generic K loop fetches one 128-pixel A box AND one weight box per filter tap: 9 x (16 + 8) KB = 216 KB
// per 128 x 64 output tile against 1152 tensor-pipe cycles -- 4900 cycles on the L2 -> SM fabric, 0.46 PFLOP/s. Here
//   * the nine 64 x 64 weight tiles (72 KB) are loaded ONCE per CTA and stay in shared memory;
//   * a k-block is a FILTER ROW: one box of 128 + 2 pixels, and the three kw taps are three MMAs whose A descriptors
//     start 0 / 1 / 2 pixel rows (128 bytes) into it (a SWIZZLE_128B operand may start at any 128-byte row with
//     base_offset 0: tools/umma_shift_test.cu) -- for the flat data-gradient kinds positions are contiguous, so the same
//     shift works there;
// i.e. 3 x 16.25 KB = 49 KB per tile. The ring holds four 17 KB stages behind the weights; the output staging and the
// barriers keep their places (72 KB + 4 x 17 KB <= 6 x 24 KB). full_bar[kRowWBar] (a stage index the ring does not use)
// is the weights' barrier.
constexpr int kRowStages = 4;
constexpr int kRowWBytes = 9 * 64 * kBlockK * 2;         // 73,728
constexpr int kRowABox = (kTileM + 2) * kBlockK * 2;     // 16,640: what one A box brings
constexpr int kRowAStage = 17 * 1024;                    // 1024-aligned stage pitch
constexpr int kRowWBar = 4;
static_assert(kRowWBytes + kRowStages * kRowAStage <= 6 * (kABytes + 64 * kBlockK * 2), "row mode must fit the BN = 64 stage area");
static_assert(kRowABox <= kRowAStage && kRowWBytes % 1024 == 0, "row-mode stage layout");

struct IgemmParams {
  // tile grid
  int batch, tiles_h, tiles_w, n_tiles;
  int tile_h, tile_w;  // tile_h * tile_w == 128
  int gemm_h, gemm_w;  // pixel grid of the GEMM; with `partial` the last tile row / column hangs over its edge
  int partial;         // 1: the grid is not a whole number of tiles (TMA zero-fills the loads and clips the stores;
                       //    the epilogue masks the overhanging pixels out of the statistics and per-lane stores)
  // K loop
  int ntaps, chunks_per_tap;
  // A tensor-map addressing
  int a_rank, dim_w, dim_h, dim_b;
  int tap_off[kMaxTaps][5];
  int b_k_offset;  // first k element of this launch inside the packed weight matrix (ConvT phases)
  int flat_pitch;  // > 0: "flat" mode -- M runs over flat positions y*pitch+x of a padded image (any H, W);
                   // rows with y >= out_h or x >= out_w are computed but neither stored nor counted
  // output addressing
  int out_h, out_w, os_h, os_w, op_h, op_w;
  int ldc;      // channels of the output tensor
  int n_valid;  // valid output channels (<= n_tiles * BN)
  int epilogue;
  int tma_store;   // 1: bf16 NHWC output leaves through shared memory + TMA store (tm_c), BN in {64, 128}
  int c_rank;      // 4: {c, w, h, b}; 5: {2c (col parity major), w, row parity, h, b} (ConvTranspose phases)
  int out_pad;     // bf16 NHWC epilogues: border of the OUTPUT tensor that is skipped over (zero, owned by the caller)
  int check_out;   // 1: rows whose output coordinate falls outside out_h x out_w are neither stored nor counted
                   //    (data gradient of the 4x4 stride-2 conv onto an odd-sized input)
  float slope;     // JPDSE_EPI_BIAS_ACT: LeakyReLU negative slope (0 = ReLU)
  int rowmode;     // 1 (BN = 64, 64 input channels, 3x3 stride 1): row-stationary K loop -- see kRow* below
  void* out;
  double* stats;
  const float* bias;
  long long* dbg;  // optional per-CTA role wait counters (tools only), else nullptr
  // developer experiments (env JPDSE_DEBUG_FLAGS, read once per process; 0 in production):
  //   1 skip statistics | 2 skip output stores | 4 spin instead of parked wait in the epilogue | 8 epilogue hands the
  //   accumulator straight back (no TMEM read) | 16 skip the MMAs | 32 skip the B (weight) TMA load | 64 plain
  //   mbarrier arrive instead of tcgen05.commit for the stage release (only with 16)
  int dbg_flags;
};

// STG = 1: bf16 NHWC output through shared-memory staging + TMA store (always for BN = 64 / 128; for BN = 256 it
// costs one pipeline stage, so only the low-K layers -- whose epilogue is exposed -- use it)
template <int BN, int STG>
struct IgemmCfg {
  static constexpr int kBBytes = BN * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // BN <= 128 gives up pipeline stages for the epilogue's output staging (2 groups x 2 buffers x 16 KiB)
  static constexpr bool kStaged = (BN == 64 || BN == 128 || (BN == 256 && STG == 1));
  static constexpr int kStageOutBytes = kStaged ? 4 * kABytes : 0;
  static constexpr int kStages = (BN >= 256) ? (kStaged ? 3 : 4) : (BN >= 128 ? 4 : (BN >= 64 ? 6 : 8));
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  static constexpr int kRedFloats = 8 * 32 * 17;  // per-epilogue-warp transpose scratch: 32 rows x 16 bf16x2 (+1 pad)
  // 1024 B alignment slack + stages + transpose scratch + barriers
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kStageOutBytes + kRedFloats * 4 + 256;
};

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// JPDSE_EPI_BIAS_ACT on one 32-column accumulator chunk: v = LeakyReLU_slope(v + bias) (every lane reads the same 32
// biases: broadcast loads out of L1)
__device__ __forceinline__ void bias_act32(uint32_t (&v)[32], const float* __restrict__ bias, float slope) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float x = __uint_as_float(v[j]) + __ldg(bias + j);
    v[j] = __float_as_uint(x > 0.f ? x : slope * x);
  }
}

template <int BN, int STG>
__global__ void __launch_bounds__(kThreads, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
             const __grid_constant__ CUtensorMap tm_c, const __grid_constant__ IgemmParams p) {
  using Cfg = IgemmCfg<BN, STG>;
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment as an OFFSET from the shared window (keeps the pointer in the shared address space -> LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_out = smem + Cfg::kStages * Cfg::kStageBytes;  // output staging (1024 B aligned: all stage sizes are)
  uint32_t* s_red = reinterpret_cast<uint32_t*>(s_out + Cfg::kStageOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + Cfg::kStageOutBytes + Cfg::kRedFloats * 4);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]        MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    if (p.tma_store) tma_prefetch_desc(&tm_c);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc<1>(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kblocks = p.ntaps * p.chunks_per_tap;
  const int m_tiles = p.batch * p.tiles_h * p.tiles_w;
  const int total_tiles = m_tiles * p.n_tiles;

  if (warp >= kProducerWarp && warp < kProducerWarp + kNumProducers) {
    // ================================================================== TMA producers (alternate k-blocks)
    if (elect_one()) {
      const uint32_t my = static_cast<uint32_t>(warp - kProducerWarp);
      uint32_t it = 0;  // global k-block counter: stage = it % kStages, phase = (it / kStages) & 1
      long long dbg_prod = 0;
      const long long dbg_t0 = p.dbg ? clock64() : 0;
      const bool row = BN == 64 && p.rowmode;
      const uint32_t n_stages = row ? kRowStages : Cfg::kStages;
      if (row && my == 0 && blockIdx.x < total_tiles) {
        // resident weights: nine {64 k, 64 n} boxes, tap-major as the packed matrix has them
        mbar_arrive_expect_tx(&full_bar[kRowWBar], kRowWBytes);
        for (int t = 0; t < 9; ++t) tma_load_2d(&tm_b, &full_bar[kRowWBar], smem + t * (64 * kBlockK * 2), t * kBlockK, 0);
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int tw = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int th = mt % p.tiles_h;
        const int b = mt / p.tiles_h;
        // Tile origin in tensor-map coordinates, straight-line (a dynamically indexed local array here costs the
        // producer thread hundreds of cycles of dependent local-memory traffic per tile -- fatal for the low-K convs).
        // rank 3 (flat): {k, position, image}; rank 4: {c, w, h, b}; rank 5: {c, w, parity, h, b}
        int b1 = tw * p.tile_w, b2, b3 = 0, b4 = 0;
        if (p.a_rank == 3) {
          b2 = b;
        } else if (p.a_rank == 4) {
          b2 = th * p.tile_h;
          b3 = b;
        } else {
          b2 = 0;
          b3 = th * p.tile_h;
          b4 = b;
        }
        int kb = 0;
        for (int t = 0; t < p.ntaps; ++t) {
          const int c1 = b1 + p.tap_off[t][1];
          const int c2 = b2 + p.tap_off[t][2];
          const int c3 = b3 + p.tap_off[t][3];
          const int c4 = b4 + p.tap_off[t][4];
          const int c0t = p.tap_off[t][0];
          for (int ch = 0; ch < p.chunks_per_tap; ++ch, ++kb, ++it) {
            if ((it % kNumProducers) != my) continue;
            const uint32_t stage = it % n_stages;
            const uint32_t phase = (it / n_stages) & 1u;
            const long long tw0 = p.dbg ? clock64() : 0;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (p.dbg) dbg_prod += clock64() - tw0;
            uint8_t* sa = row ? smem + kRowWBytes + stage * kRowAStage : smem + stage * Cfg::kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_arrive_expect_tx(&full_bar[stage], row ? kRowABox : ((p.dbg_flags & 32) ? kABytes : Cfg::kStageBytes));
            const int c0 = c0t + ch * kBlockK;
            if (p.a_rank == 3)
              tma_load_3d(&tm_a, &full_bar[stage], sa, c0, c1, c2);
            else if (p.a_rank == 4)
              tma_load_4d(&tm_a, &full_bar[stage], sa, c0, c1, c2, c3);
            else
              tma_load_5d(&tm_a, &full_bar[stage], sa, c0, c1, c2, c3, c4);
            if (!row && !(p.dbg_flags & 32)) tma_load_2d(&tm_b, &full_bar[stage], sb, p.b_k_offset + kb * kBlockK, nt * BN);
          }
        }
      }
      if (p.dbg && my == 0) {
        p.dbg[blockIdx.x * 16 + 0] = dbg_prod;
        p.dbg[blockIdx.x * 16 + 1] = clock64() - dbg_t0;
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTileM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long dbg_full = 0, dbg_tempty = 0;
      const long long dbg_t0 = p.dbg ? clock64() : 0;
      const bool row = BN == 64 && p.rowmode;
      const int n_stages = row ? kRowStages : Cfg::kStages;
      if (row && blockIdx.x < total_tiles) {
        mbar_wait(&full_bar[kRowWBar], 0);  // the resident weights have landed
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const long long tw0 = p.dbg ? clock64() : 0;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        if (p.dbg) dbg_tempty += clock64() - tw0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb) {
          const long long tw1 = p.dbg ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          if (p.dbg) dbg_full += clock64() - tw1;
          tc_fence_after();
          if (row) {
            // k-block = filter row kb: A = the 130-pixel box, shifted by kw pixel rows; B = resident tile (kb, kw)
            const uint32_t sa = smem_u32(smem + kRowWBytes + stage * kRowAStage);
            const uint32_t sw = smem_u32(smem) + static_cast<uint32_t>(kb * 3) * (64 * kBlockK * 2);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint64_t adesc = umma_smem_desc_sw128(sa + kw * 128);
              const uint64_t bdesc = umma_smem_desc_sw128(sw + kw * (64 * kBlockK * 2));
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                umma_bf16<1>(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                             (kb | kw | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
            if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);
            if (++stage == n_stages) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_smem_desc_sw128(sa);
          const uint64_t bdesc = umma_smem_desc_sw128(sa + kABytes);
          if (!(p.dbg_flags & 16)) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // advancing K inside the 128-byte swizzle row: +32 bytes = +2 in the (>>4) address field
              umma_bf16<1>(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                           (kb | k) != 0 ? 1u : 0u);
            }
          }
          if (p.dbg_flags & 64)
            mbar_arrive(&empty_bar[stage]);  // experiment (only valid with the MMAs switched off)
          else
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (kb == kblocks - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (p.dbg) {
        p.dbg[blockIdx.x * 16 + 2] = dbg_full;
        p.dbg[blockIdx.x * 16 + 3] = dbg_tempty;
        p.dbg[blockIdx.x * 16 + 4] = clock64() - dbg_t0;
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue (warps 0..3 / 4..7)
    const int group = warp >> 2;         // epilogue warpgroup == TMEM accumulator buffer it drains
    const int quarter = warp & 3;        // TMEM lanes [32*quarter, 32*quarter+32)
    const int m = quarter * 32 + lane;   // tile row = output pixel within the tile
    uint32_t* s_t = s_red + warp * (32 * 17);  // this warp's transpose scratch
    const int acc = group;
    uint32_t acc_phase = 0;
    // InstanceNorm statistics: running column sums in registers, flushed with one atomic per column when the
    // (image, channel tile) changes -- at full resolution a CTA stays on one image for ~100 tiles.
    constexpr int kChunks = BN >= 32 ? BN / 32 : 1;
    float run_s1a[kChunks], run_s1b[kChunks], run_s2a[kChunks], run_s2b[kChunks];
#pragma unroll
    for (int ch = 0; ch < kChunks; ++ch) run_s1a[ch] = run_s1b[ch] = run_s2a[ch] = run_s2b[ch] = 0.f;
    int cur_b = -1, cur_n0 = 0;
    auto flush = [&]() {
      if (cur_b >= 0 && lane < 16) {
#pragma unroll
        for (int ch = 0; ch < kChunks; ++ch) {
          double* st = p.stats + (static_cast<size_t>(cur_b) * p.ldc + cur_n0 + ch * 32 + lane * 2) * 2;
          atomicAdd(st + 0, static_cast<double>(run_s1a[ch]));
          atomicAdd(st + 1, static_cast<double>(run_s2a[ch]));
          atomicAdd(st + 2, static_cast<double>(run_s1b[ch]));
          atomicAdd(st + 3, static_cast<double>(run_s2b[ch]));
          run_s1a[ch] = run_s1b[ch] = run_s2a[ch] = run_s2b[ch] = 0.f;
        }
      }
    };
    // column sums of a 32-row x 32-column bf16x2 chunk (pk: this lane's row) through a 32x16-word shared transpose
    // (bank-conflict free): lane l then owns column pair (l & 15) over rows 16*(l >> 4) .. +15
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
    int tile_i = 0;
    int out_buf = 0;  // running index of this group's output staging buffer
    long long dbg_epi = 0, dbg_read = 0, dbg_tiles = 0;
    const long long dbg_t0 = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_i) {
      if ((tile_i & 1) != group) continue;
      const int nt = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int tw = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int b = mt / p.tiles_h;
      const int r = m / p.tile_w;
      const int c = m - r * p.tile_w;
      int oh = (th * p.tile_h + r) * p.os_h + p.op_h;
      int ow = (tw * p.tile_w + c) * p.os_w + p.op_w;
      bool valid = true;
      if (p.flat_pitch > 0) {  // flat mode: ow is the flat position
        oh = ow / p.flat_pitch;
        ow -= oh * p.flat_pitch;
        valid = oh < p.out_h && ow < p.out_w;
      } else if (p.partial) {
        valid = th * p.tile_h + r < p.gemm_h && tw * p.tile_w + c < p.gemm_w;
      }
      if (p.check_out) valid = valid && oh < p.out_h && ow < p.out_w;
      const int n0 = nt * BN;

      const long long tw2 = p.dbg ? clock64() : 0;
      if (p.dbg_flags & 4)
        mbar_wait(&tfull_bar[acc], acc_phase);
      else
        mbar_wait_parked(&tfull_bar[acc], acc_phase);
      const long long tr0 = p.dbg ? clock64() : 0;
      if (p.dbg) dbg_epi += tr0 - tw2;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);

      if (p.dbg_flags & 8) {  // experiment: hand the accumulator straight back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      } else if (Cfg::kStaged && p.tma_store) {
        // ---- bf16 NHWC tile -> swizzled shared-memory staging -> one TMA store per 64-channel half. Per-lane global
        // stores (16 B at a pixel stride) cost the epilogue warps ~2-3k cycles per 32-column chunk and made every
        // low-K conv epilogue-bound; here the warps only convert and write shared memory.
        if constexpr (Cfg::kStaged) {
          const bool want_stats = p.epilogue == JPDSE_EPI_RAW_STATS;
          if (want_stats && (b != cur_b || n0 != cur_n0)) {
            flush();
            cur_b = b;
            cur_n0 = n0;
          }
          const bool issuer = quarter == 0 && lane == 0;
#pragma unroll
          for (int half = 0; half < BN / 64; ++half) {
            uint8_t* buf = s_out + (group * 2 + (out_buf & 1)) * kABytes;
            if (issuer) tma_store_wait_read<1>();  // the store that last read this buffer (two halves ago) is done
            named_bar_sync(1 + group, 128);
#pragma unroll
            for (int c2 = 0; c2 < 2; ++c2) {
              const int ch = half * 2 + c2;
              uint32_t v[32];
              tmem_ld_32x32b_x32(taddr + ch * 32, v);
              tmem_ld_wait();
              if (p.epilogue == JPDSE_EPI_BIAS_ACT) bias_act32(v, p.bias + n0 + ch * 32, p.slope);
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
                pk[j] = *reinterpret_cast<uint32_t*>(&h);
              }
              // row m of the staged box = 128 B; SWIZZLE_128B: 16-byte chunk index ^= (row & 7)
              uint4* rowp = reinterpret_cast<uint4*>(buf + m * 128);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                rowp[(c2 * 4 + j) ^ (m & 7)] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
              if (p.partial && !valid) {  // the TMA store clips this pixel; keep it out of the sums too
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[j] = 0u;
              }
              if (want_stats && !(p.dbg_flags & 1)) chunk_stats(pk, ch);
            }
            if (half == BN / 64 - 1) {  // accumulator fully read: hand it back before the store is even issued
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            fence_proxy_async_smem();
            named_bar_sync(1 + group, 128);
            if (issuer && !(p.dbg_flags & 2)) {
              const int c0 = n0 + half * 64;
              if (p.c_rank == 4)
                tma_store_4d(&tm_c, buf, c0, tw * p.tile_w, th * p.tile_h, b);
              else
                tma_store_5d(&tm_c, buf, p.op_w * p.ldc + c0, tw * p.tile_w, p.op_h, th * p.tile_h, b);
              tma_store_commit();
            }
            ++out_buf;
          }
        }
      } else if (p.epilogue == JPDSE_EPI_RAW_STATS || p.epilogue == JPDSE_EPI_RAW || p.epilogue == JPDSE_EPI_BIAS_ACT) {
        const bool want_stats = p.epilogue == JPDSE_EPI_RAW_STATS;
        if (want_stats && (b != cur_b || n0 != cur_n0)) {
          flush();
          cur_b = b;
          cur_n0 = n0;
        }
        __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) +
                              ((static_cast<size_t>(b) * (p.out_h + 2 * p.out_pad) + oh + p.out_pad) * (p.out_w + 2 * p.out_pad) +
                               ow + p.out_pad) * p.ldc + n0;
        if constexpr (BN >= 32) {
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(taddr + ch * 32, v);
            tmem_ld_wait();
            if (p.epilogue == JPDSE_EPI_BIAS_ACT) bias_act32(v, p.bias + n0 + ch * 32, p.slope);
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
              pk[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            uint4* dst = reinterpret_cast<uint4*>(orow + ch * 32);
            if (valid && !(p.dbg_flags & 2)) {
#pragma unroll
              for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            }
            if (!want_stats || (p.dbg_flags & 1)) continue;
            if (!valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = 0u;
            }
            chunk_stats(pk, ch);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      } else {
        // terminal convs: few output channels, fp32 NCHW output
        if constexpr (BN <= 128) {
          float* obase = reinterpret_cast<float*>(p.out);
          const size_t plane = static_cast<size_t>(p.out_h) * p.out_w;
          const size_t pix = static_cast<size_t>(oh) * p.out_w + ow;
#pragma unroll 1
          for (int ch = 0; ch < BN / 16; ++ch) {
            if (n0 + ch * 16 >= p.n_valid) break;
            uint32_t v[16];
            tmem_ld_32x32b_x16(taddr + ch * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + ch * 16 + j;
              if (n < p.n_valid && valid) {
                float x = __uint_as_float(v[j]);
                float y;
                if (p.epilogue == JPDSE_EPI_BIAS_TANH_NCHW) {
                  y = tanhf(x + p.bias[n]);
                } else if (p.epilogue == JPDSE_EPI_BIAS_NCHW) {
                  y = x + p.bias[n];
                } else {
                  // torch.sign(torch.tanh(x)) == (x > 0) - (x < 0): tanh keeps the sign of every
                  // non-zero float (denormals included) and torch.sign maps NaN and -0 to 0
                  y = static_cast<float>((x > 0.f) - (x < 0.f));
                }
                obase[(static_cast<size_t>(b) * p.n_valid + n) * plane + pix] = y;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      acc_phase ^= 1;
      if (p.dbg) {
        dbg_read += clock64() - tr0;
        ++dbg_tiles;
      }
    }
    if (p.epilogue == JPDSE_EPI_RAW_STATS) flush();
    if (p.tma_store && quarter == 0 && lane == 0) tma_store_wait_read<0>();  // staging must outlive its readers
    if (p.dbg && lane == 0 && (warp == 0 || warp == 4)) {
      p.dbg[blockIdx.x * 16 + 5 + group] = dbg_epi;
      if (group == 0) p.dbg[blockIdx.x * 16 + 7] = clock64() - dbg_t0;
      p.dbg[blockIdx.x * 16 + 8 + group] = dbg_read;
      p.dbg[blockIdx.x * 16 + 10 + group] = dbg_tiles;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                   const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return fail(JPDSE_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(JPDSE_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
                static_cast<int>(r), rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                (unsigned long long)(rank > 4 ? gdim[4] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
                rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0);
  }
  return JPDSE_OK;
}

static int pick_bn(int cout) {
  if (cout >= 256) return 256;
  if (cout > 64) return 128;
  if (cout > 32) return 64;
  if (cout > 16) return 32;
  return 16;
}

// ConvTranspose2d(k=3,s=2,p=1,op=1): out[2i-1+kh, 2j-1+kw] += in[i,j] * W[kh,kw].
// Output phase 0 (even coordinate 2a) takes tap kh=1 at i=a; phase 1 (odd, 2a+1) takes kh=0 at i=a+1
// and kh=2 at i=a.
static int convt_taps(int phase, int kk[2], int dd[2]) {
  if (phase == 0) {
    kk[0] = 1;
    dd[0] = 0;
    return 1;
  }
  kk[0] = 0;
  dd[0] = 1;
  kk[1] = 2;
  dd[1] = 0;
  return 2;
}

int conv_geom(const jpdse_conv_desc* d, ConvGeom* g) {
  if (d == nullptr) return fail(JPDSE_ERR_INVALID, "conv desc is NULL");
  if (d->batch <= 0 || d->in_h <= 0 || d->in_w <= 0 || d->cin <= 0 || d->cout <= 0 || d->cin_real <= 0 ||
      d->cin_real > d->cin || d->in_pad < 0)
    return fail(JPDSE_ERR_INVALID, "conv desc: bad sizes");
  if (d->kind == JPDSE_CONV3X3_PAD1 && d->in_pad != 1) return fail(JPDSE_ERR_INVALID, "CONV3X3_PAD1 needs in_pad == 1");
  if (d->kind == JPDSE_CONV7X7_PAD3 && d->in_pad != 3) return fail(JPDSE_ERR_INVALID, "CONV7X7_PAD3 needs in_pad == 3");
  g->bn = pick_bn(d->cout);
  g->rows = ((d->cout + g->bn - 1) / g->bn) * g->bn;
  g->path = kPathIgemm;
  switch (d->kind) {
    case JPDSE_CONV3X3_FULL:
    case JPDSE_CONV3X3_FULL_SHARED:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv full: cin must be a multiple of 64 (got %d)", d->cin);
      if (d->in_pad != 2) return fail(JPDSE_ERR_INVALID, "CONV3X3_FULL needs in_pad == 2");
      g->out_h = g->gemm_h = d->in_h + 2;
      g->out_w = g->gemm_w = d->in_w + 2;
      g->cpt = d->cin / 64;
      g->ktot = 9 * d->cin;
      break;
    case JPDSE_CONV7X7_FULL:
      if ((d->cin * 2) % 16 || 7 * d->cin > 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv7x7 full: need cin*2 %% 16 == 0 and 7*cin <= 64");
      if (d->in_pad != 6) return fail(JPDSE_ERR_INVALID, "CONV7X7_FULL needs in_pad == 6");
      g->out_h = g->gemm_h = d->in_h + 6;
      g->out_w = g->gemm_w = d->in_w + 6;
      g->cpt = 1;
      g->ktot = 7 * 64;
      break;
    case JPDSE_CONV3X3_PAD1:
    case JPDSE_CONV1X1:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv: cin must be a multiple of 64 (got %d)", d->cin);
      g->out_h = g->gemm_h = d->in_h;
      g->out_w = g->gemm_w = d->in_w;
      g->cpt = d->cin / 64;
      g->ktot = (d->kind == JPDSE_CONV1X1 ? 1 : 9) * d->cin;
      if (pair_conv_applicable(d) && d->out_pad == 0 && d->epilogue != JPDSE_EPI_BIAS_ACT)
        g->path = kPathPair;  // same packed weights as the single-CTA kernel
      break;
    case JPDSE_CONV3X3_S2:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv s2: cin must be a multiple of 64 (got %d)", d->cin);
      if ((d->in_h & 1) || (d->in_w & 1)) return fail(JPDSE_ERR_UNSUPPORTED, "conv s2: odd input size");
      g->out_h = g->gemm_h = d->in_h / 2;
      g->out_w = g->gemm_w = d->in_w / 2;
      g->cpt = d->cin / 64;
      g->ktot = 9 * d->cin;
      break;
    case JPDSE_CONVT3X3_S2:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "convT: cin must be a multiple of 64 (got %d)", d->cin);
      g->out_h = 2 * d->in_h;
      g->out_w = 2 * d->in_w;
      g->gemm_h = d->in_h;
      g->gemm_w = d->in_w;
      g->cpt = d->cin / 64;
      g->ktot = 9 * d->cin;
      if (convt_fused_applicable(d)) g->path = kPathConvtFused;
      break;
    case JPDSE_CONV4X4_S2:
    case JPDSE_CONV4X4_S1:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv4x4: cin must be a multiple of 64 (got %d)", d->cin);
      if (d->in_pad != 2) return fail(JPDSE_ERR_INVALID, "CONV4X4 kinds need in_pad == 2 (the conv's zero padding)");
      g->out_h = g->gemm_h = d->kind == JPDSE_CONV4X4_S2 ? d->in_h / 2 + 1 : d->in_h + 1;
      g->out_w = g->gemm_w = d->kind == JPDSE_CONV4X4_S2 ? d->in_w / 2 + 1 : d->in_w + 1;
      g->cpt = d->kind == JPDSE_CONV4X4_S2 ? 2 * d->cin / 64 : d->cin / 64;  // stride 2: one tap = a column PAIR
      g->ktot = 16 * d->cin;
      break;
    case JPDSE_CONV4X4_S2_DGRAD:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv4x4 dgrad: cin must be a multiple of 64 (got %d)", d->cin);
      if (d->in_pad != 2) return fail(JPDSE_ERR_INVALID, "CONV4X4_S2_DGRAD needs in_pad == 2");
      if (d->out_h <= 0 || d->out_w <= 0 || d->out_h / 2 + 1 != d->in_h || d->out_w / 2 + 1 != d->in_w)
        return fail(JPDSE_ERR_INVALID, "CONV4X4_S2_DGRAD: out_h/out_w (%d x %d) are not a forward input size of a %d x %d output",
                    d->out_h, d->out_w, d->in_h, d->in_w);
      g->out_h = d->out_h;
      g->out_w = d->out_w;
      g->gemm_h = (d->out_h + 1) / 2;
      g->gemm_w = (d->out_w + 1) / 2;
      g->cpt = d->cin / 64;
      g->ktot = 16 * d->cin;  // four phase blocks of 4 taps
      break;
    case JPDSE_CONV4X4_S1_FULL:
      if (d->cin % 64) return fail(JPDSE_ERR_UNSUPPORTED, "conv4x4 full: cin must be a multiple of 64 (got %d)", d->cin);
      if (d->in_pad != 2) return fail(JPDSE_ERR_INVALID, "CONV4X4_S1_FULL needs in_pad == 2");
      if (d->in_h < 2 || d->in_w < 2) return fail(JPDSE_ERR_INVALID, "CONV4X4_S1_FULL: gradient smaller than 2x2");
      g->out_h = g->gemm_h = d->in_h - 1;
      g->out_w = g->gemm_w = d->in_w - 1;
      g->cpt = d->cin / 64;
      g->ktot = 16 * d->cin;
      break;
    case JPDSE_CONV3X3_PAD1_NARROW:
      if ((d->cin * 2) % 16 || 3 * d->cin > 64)
        return fail(JPDSE_ERR_UNSUPPORTED, "conv3x3 narrow: need cin*2 %% 16 == 0 and 3*cin <= 64 (got %d)", d->cin);
      if (d->in_pad != 1) return fail(JPDSE_ERR_INVALID, "CONV3X3_PAD1_NARROW needs in_pad == 1");
      g->out_h = g->gemm_h = d->in_h;
      g->out_w = g->gemm_w = d->in_w;
      g->cpt = 1;
      g->ktot = 3 * 64;
      break;
    case JPDSE_CONV7X7_PAD3:
      if ((d->cin * 2) % 16) return fail(JPDSE_ERR_UNSUPPORTED, "conv7x7: cin*2 bytes must be a multiple of 16");
      g->out_h = g->gemm_h = d->in_h;
      g->out_w = g->gemm_w = d->in_w;
      g->cpt = (7 * d->cin + 63) / 64;
      g->ktot = 7 * g->cpt * 64;
      // full-resolution stem / head: row-stationary kernels (conv_rowstat.cu) when the shape allows
      if (d->epilogue == JPDSE_EPI_BIAS_TANH_NCHW && d->cin == 64 && d->cout <= 4 && d->in_w >= 128) {
        g->path = kPathRowHead;
        g->rows = 7 * 32;  // [q = 6-kh][kw*cout+co padded to 32]
        g->ktot = 64;
      } else if (d->epilogue == JPDSE_EPI_RAW_STATS && d->cin == 40 && d->cout % 32 == 0 && d->cout <= 128 &&
                 d->in_w % 128 == 0) {
        g->path = kPathRowStem;
        g->rows = (d->cout / 32) * 7 * 5 * 32;  // [split][kb][q = 6-kh][32 channels]
        g->ktot = 64;
      }
      break;
    default:
      return fail(JPDSE_ERR_INVALID, "conv desc: unknown kind %d", d->kind);
  }
  // Small problems (batch 1 at the bottleneck): with N = 256 tiles fewer than ~3/4 of the SMs get a tile; halve N so
  // the tile count doubles (an N = 128 MMA costs about half an N = 256 one, so a lone tile also finishes sooner).
  // (A wave-quantisation-aware choice for the flat data-gradient kinds was tried and lost: N = 128 doubles the unique
  // A traffic and falls back to per-lane stores there.)
  if (g->bn == 256 && g->path == kPathIgemm && d->kind != JPDSE_CONV3X3_FULL && d->kind != JPDSE_CONV3X3_FULL_SHARED &&
      d->kind != JPDSE_CONV4X4_S1_FULL) {
    const long long m_tiles = (static_cast<long long>(d->batch) * g->gemm_h * g->gemm_w + 127) / 128;
    if (m_tiles * (g->rows / 256) * 4 < static_cast<long long>(num_sms()) * 3) g->bn = 128;
  }
  if (d->out_pad < 0) return fail(JPDSE_ERR_INVALID, "conv desc: out_pad < 0");
  if (d->epilogue == JPDSE_EPI_RAW_STATS || d->epilogue == JPDSE_EPI_RAW || d->epilogue == JPDSE_EPI_BIAS_ACT) {
    if (d->cout % g->bn || g->bn < 32)
      return fail(JPDSE_ERR_UNSUPPORTED, "bf16 NHWC epilogues need cout %% %d == 0 (got %d)", g->bn, d->cout);
    if (d->out_pad && g->path != kPathIgemm)
      return fail(JPDSE_ERR_UNSUPPORTED, "out_pad is only supported by the generic implicit-GEMM path");
  } else if (d->epilogue == JPDSE_EPI_BIAS_TANH_NCHW || d->epilogue == JPDSE_EPI_SIGN_NCHW || d->epilogue == JPDSE_EPI_BIAS_NCHW) {
    if (g->bn > 128) return fail(JPDSE_ERR_UNSUPPORTED, "NCHW epilogues support cout <= 128 (got %d)", d->cout);
  } else {
    return fail(JPDSE_ERR_INVALID, "conv desc: unknown epilogue %d", d->epilogue);
  }
  return JPDSE_OK;
}

// M tiling: 128 pixels = tile_h rows x tile_w columns of the GEMM pixel grid; tile_w is 128 or, for narrower grids,
// the next power of two. Grids that are not a whole number of tiles get overhanging edge tiles (IgemmParams::partial).
static void pick_tile(int gw, int* th, int* tw, int gh = 0) {
  int w = 128;
  if (gw < 128) {
    w = 1;
    while (w < gw) w <<= 1;
  } else if (gh > 0) {
    // grids that are not a whole number of 128-pixel row pieces (the PatchGAN's 513 / 257 / 129-wide maps: 128 + 128 + 1
    // pixels are THREE tiles a row): the 128-pixel tile shape (128x1 ... 16x8) that covers the grid with the least overhang;
    // ties keep the widest (every power-of-two width keeps 128x1). JPDSE_TILE_SHAPE=0: always 128x1.
    static int on = -1;
    if (on < 0) {
      const char* e = getenv("JPDSE_TILE_SHAPE");
      on = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    long long best = -1;
    for (int cw = 128; cw >= 16 && on; cw >>= 1) {
      const int ch = 128 / cw;
      const long long area = static_cast<long long>((gw + cw - 1) / cw) * cw * ((gh + ch - 1) / ch) * ch;
      if (best < 0 || area < best) {
        best = area;
        w = cw;
      }
    }
  }
  *tw = w;
  *th = 128 / w;
}

static long long* g_dbg = nullptr;  // role counters of the LAST igemm launch when enabled (tools/role_times.py)
static bool g_dbg_enabled = false;

template <int BN, int STG>
static int launch_igemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const IgemmParams& p,
                        cudaStream_t stream) {
  using Cfg = IgemmCfg<BN, STG>;
  static DeviceOnce configured;  // the attribute is per device: set it on each device this process uses
  if (configured.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(igemm_kernel<BN, STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    configured.done();
  }
  const int total = p.batch * p.tiles_h * p.tiles_w * p.n_tiles;
  int grid = num_sms();
  if (grid > total) grid = total;
  igemm_kernel<BN, STG><<<grid, kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tc, p);
  return check_launch("igemm_kernel");
}

static int launch_igemm_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                           const IgemmParams& p, cudaStream_t stream) {
  switch (bn) {
    case 256: return p.tma_store ? launch_igemm<256, 1>(ta, tb, tc, p, stream) : launch_igemm<256, 0>(ta, tb, tc, p, stream);
    case 128: return launch_igemm<128, 1>(ta, tb, tc, p, stream);
    case 64: return launch_igemm<64, 1>(ta, tb, tc, p, stream);
    case 32: return launch_igemm<32, 0>(ta, tb, tc, p, stream);
    case 16: return launch_igemm<16, 0>(ta, tb, tc, p, stream);
  }
  return fail(JPDSE_ERR_INVALID, "no igemm instantiation for BN=%d", bn);
}

// ------------------------------------------------------------------------------------------ weight packing
struct PackParams {
  int kind, cin, cin_real, cout, rows, ktot, cpt, path, cout_real;
};

__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, PackParams q) {
  const size_t total = static_cast<size_t>(q.rows) * q.ktot;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float val = 0.f;
    if (q.path == kPathRowHead) {
      // [q = 6 - kh][n = kw*cout + co (zero padded to 32)][c]
      const int c = static_cast<int>(i % 64);
      const int n = static_cast<int>((i / 64) % 32);
      const int kh = 6 - static_cast<int>(i / (64 * 32));
      const int kw = n / q.cout, co = n % q.cout;
      if (kw < 7 && c < q.cin_real) val = w[((static_cast<size_t>(co) * q.cin_real + c) * 7 + kh) * 7 + kw];
    } else if (q.path == kPathRowStem) {
      // [split][kb][q = 6 - kh][n (32)][k (64)], element e = kb*64+k of the 7*cin window under filter row kh
      const int k = static_cast<int>(i % 64);
      size_t r = i / 64;
      const int n = static_cast<int>(r % 32);
      r /= 32;
      const int kh = 6 - static_cast<int>(r % 7);
      r /= 7;
      const int kb = static_cast<int>(r % 5);
      const int split = static_cast<int>(r / 5);
      const int e = kb * 64 + k;
      const int kw = e / q.cin, c = e % q.cin;
      if (kw < 7 && c < q.cin_real)
        val = w[((static_cast<size_t>(split * 32 + n) * q.cin_real + c) * 7 + kh) * 7 + kw];
    } else if (q.kind == JPDSE_CONVT3X3_S2) {
      // layout: phase blocks [(ph,pw) = 00,01,10,11], each (rows x ntaps*cin) row-major
      size_t rem = i;
      int ph = 0, pw = 0, nt = 1;
      for (int pidx = 0; pidx < 4; ++pidx) {
        ph = pidx >> 1;
        pw = pidx & 1;
        nt = (ph ? 2 : 1) * (pw ? 2 : 1);
        const size_t blk = static_cast<size_t>(q.rows) * nt * q.cin;
        if (rem < blk) break;
        rem -= blk;
      }
      const int kp = nt * q.cin;
      const int n = static_cast<int>(rem / kp);
      const int k = static_cast<int>(rem % kp);
      const int t = k / q.cin, c = k % q.cin;
      const int ntw = pw ? 2 : 1;
      const int ti = t / ntw, tj = t % ntw;
      const int kh = ph ? (ti == 0 ? 0 : 2) : 1;
      const int kw = pw ? (tj == 0 ? 0 : 2) : 1;
      if (n < q.cout && c < q.cin_real) val = w[((static_cast<size_t>(c) * q.cout + n) * 3 + kh) * 3 + kw];
    } else {
      const int n = static_cast<int>(i / q.ktot);
      const int k = static_cast<int>(i % q.ktot);
      if (q.kind == JPDSE_CONV3X3_FULL) {
        // W'[n = ci_fwd][t' = kh'*3+kw'][c = co_fwd] = W_fwd[co_fwd][ci_fwd][2-kh'][2-kw']
        // (the forward weight is (cin_real, cout_real, 3, 3): cout_real < cout for the VGG's 3-channel input conv)
        const int t = k / q.cin, c = k % q.cin;
        const int kh = 2 - t / 3, kw = 2 - t % 3;
        if (n < q.cout_real && c < q.cin_real) val = w[((static_cast<size_t>(c) * q.cout_real + n) * 3 + kh) * 3 + kw];
      } else if (q.kind == JPDSE_CONV7X7_FULL) {
        // per filter row kh': window element e = kw'*cin + c ; W'[n][kh'][e] = W_fwd[c][n][6-kh'][6-kw']
        const int khp = k / 64, e = k % 64;
        const int kwp = e / q.cin, c = e % q.cin;
        if (n < q.cout && kwp < 7 && c < q.cin_real)
          val = w[((static_cast<size_t>(c) * q.cout + n) * 7 + (6 - khp)) * 7 + (6 - kwp)];
      } else if (q.kind == JPDSE_CONV7X7_PAD3 || q.kind == JPDSE_CONV3X3_PAD1_NARROW) {
        const int ks = q.kind == JPDSE_CONV7X7_PAD3 ? 7 : 3;
        const int kh = k / (q.cpt * 64);
        const int e = k % (q.cpt * 64);
        const int kw = e / q.cin, c = e % q.cin;
        if (n < q.cout && kw < ks && c < q.cin_real)
          val = w[((static_cast<size_t>(n) * q.cin_real + c) * ks + kh) * ks + kw];
      } else if (q.kind == JPDSE_CONV1X1) {
        if (n < q.cout && k < q.cin_real) val = w[static_cast<size_t>(n) * q.cin_real + k];
      } else if (q.kind == JPDSE_CONV4X4_S2 || q.kind == JPDSE_CONV4X4_S1) {
        // [n][t = kh*4+kw][c]; the stride-2 kernel walks it as (kh, column pair) taps of 2*cin contiguous elements
        const int t = k / q.cin, c = k % q.cin;
        if (n < q.cout && c < q.cin_real) val = w[(static_cast<size_t>(n) * q.cin_real + c) * 16 + t];
      } else if (q.kind == JPDSE_CONV4X4_S1_FULL) {
        // W'[n = ci_fwd][t' = kh'*4+kw'][c = co_fwd] = W_fwd[co_fwd][ci_fwd][3-kh'][3-kw'], W_fwd is (cin_real, cout_real, 4, 4)
        const int t = k / q.cin, c = k % q.cin;
        const int kh = 3 - t / 4, kw = 3 - t % 4;
        if (n < q.cout_real && c < q.cin_real) val = w[((static_cast<size_t>(c) * q.cout_real + n) * 4 + kh) * 4 + kw];
      } else if (q.kind == JPDSE_CONV4X4_S2_DGRAD) {
        // four phase blocks (r, s), each rows x (4 taps x cin): tap t = i*2+j -> W_fwd[co][ci][2i+r][2j+s]
        const size_t blk = static_cast<size_t>(q.rows) * 4 * q.cin;
        const int ph = static_cast<int>(i / blk);
        const size_t rem = i % blk;
        const int nn = static_cast<int>(rem / (4 * q.cin));
        const int kk = static_cast<int>(rem % (4 * q.cin));
        const int t = kk / q.cin, c = kk % q.cin;
        const int kh = 2 * (t >> 1) + (ph >> 1), kw = 2 * (t & 1) + (ph & 1);
        if (nn < q.cout_real && c < q.cin_real) val = w[((static_cast<size_t>(c) * q.cout_real + nn) * 4 + kh) * 4 + kw];
      } else {
        const int t = k / q.cin, c = k % q.cin;
        if (n < q.cout && c < q.cin_real) val = w[(static_cast<size_t>(n) * q.cin_real + c) * 9 + t];
      }
    }
    out[i] = __float2bfloat16_rn(val);
  }
}

}  // namespace jpdse

using namespace jpdse;

// Developer instrumentation (not part of the documented ABI): per-CTA role wait cycles of the last
// implicit-GEMM launch: [producer wait-empty, producer total, mma wait-full, mma wait-tmem-empty, mma total,
// epilogue0 wait-tmem-full, epilogue1 wait-tmem-full, epilogue total] x 148 CTAs.
extern "C" int jpdse_debug_role_counters(int enable, long long* host_out, int max_values) {
  if (g_dbg == nullptr) {
    if (cudaMalloc(&g_dbg, sizeof(long long) * 16 * 256) != cudaSuccess) return fail(JPDSE_ERR_CUDA, "debug buffer alloc failed");
    cudaMemset(g_dbg, 0, sizeof(long long) * 16 * 256);
  }
  if (host_out != nullptr && max_values > 0) {
    cudaDeviceSynchronize();
    cudaMemcpy(host_out, g_dbg, sizeof(long long) * (max_values < 16 * 256 ? max_values : 16 * 256), cudaMemcpyDeviceToHost);
  }
  g_dbg_enabled = enable != 0;
  return JPDSE_OK;
}

extern "C" size_t jpdse_conv_packed_weight_bytes(const jpdse_conv_desc* d) {
  ConvGeom g;
  if (conv_geom(d, &g) != JPDSE_OK) return 0;
  return static_cast<size_t>(g.rows) * g.ktot * 2;
}

extern "C" int jpdse_conv_launch_count(const jpdse_conv_desc* d) {
  ConvGeom g;
  if (conv_geom(d, &g) != JPDSE_OK) return 0;
  if (d->kind == JPDSE_CONV4X4_S2_DGRAD) return 4;
  return (d->kind == JPDSE_CONVT3X3_S2 && g.path == kPathIgemm) ? 4 : 1;  // generic ConvTranspose = one launch per phase
}

extern "C" double jpdse_conv_flops(const jpdse_conv_desc* d) {
  ConvGeom g;
  if (conv_geom(d, &g) != JPDSE_OK) return 0.0;
  int taps = d->kind == JPDSE_CONV7X7_PAD3 ? 49 : (d->kind == JPDSE_CONV1X1 ? 1 : 9);
  if (d->kind == JPDSE_CONV4X4_S2 || d->kind == JPDSE_CONV4X4_S1 || d->kind == JPDSE_CONV4X4_S1_FULL) taps = 16;
  const int co = (d->cout_real > 0 && d->cout_real <= d->cout) ? d->cout_real : d->cout;
  if (d->kind == JPDSE_CONV4X4_S2_DGRAD)  // 4 of the 16 taps reach each output pixel
    return 2.0 * d->batch * static_cast<double>(g.out_h) * g.out_w * 4 * d->cin_real * co;
  // ConvT counted as 9 taps per *input* pixel
  return 2.0 * d->batch * static_cast<double>(g.gemm_h) * g.gemm_w * taps * d->cin_real * co;
}

namespace jpdse {

// Coalesced packers for the layouts that carry 98 % of the generator's weights (a training step re-packs every conv
// after the optimizer step). The generic kernel above gathers one element per thread at a 36-byte (or, for the
// data-gradient layout, 36 KiB) stride; these stage a tile in shared memory so both sides move in contiguous runs.
// Same values bit for bit.
//   3x3 forward layout  out[n][t][c] = w[n][c][t]                 (CONV3X3_PAD1, CONV3X3_S2)
template <int kChunk>
__global__ void __launch_bounds__(kChunk / 2)
pack3x3_fwd_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cin) {
  __shared__ __align__(16) float s[kChunk * 9];
  const int n = blockIdx.x, c0 = blockIdx.y * kChunk;
  const float4* src = reinterpret_cast<const float4*>(w + (static_cast<size_t>(n) * cin + c0) * 9);
  for (int i = threadIdx.x; i < kChunk * 9 / 4; i += kChunk / 2) reinterpret_cast<float4*>(s)[i] = __ldg(src + i);
  __syncthreads();
  const int c = 2 * threadIdx.x;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(s[c * 9 + t], s[(c + 1) * 9 + t]);
    *reinterpret_cast<__nv_bfloat162*>(out + (static_cast<size_t>(n) * 9 + t) * cin + c0 + c) = h;
  }
}

//   3x3 data-gradient layout  out[n][t'][c] = w[c][n][8 - t']     (CONV3X3_FULL: n = forward ci, c = forward co)
__global__ void __launch_bounds__(256)
pack3x3_full_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cdim, int ndim) {
  constexpr int kRow = 16 * 9 + 1;
  __shared__ float s[64 * kRow];
  const int n0 = blockIdx.x * 16, c0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 64 * 36; i += 256) {
    const int c = i / 36, v = i - c * 36;
    const float4 f = __ldg(reinterpret_cast<const float4*>(w + (static_cast<size_t>(c0 + c) * ndim + n0) * 9) + v);
    float* dst = s + c * kRow + v * 4;
    dst[0] = f.x; dst[1] = f.y; dst[2] = f.z; dst[3] = f.w;
  }
  __syncthreads();
  const int pair = threadIdx.x & 31;
  for (int r = threadIdx.x >> 5; r < 16 * 9; r += 8) {
    const int nn = r / 9, t = r - nn * 9;
    const int k = nn * 9 + 8 - t;
    const __nv_bfloat162 h = __floats2bfloat162_rn(s[(2 * pair) * kRow + k], s[(2 * pair + 1) * kRow + k]);
    *reinterpret_cast<__nv_bfloat162*>(out + (static_cast<size_t>(n0 + nn) * 9 + t) * cdim + c0 + 2 * pair) = h;
  }
}

}  // namespace jpdse

extern "C" int jpdse_conv_pack_weights(const jpdse_conv_desc* d, const float* w, void* w_packed, void* stream) {
  ConvGeom g;
  int rc = conv_geom(d, &g);
  if (rc != JPDSE_OK) return rc;
  if (w == nullptr || w_packed == nullptr) return fail(JPDSE_ERR_INVALID, "pack_weights: NULL pointer");
  if (g.path == kPathConvtFused) return convt_fused_pack(d, w, w_packed, static_cast<cudaStream_t>(stream));
  {
    const char* e = getenv("JPDSE_GENERIC_PACK");  // tests: force the generic gather kernel
    const bool fast = !(e && e[0] == '1') && g.path != kPathRowHead && g.path != kPathRowStem && d->cin_real == d->cin &&
                      g.rows == d->cout && d->cin % 64 == 0 && d->cout % 16 == 0 &&
                      (d->cout_real == 0 || d->cout_real == d->cout);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (fast && (d->kind == JPDSE_CONV3X3_PAD1 || d->kind == JPDSE_CONV3X3_S2)) {
      if (d->cin % 256 == 0)
        pack3x3_fwd_kernel<256><<<dim3(d->cout, d->cin / 256), 128, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_packed), d->cin);
      else
        pack3x3_fwd_kernel<64><<<dim3(d->cout, d->cin / 64), 32, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_packed), d->cin);
      return check_launch("pack3x3_fwd_kernel");
    }
    if (fast && (d->kind == JPDSE_CONV3X3_FULL || d->kind == JPDSE_CONV3X3_FULL_SHARED)) {
      pack3x3_full_kernel<<<dim3(d->cout / 16, d->cin / 64), 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(w_packed), d->cin, d->cout);
      return check_launch("pack3x3_full_kernel");
    }
  }
  PackParams q{d->kind == JPDSE_CONV3X3_FULL_SHARED ? JPDSE_CONV3X3_FULL : d->kind, d->cin, d->cin_real, d->cout, g.rows, g.ktot, g.cpt, g.path,
               (d->cout_real > 0 && d->cout_real <= d->cout) ? d->cout_real : d->cout};
  const size_t total = static_cast<size_t>(g.rows) * g.ktot;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  pack_weights_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, static_cast<__nv_bfloat16*>(w_packed), q);
  return check_launch("pack_weights_kernel");
}

extern "C" int jpdse_conv_forward(const jpdse_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                                  void* y, double* stats, void* stream_v) {
  ConvGeom g;
  int rc = conv_geom(d, &g);
  if (rc != JPDSE_OK) return rc;
  if (x == nullptr || w_packed == nullptr || y == nullptr) return fail(JPDSE_ERR_INVALID, "conv_forward: NULL pointer");
  if (d->epilogue == JPDSE_EPI_RAW_STATS && stats == nullptr) return fail(JPDSE_ERR_INVALID, "conv_forward: stats is NULL");
  if ((d->epilogue == JPDSE_EPI_BIAS_TANH_NCHW || d->epilogue == JPDSE_EPI_BIAS_ACT || d->epilogue == JPDSE_EPI_BIAS_NCHW) &&
      bias == nullptr)
    return fail(JPDSE_ERR_INVALID, "conv_forward: bias is NULL");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w_packed) & 15) ||
      (reinterpret_cast<uintptr_t>(y) & 15))
    return fail(JPDSE_ERR_INVALID, "conv_forward: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (g.path == kPathConvtFused) return convt_fused_forward(d, x, w_packed, y, stats, stream);
  if (g.path == kPathPair) return pair_conv_forward(d, x, w_packed, y, stats, stream);
  if (g.path != kPathIgemm) return rowconv_forward(d, g.path == kPathRowHead, x, w_packed, bias, y, stats, stream);

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  const bool shared = d->kind == JPDSE_CONV3X3_FULL_SHARED;
  // The PatchGAN's stride-1 4x4 convs run on 131 / 67-pixel-wide grids: one-row tiles of 128 pixels waste half of every
  // second tile (131 = 128 + 3) or half of every tile (67 of 128). Over flat positions of the zero-bordered input
  // (pitch W+4: 3 dead columns per row) the same kernel wastes 2-4 %; taken when it saves more than 10 % of the tiles.
  bool flat_s1 = false;
  if (d->kind == JPDSE_CONV4X4_S1) {
    const char* e = getenv("JPDSE_FLAT_S1");  // "0": rectangular tiles (read per call: tests toggle it)
    int th = 0, tw = 0;
    pick_tile(g.gemm_w, &th, &tw, g.gemm_h);
    const long long rect = static_cast<long long>((g.gemm_h + th - 1) / th) * ((g.gemm_w + tw - 1) / tw);
    const long long pitch = d->in_w + 2 * d->in_pad;
    const long long flat_tiles = ((g.out_h - 1) * pitch + g.out_w + 127) / 128;
    flat_s1 = !(e != nullptr && e[0] == '0') && flat_tiles * 10 < rect * 9;
  }
  const bool flat = d->kind == JPDSE_CONV3X3_FULL || d->kind == JPDSE_CONV7X7_FULL || d->kind == JPDSE_CONV4X4_S1_FULL || shared ||
                    flat_s1;
  p.batch = d->batch;
  if (shared) {
    // the whole batch is one run of positions, pitch = output width: position m IS output pixel m of the dense
    // (B, H+2, W+2, Cout) result -- the kernel sees one image of B*(H+2) rows
    p.batch = 1;
    p.flat_pitch = g.out_w;
    p.tile_h = 1;
    p.tile_w = 128;
    p.tiles_h = 1;
    p.tiles_w = static_cast<int>((static_cast<long long>(d->batch) * g.out_h * g.out_w + 127) / 128);
  } else if (flat) {
    // M runs over flat positions y*pitch + x of the zero-bordered input; pitch = stored width
    p.flat_pitch = d->in_w + 2 * d->in_pad;
    const long long last = static_cast<long long>(g.out_h - 1) * p.flat_pitch + g.out_w;  // positions needed
    p.tile_h = 1;
    p.tile_w = 128;
    p.tiles_h = 1;
    p.tiles_w = static_cast<int>((last + 127) / 128);
  } else {
    pick_tile(g.gemm_w, &p.tile_h, &p.tile_w, g.gemm_h);
    p.tiles_h = (g.gemm_h + p.tile_h - 1) / p.tile_h;
    p.tiles_w = (g.gemm_w + p.tile_w - 1) / p.tile_w;
    p.gemm_h = g.gemm_h;
    p.gemm_w = g.gemm_w;
    p.partial = (g.gemm_h % p.tile_h || g.gemm_w % p.tile_w) ? 1 : 0;
  }
  p.n_tiles = g.rows / g.bn;
  p.chunks_per_tap = g.cpt;
  {
    // row mode (resident weights, one 130-pixel box per filter row; see kRow* above): 64 -> 64 channel 3x3 stride-1 convs
    // whose tile is one 128-pixel piece of a row (PAD1) or of the flat position run (the data-gradient kinds)
    const char* e = getenv("JPDSE_ROW_MODE");  // "0": the generic tap-by-tap K loop (read per call: tests toggle it)
    const bool on = !(e != nullptr && e[0] == '0');
    const bool shape = g.bn == 64 && g.rows == 64 && d->cin == 64 && g.cpt == 1 && g.ktot == 9 * 64;
    const bool tiles = (d->kind == JPDSE_CONV3X3_PAD1 && p.tile_h == 1 && p.tile_w == 128) || d->kind == JPDSE_CONV3X3_FULL || shared;
    p.rowmode = (on && shape && tiles) ? 1 : 0;
  }
  p.out_h = shared ? d->batch * g.out_h : g.out_h;
  p.out_w = g.out_w;
  p.os_h = p.os_w = 1;
  p.ldc = d->cout;
  p.n_valid = d->cout;
  p.epilogue = d->epilogue;
  p.out_pad = d->out_pad;
  p.slope = d->slope;
  p.out = y;
  p.stats = stats;
  p.bias = bias;
  p.dbg = g_dbg_enabled ? g_dbg : nullptr;
  {
    static int flags = -1;
    if (flags < 0) {
      const char* e = getenv("JPDSE_DEBUG_FLAGS");
      flags = e ? atoi(e) : 0;
    }
    p.dbg_flags = flags;
  }

  // bf16 NHWC outputs of the BN = 64 / 128 instantiations leave through shared memory + TMA store
  // (BN = 256 gives up a pipeline stage for the staging, so only where the epilogue is exposed: <= 48 k-blocks a tile)
  static int staged_limit = -1;  // k-blocks per tile up to which N = 256 trades its fourth stage for the staged epilogue
  if (staged_limit < 0) {
    const char* e = getenv("JPDSE_STAGED_KBLOCKS");
    staged_limit = e ? atoi(e) : 48;
  }
  int dev_taps = d->kind == JPDSE_CONV1X1 ? 1 : 9;  // taps of one tile's K loop as the kernel walks them
  if (d->kind == JPDSE_CONV4X4_S2) dev_taps = 8;
  if (d->kind == JPDSE_CONV4X4_S1) dev_taps = 16;
  if (d->kind == JPDSE_CONV4X4_S2_DGRAD) dev_taps = 4;
  if (d->kind == JPDSE_CONV3X3_PAD1_NARROW) dev_taps = 3;
  const int kblocks_per_tile = dev_taps * g.cpt;
  const bool odd_phase_out = d->kind == JPDSE_CONV4X4_S2_DGRAD && ((g.out_h | g.out_w) & 1);  // no {2C, W/2, 2, H/2} view
  const bool staged_out = (g.bn == 64 || g.bn == 128 || (g.bn == 256 && kblocks_per_tile <= staged_limit)) && !flat &&
                          (d->cout % 64) == 0 && d->kind != JPDSE_CONV7X7_PAD3 && !odd_phase_out &&
                          (d->epilogue == JPDSE_EPI_RAW_STATS || d->epilogue == JPDSE_EPI_RAW || d->epilogue == JPDSE_EPI_BIAS_ACT);
  const uint64_t C = static_cast<uint64_t>(d->cin);
  const uint64_t H = static_cast<uint64_t>(d->in_h), W = static_cast<uint64_t>(d->in_w), B = static_cast<uint64_t>(d->batch);
  // physical (stored) extent of x and the address of its logical pixel (0,0)
  const uint64_t Hp = H + 2 * static_cast<uint64_t>(d->in_pad), Wp = W + 2 * static_cast<uint64_t>(d->in_pad);
  const uint8_t* xin = static_cast<const uint8_t*>(x) + (static_cast<uint64_t>(d->in_pad) * Wp + d->in_pad) * C * 2;
  CUtensorMap ta, tb;
  uint64_t dims[5], strides[4];
  uint32_t box[5];

  if (d->kind == JPDSE_CONVT3X3_S2) {
    // A: (B,H,W,C) as {C, W, H, B}; +1 neighbours past the bottom/right edge are TMA zero fill
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = B;
    strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
    box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
    rc = make_tmap_bf16(&ta, xin, 4, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
    p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
    p.os_h = p.os_w = 2;
    CUtensorMap tc = ta;
    if (staged_out) {
      // output (B,2H,2W,Cout) as {2*Cout (column parity major), W, 2 (row parity), H, B}: a phase's tile is one box
      const uint64_t Co = static_cast<uint64_t>(d->cout);
      uint64_t cd[5] = {2 * Co, W, 2, H, B};
      uint64_t cs[4] = {2 * Co * 2, 2 * W * Co * 2, 2 * 2 * W * Co * 2, 2 * H * 2 * W * Co * 2};
      uint32_t cb[5] = {64, static_cast<uint32_t>(p.tile_w), 1, static_cast<uint32_t>(p.tile_h), 1};
      rc = make_tmap_bf16(&tc, y, 5, cd, cs, cb);
      if (rc != JPDSE_OK) return rc;
      p.tma_store = 1;
      p.c_rank = 5;
    }
    size_t k_elems_before = 0;  // rows * K of the previous phase blocks
    for (int pidx = 0; pidx < 4; ++pidx) {
      const int ph = pidx >> 1, pw = pidx & 1;
      int kh[2], dh[2], kw[2], dw[2];
      const int nth = convt_taps(ph, kh, dh), ntw = convt_taps(pw, kw, dw);
      p.ntaps = nth * ntw;
      memset(p.tap_off, 0, sizeof(p.tap_off));
      for (int i = 0; i < nth; ++i)
        for (int j = 0; j < ntw; ++j) {
          p.tap_off[i * ntw + j][1] = dw[j];
          p.tap_off[i * ntw + j][2] = dh[i];
        }
      p.op_h = ph; p.op_w = pw;
      const uint64_t kp = static_cast<uint64_t>(p.ntaps) * C;
      uint64_t bd[2] = {kp, static_cast<uint64_t>(g.rows)};
      uint64_t bs[1] = {kp * 2};
      uint32_t bb[2] = {64, static_cast<uint32_t>(g.bn)};
      rc = make_tmap_bf16(&tb, static_cast<const uint8_t*>(w_packed) + k_elems_before * 2, 2, bd, bs, bb);
      if (rc != JPDSE_OK) return rc;
      p.b_k_offset = 0;
      rc = launch_igemm_bn(g.bn, ta, tb, tc, p, stream);
      if (rc != JPDSE_OK) return rc;
      k_elems_before += static_cast<size_t>(g.rows) * kp;
    }
    return JPDSE_OK;
  }

  if (d->kind == JPDSE_CONV4X4_S2_DGRAD) {
    // dx_padded[2a+r, 2b+s] = sum_{i,j in {0,1}} dy[a-i, b-j] * W[2i+r][2j+s] in the coordinates of the forward input
    // padded by 2; with y = 2a'+r the interior pixel is a = a'+1, so phase (r, s) is a 2x2-tap GEMM over dy rows
    // a'+1-i, columns b'+1-j (index -1 never occurs; index h / w is the zero border behind the gradient).
    if (d->out_pad) return fail(JPDSE_ERR_UNSUPPORTED, "CONV4X4_S2_DGRAD writes a dense output (out_pad must be 0)");
    dims[0] = C; dims[1] = W + d->in_pad; dims[2] = H + d->in_pad; dims[3] = B;
    strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
    box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
    rc = make_tmap_bf16(&ta, xin, 4, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
    p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
    p.os_h = p.os_w = 2;
    p.check_out = 1;
    CUtensorMap tc = ta;
    if (staged_out) {
      const uint64_t Co = static_cast<uint64_t>(d->cout), GW = static_cast<uint64_t>(g.gemm_w), GH = static_cast<uint64_t>(g.gemm_h);
      uint64_t cd[5] = {2 * Co, GW, 2, GH, B};
      uint64_t cs[4] = {2 * Co * 2, 2 * GW * Co * 2, 2 * 2 * GW * Co * 2, 2 * GH * 2 * GW * Co * 2};
      uint32_t cb[5] = {64, static_cast<uint32_t>(p.tile_w), 1, static_cast<uint32_t>(p.tile_h), 1};
      rc = make_tmap_bf16(&tc, y, 5, cd, cs, cb);
      if (rc != JPDSE_OK) return rc;
      p.tma_store = 1;
      p.c_rank = 5;
    }
    const uint64_t kp = 4 * C;
    for (int pidx = 0; pidx < 4; ++pidx) {
      p.ntaps = 4;
      memset(p.tap_off, 0, sizeof(p.tap_off));
      for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) {
          p.tap_off[i * 2 + j][1] = 1 - j;
          p.tap_off[i * 2 + j][2] = 1 - i;
        }
      p.op_h = pidx >> 1; p.op_w = pidx & 1;
      uint64_t bd[2] = {kp, static_cast<uint64_t>(g.rows)};
      uint64_t bs[1] = {kp * 2};
      uint32_t bb[2] = {64, static_cast<uint32_t>(g.bn)};
      rc = make_tmap_bf16(&tb, static_cast<const uint8_t*>(w_packed) + static_cast<size_t>(pidx) * g.rows * kp * 2, 2, bd, bs, bb);
      if (rc != JPDSE_OK) return rc;
      p.b_k_offset = 0;
      rc = launch_igemm_bn(g.bn, ta, tb, tc, p, stream);
      if (rc != JPDSE_OK) return rc;
    }
    return JPDSE_OK;
  }

  if (flat) {
    // {K-inner, flat position, image}; reads past the last position of an image meet the next image's zero
    // border (or TMA zero fill behind the last image) and only feed rows that are never stored
    const uint64_t L = Hp * Wp;
    xin = static_cast<const uint8_t*>(x);
    dims[1] = L; dims[2] = B;
    strides[0] = C * 2; strides[1] = L * C * 2;
    box[0] = 64; box[1] = 128; box[2] = 1;
    p.a_rank = 3; p.dim_w = 1; p.dim_h = 1; p.dim_b = 2;
    if (shared) {
      // JPDSE_PAD_SHARED layout: B*(H+2)*(W+2) + 2*(W+2) + 2 stored positions, pitch W+2
      const uint64_t P = W + 2;
      dims[0] = C; dims[1] = B * (H + 2) * P + 2 * P + 2; dims[2] = 1;
      strides[1] = dims[1] * C * 2;
      p.ntaps = 9;
      for (int t = 0; t < 9; ++t) p.tap_off[t][1] = (t / 3) * static_cast<int>(P) + (t % 3);
      if (p.rowmode) {
        p.ntaps = 3;
        box[1] = 130;
        for (int t = 0; t < 3; ++t) p.tap_off[t][1] = t * static_cast<int>(P);
      }
    } else if (d->kind == JPDSE_CONV3X3_FULL) {
      dims[0] = C;
      p.ntaps = 9;
      for (int t = 0; t < 9; ++t) p.tap_off[t][1] = (t / 3) * static_cast<int>(Wp) + (t % 3);
      if (p.rowmode) {
        p.ntaps = 3;
        box[1] = 130;
        for (int t = 0; t < 3; ++t) p.tap_off[t][1] = t * static_cast<int>(Wp);
      }
    } else if (d->kind == JPDSE_CONV4X4_S1_FULL) {
      // dx[y, x] = sum_{kh', kw'} dy_padded[y + 1 + kh', x + 1 + kw'] * W[3-kh'][3-kw'] (dy stored with a zero border of 2)
      dims[0] = C;
      p.ntaps = 16;
      for (int t = 0; t < 16; ++t) p.tap_off[t][1] = (1 + t / 4) * static_cast<int>(Wp) + (1 + t % 4);
    } else if (d->kind == JPDSE_CONV4X4_S1) {
      // y[oy, ox] = sum_{kh, kw} x_stored[oy + kh, ox + kw] * W[kh][kw] (x stored with its zero border of 2)
      dims[0] = C;
      p.ntaps = 16;
      for (int t = 0; t < 16; ++t) p.tap_off[t][1] = (t / 4) * static_cast<int>(Wp) + (t % 4);
    } else {
      dims[0] = 64;  // 7*C window elements (+ zero-weight tail) under a filter row
      p.ntaps = 7;
      for (int t = 0; t < 7; ++t) p.tap_off[t][1] = t * static_cast<int>(Wp);
    }
  }
  switch (flat ? -1 : d->kind) {
    case JPDSE_CONV3X3_PAD1: {
      xin = static_cast<const uint8_t*>(x);  // taps address the stored border directly
      dims[0] = C; dims[1] = Wp; dims[2] = Hp; dims[3] = B;
      strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
      p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
      p.ntaps = 9;
      for (int t = 0; t < 9; ++t) {
        p.tap_off[t][1] = t % 3;
        p.tap_off[t][2] = t / 3;
      }
      if (p.rowmode) {
        p.ntaps = 3;
        box[1] = 130;  // the tile's 128 pixels + the two neighbours the kw = 1, 2 taps reach
        for (int t = 0; t < 3; ++t) {
          p.tap_off[t][1] = 0;
          p.tap_off[t][2] = t;
        }
      }
      break;
    }
    case JPDSE_CONV4X4_S1: {
      xin = static_cast<const uint8_t*>(x);  // taps address the stored zero border directly
      dims[0] = C; dims[1] = Wp; dims[2] = Hp; dims[3] = B;
      strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
      p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
      p.ntaps = 16;
      for (int t = 0; t < 16; ++t) {
        p.tap_off[t][1] = t % 4;
        p.tap_off[t][2] = t / 4;
      }
      break;
    }
    case JPDSE_CONV4X4_S2: {
      // the stored (B,H+4,W+4,C) tensor viewed as {2C, (W+4)/2, 2, (H+4)/2, B}: tap (kh, column pair kp) covers kw = 2kp
      // and 2kp+1 as 2C contiguous elements; output (oh, ow) reads row 2*oh+kh = pair oh + kh/2, parity kh % 2
      xin = static_cast<const uint8_t*>(x);
      dims[0] = 2 * C; dims[1] = Wp / 2; dims[2] = 2; dims[3] = Hp / 2; dims[4] = B;
      strides[0] = 2 * C * 2; strides[1] = Wp * C * 2; strides[2] = 2 * Wp * C * 2; strides[3] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = 1; box[3] = p.tile_h; box[4] = 1;
      p.a_rank = 5; p.dim_w = 1; p.dim_h = 3; p.dim_b = 4;
      p.ntaps = 8;
      for (int t = 0; t < 8; ++t) {
        const int kh = t / 2, kp = t % 2;
        p.tap_off[t][1] = kp;
        p.tap_off[t][2] = kh & 1;
        p.tap_off[t][3] = kh >> 1;
      }
      break;
    }
    case JPDSE_CONV1X1: {
      dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = B;
      strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
      p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
      p.ntaps = 1;
      break;
    }
    case JPDSE_CONV3X3_S2: {
      // (B,H,W,C) viewed as {2C, W/2, 2, H/2, B}: a column pair is one 2C-wide "pixel", rows split
      // into (pair, parity). Tap kh -> input row 2*oh+kh-1 = pair oh-1 parity 1 | pair oh parity 0 | 1.
      dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B;
      strides[0] = 2 * C * 2; strides[1] = Wp * C * 2; strides[2] = 2 * Wp * C * 2; strides[3] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = 1; box[3] = p.tile_h; box[4] = 1;
      p.a_rank = 5; p.dim_w = 1; p.dim_h = 3; p.dim_b = 4;
      p.ntaps = 9;
      for (int t = 0; t < 9; ++t) {
        const int kh = t / 3, kw = t % 3;
        p.tap_off[t][0] = (kw == 1) ? 0 : static_cast<int>(C);
        p.tap_off[t][1] = (kw == 0) ? -1 : 0;
        p.tap_off[t][2] = (kh == 1) ? 0 : 1;
        p.tap_off[t][3] = (kh == 0) ? -1 : 0;
      }
      break;
    }
    case JPDSE_CONV3X3_PAD1_NARROW: {
      // as the 7x7 case below with three filter rows: window {64 (3*C real), W, H+2, B}, pixel stride C elements
      xin = static_cast<const uint8_t*>(x);
      dims[0] = 64; dims[1] = W; dims[2] = Hp; dims[3] = B;
      strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
      p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
      p.ntaps = 3;
      for (int t = 0; t < 3; ++t) p.tap_off[t][2] = t;
      break;
    }
    case JPDSE_CONV7X7_PAD3: {
      // one "tap" per filter row: the 7*C elements under a filter row are contiguous in NHWC, so
      // the A operand is an overlapping-window view {7C (padded to cpt*64), W, H+6, B} with a
      // pixel stride of C elements. Elements past 7*C meet zero weights.
      xin = static_cast<const uint8_t*>(x);
      dims[0] = static_cast<uint64_t>(g.cpt) * 64; dims[1] = W; dims[2] = Hp; dims[3] = B;
      strides[0] = C * 2; strides[1] = Wp * C * 2; strides[2] = Hp * Wp * C * 2;
      box[0] = 64; box[1] = p.tile_w; box[2] = p.tile_h; box[3] = 1;
      p.a_rank = 4; p.dim_w = 1; p.dim_h = 2; p.dim_b = 3;
      p.ntaps = 7;
      for (int t = 0; t < 7; ++t) p.tap_off[t][2] = t;
      break;
    }
  }
  rc = make_tmap_bf16(&ta, xin, p.a_rank, dims, strides, box);
  if (rc != JPDSE_OK) return rc;
  uint64_t bd[2] = {static_cast<uint64_t>(g.ktot), static_cast<uint64_t>(g.rows)};
  uint64_t bs[1] = {static_cast<uint64_t>(g.ktot) * 2};
  uint32_t bb[2] = {64, static_cast<uint32_t>(g.bn)};
  rc = make_tmap_bf16(&tb, w_packed, 2, bd, bs, bb);
  if (rc != JPDSE_OK) return rc;
  CUtensorMap tc = ta;
  if (staged_out) {
    // dims = the logical output (the store clips overhanging tiles there); strides / base step over the caller's border
    const uint64_t Co = static_cast<uint64_t>(d->cout), OW = static_cast<uint64_t>(g.out_w), OH = static_cast<uint64_t>(g.out_h);
    const uint64_t OP = static_cast<uint64_t>(d->out_pad), OWp = OW + 2 * OP, OHp = OH + 2 * OP;
    uint64_t cd[4] = {Co, OW, OH, B};
    uint64_t cs[3] = {Co * 2, OWp * Co * 2, OHp * OWp * Co * 2};
    uint32_t cb[4] = {64, static_cast<uint32_t>(p.tile_w), static_cast<uint32_t>(p.tile_h), 1};
    rc = make_tmap_bf16(&tc, static_cast<uint8_t*>(y) + (OP * OWp + OP) * Co * 2, 4, cd, cs, cb);
    if (rc != JPDSE_OK) return rc;
    p.tma_store = 1;
    p.c_rank = 4;
  }
  return launch_igemm_bn(g.bn, ta, tb, tc, p, stream);
}
