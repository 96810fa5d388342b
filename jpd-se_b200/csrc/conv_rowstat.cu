// Row-stationary 7x7 convolutions (the full-resolution stem and head of GlobalGenerator,
// reference ctu/models/pix2pixHD_networks/networks.py:210 and :246).
//
// The generic implicit-GEMM kernel re-fetches every activation once per filter tap; at full resolution
// that makes the two 7x7 convs L2-bandwidth bound (49 taps). Here each 128-pixel piece of an INPUT row is
// fetched by TMA exactly once and multiplied by the weights of all 7 filter rows while it sits in shared
// memory: ONE tcgen05.mma of N = 7 x 32 columns per K step sends the 7 products to the TMEM accumulators of
// the 7 OUTPUT rows that input row touches (a ring of 16 accumulator slots of 32 columns; consecutive output
// rows are consecutive slots, so the 7 filter rows are consecutive columns). All weights stay resident in
// shared memory. Small-N MMAs cost ~68 cycles regardless of N (measured, tools/mma_bench.cu), which is why the
// 7 filter rows must share one instruction.
//
//   head  (64 -> 3, +bias, tanh):  N = (kw, cout) = 21 -> 32 columns, K = 64 channels. The kw shift is resolved
//          in the epilogue: out[ow] = sum_kw P[ow + kw][kw, co]  (a 7-wide shifted sum through shared memory).
//   stem  (39(40) -> 64, raw + InstanceNorm statistics): K = the 7*40 contiguous elements under a filter row,
//          read through an overlapping-window tensor map (pixel stride 40 elements, 5 k-blocks of 64);
//          N = 32 of the 64 output channels per CTA so that 7 x 5 x 4 KiB of weights fit in shared memory.
//
// Warp roles as in conv_igemm.cu: warps 0..3 / 4..7 two epilogue warpgroups that drain alternate output rows,
// warp 8 TMA producer, warp 9 TMEM owner + MMA issuer.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "conv_shared.cuh"
#include "ptx.cuh"

namespace jpdse {

constexpr int kRowN = 32;                          // accumulator columns per slot / MMA N
constexpr int kRowSlots = 16;                      // accumulator ring (7 live output rows + 9 draining / free)
constexpr int kRowBTile = kRowN * 64 * 2;          // one (kb, q) weight tile, bytes; q = 6 - kh
constexpr int kRowATile = 128 * 64 * 2;            // one A k-block, bytes
constexpr int kRowThreads = 320;                   // two epilogue warpgroups (alternate rows) + 2 control warps
constexpr int kRowProducerWarp = 8;                // control warps take the highest ids (arbiter priority)
constexpr int kRowMmaWarp = 9;

struct RowConvParams {
  int batch, height, width;   // output = logical input size
  int strips, strip_step;     // column strips per row; first position of strip s = s * strip_step
  int strip_valid;            // output columns produced per strip (122 for the head, 128 for the stem)
  int chunks, chunk_rows;     // row chunks per strip
  int n_splits;               // 1 (head) or 2 (stem: output channel halves)
  int cout;                   // real output channels
  void* out;
  double* stats;
  const float* bias;
  int mc;         // 1: the two channel-split CTAs of a strip form a cluster and share every A box by TMA multicast
  int dbg_flags;  // developer experiments (JPDSE_DEBUG_FLAGS, as in conv_igemm.cu): 1 skip statistics | 2 skip the output
                  // stores | 8 hand the accumulator straight back | 16 skip the MMAs
};

template <int KB, bool kHead>
struct RowCfg {
  static constexpr int kStages = kHead ? 8 : 4;
  static constexpr int kBBytes = 7 * KB * kRowBTile;
  static constexpr int kScratchFloats = kHead ? 2 * 2 * 28 * 136 : 8 * 32 * 17;
  static constexpr int kSmemBytes = 1024 + kBBytes + kStages * kRowATile + kScratchFloats * 4 + 512;
};

template <int KB, bool kHead, bool kMc>
__global__ void __launch_bounds__(kRowThreads, 1)
rowconv_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
               const __grid_constant__ RowConvParams p) {
  using Cfg = RowCfg<KB, kHead>;
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment as an OFFSET from the shared window (keeps the pointers in the shared address space)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_b = smem;                                   // resident weights
  uint8_t* s_a = smem + Cfg::kBBytes;                    // A ring
  float* s_scr = reinterpret_cast<float*>(s_a + Cfg::kStages * kRowATile);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_scr + Cfg::kScratchFloats);
  uint64_t* full_bar = bars;                             // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;             // [kStages]
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;        // [kRowSlots]
  uint64_t* tempty_bar = tfull_bar + kRowSlots;          // [kRowSlots]
  uint64_t* bfull_bar = tempty_bar + kRowSlots;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == kRowProducerWarp && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kMc ? 2 : 1);  // multicast: a slot is free once BOTH CTAs' MMAs have read it
    }
    for (int i = 0; i < kRowSlots; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(bfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == kRowMmaWarp) {
    tmem_alloc<1>(tmem_slot, kRowSlots * kRowN);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  if (kMc) cluster_sync_all();  // the peer's barriers exist before anything is multicast at them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work decomposition: item -> (image, strip, row chunk); CTAs with the same channel split share `split`
  const int split = kHead ? 0 : static_cast<int>(blockIdx.x) % p.n_splits;
  const int cta = kHead ? static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x) / p.n_splits;
  const int ctas = kHead ? static_cast<int>(gridDim.x) : static_cast<int>(gridDim.x) / p.n_splits;
  const int items = p.batch * p.strips * p.chunks;

  if (warp == kRowProducerWarp) {
    // ================================================================== TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(bfull_bar, Cfg::kBBytes);
      for (int kb = 0; kb < KB; ++kb)
        tma_load_2d(&tm_b, bfull_bar, s_b + kb * 7 * kRowBTile, 0, (split * KB + kb) * 7 * kRowN);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t it = 0;  // running A box number
      for (int item = cta; item < items; item += ctas) {
        const int chunk = item % p.chunks;
        const int strip = (item / p.chunks) % p.strips;
        const int b = item / (p.chunks * p.strips);
        const int r0 = chunk * p.chunk_rows;
        const int rows = min(p.chunk_rows, p.height - r0);
        const int iw0 = strip * p.strip_step;
        // the shared-memory ring is shallow (the weights take most of it): pull rows into L2 ahead of the loads
        constexpr int kAhead = kHead ? 8 : 4;
        if (split == 0)
          for (int i = 0; i < kAhead && i < rows + 6; ++i)
            for (int kb = 0; kb < KB; ++kb) tma_prefetch_4d(&tm_a, kb * 64, iw0, r0 + i, b);
        for (int i = 0; i < rows + 6; ++i) {
          if (split == 0 && i + kAhead < rows + 6)
            for (int kb = 0; kb < KB; ++kb) tma_prefetch_4d(&tm_a, kb * 64, iw0, r0 + i + kAhead, b);
          for (int kb = 0; kb < KB; ++kb, ++it) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full_bar[stage], kRowATile);
            if (!kMc)
              tma_load_4d(&tm_a, &full_bar[stage], s_a + stage * kRowATile, kb * 64, iw0, r0 + i, b);
            else if ((it & 1u) == static_cast<uint32_t>(split))  // the pair walks the same boxes: each loads every other one
              tma_load_4d_mc(&tm_a, &full_bar[stage], s_a + stage * kRowATile, kb * 64, iw0, r0 + i, b, 0x3);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kRowMmaWarp) {
    // ================================================================== MMA issuer
    if (elect_one()) {
      mbar_wait(bfull_bar, 0);
      tc_fence_after();
      const uint64_t b_desc0 = umma_smem_desc_sw128(smem_u32(s_b));
      int stage = 0;
      uint32_t phase = 0;
      uint32_t n0 = 0;  // output rows started by this CTA so far (row n lives in ring slot n % 16)
      for (int item = cta; item < items; item += ctas) {
        const int chunk = item % p.chunks;
        const int rows = min(p.chunk_rows, p.height - chunk * p.chunk_rows);
        for (int i = 0; i < rows + 6; ++i) {
          if (i < rows) {  // output row i starts with this input row: its accumulator slot must be drained
            const uint32_t n = n0 + i;
            mbar_wait(&tempty_bar[n & (kRowSlots - 1)], ((n / kRowSlots) & 1) ^ 1);
            tc_fence_after();
          }
          // input row i feeds output rows j = i-6+q, q = 0..6 (filter row kh = 6-q); keep those inside the chunk.
          // The MMA "program" of this input row is built once (the single issuing thread is latency bound, so the
          // K loop below must be nothing but descriptor adds + tcgen05.mma): runs of consecutive ring slots, one
          // MMA of N = 32 * len columns each; split where the ring wraps, and -- on the very first K step only --
          // in front of the row that starts here (q = 6, kh = 0), which is zero-initialised instead of accumulated.
          const int q_lo = max(0, 6 - i);
          const int q_hi = min(6, rows + 5 - i);
          const bool starts = (q_hi == 6);  // an output row starts with this input row (its slot is zero-initialised)
          const uint32_t slot_lo = (n0 + static_cast<uint32_t>(i - 6 + q_lo)) & (kRowSlots - 1);
          // regular K step: rows q_lo..q_hi -> run A (up to the ring end) + run B (wrapped remainder, may be empty)
          const int tot = q_hi - q_lo + 1;
          const int len_a = min(tot, kRowSlots - static_cast<int>(slot_lo));
          const int len_b = tot - len_a;
          const uint32_t d_a = tmem_base + slot_lo * kRowN, d_b = tmem_base;
          const uint64_t b_a = static_cast<uint64_t>((q_lo * kRowBTile) >> 4);
          const uint64_t b_b = static_cast<uint64_t>(((q_lo + len_a) * kRowBTile) >> 4);
          const uint32_t id_a = umma_idesc_bf16(128, kRowN * len_a), id_b = umma_idesc_bf16(128, kRowN * (len_b > 0 ? len_b : 1));
          // first K step of a starting row: rows q_lo..5 accumulate (runs A0/B0), row 6 overwrites (run Z)
          const int tot0 = starts ? tot - 1 : tot;
          const int len_a0 = min(tot0, kRowSlots - static_cast<int>(slot_lo));
          const int len_b0 = tot0 - len_a0;
          const uint64_t b_b0 = static_cast<uint64_t>(((q_lo + len_a0) * kRowBTile) >> 4);
          const uint32_t id_a0 = umma_idesc_bf16(128, kRowN * (len_a0 > 0 ? len_a0 : 1));
          const uint32_t id_b0 = umma_idesc_bf16(128, kRowN * (len_b0 > 0 ? len_b0 : 1));
          const uint32_t d_z = tmem_base + ((n0 + static_cast<uint32_t>(i)) & (kRowSlots - 1)) * kRowN;
          const uint64_t b_z = static_cast<uint64_t>((6 * kRowBTile) >> 4);
          constexpr uint32_t id_z = umma_idesc_bf16(128, kRowN);
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {  // fully unrolled: the kb == 0 / last-block cases below fold at compile time
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t adesc = umma_smem_desc_sw128(smem_u32(s_a + stage * kRowATile));
            const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((kb * 7 * kRowBTile) >> 4);
            if (p.dbg_flags & 16) {
              if (kMc) umma_commit_mc(&empty_bar[stage], 0x3);
              else umma_commit(&empty_bar[stage]);
              if (++stage == Cfg::kStages) {
                stage = 0;
                phase ^= 1;
              }
              continue;
            }
            if (kb == 0) {
              if (len_a0 > 0) umma_bf16<1>(d_a, adesc, bdesc + b_a, id_a0, 1u);
              if (len_b0 > 0) umma_bf16<1>(d_b, adesc, bdesc + b_b0, id_b0, 1u);
              if (starts) umma_bf16<1>(d_z, adesc, bdesc + b_z, id_z, 0u);
            }
            // stem: the window under a filter row has 7 * 40 = 280 elements; the last k-block (256..319) only carries
            // 24 of them, so its upper two K = 16 steps meet nothing but zero weights
            const int ksteps = (!kHead && kb == KB - 1) ? 2 : 4;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if ((kb == 0 && k == 0) || k >= ksteps) continue;
              umma_bf16<1>(d_a, adesc + static_cast<uint64_t>(k * 2), bdesc + b_a + static_cast<uint64_t>(k * 2), id_a, 1u);
              if (len_b > 0)
                umma_bf16<1>(d_b, adesc + static_cast<uint64_t>(k * 2), bdesc + b_b + static_cast<uint64_t>(k * 2), id_b, 1u);
            }
            if (kMc) umma_commit_mc(&empty_bar[stage], 0x3);
            else umma_commit(&empty_bar[stage]);
            if (++stage == Cfg::kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
          if (i >= 6) umma_commit(&tfull_bar[(n0 + i - 6) & (kRowSlots - 1)]);  // row i-6 has all 7 filter rows
        }
        n0 += rows;
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue (warps 0..3 / 4..7)
    const int group = warp >> 2;          // warpgroup g drains the output rows whose running number is == g (mod 2)
    const int quarter = warp & 3;
    const int pos = quarter * 32 + lane;  // position inside the strip = TMEM lane
    uint32_t n0 = 0;
    uint32_t grp_rows = 0;                // rows this group has drained (selects the head's double buffer)
    for (int item = cta; item < items; item += ctas) {
      const int chunk = item % p.chunks;
      const int strip = (item / p.chunks) % p.strips;
      const int b = item / (p.chunks * p.strips);
      const int r0 = chunk * p.chunk_rows;
      const int rows = min(p.chunk_rows, p.height - r0);
      const int ow = strip * p.strip_step + pos;
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;  // stem: column-pair statistics over this item's rows
      for (int j = 0; j < rows; ++j) {
        const uint32_t n = n0 + j;
        if (static_cast<int>(n & 1) != group) continue;
        const int slot = n & (kRowSlots - 1);
        mbar_wait_parked(&tfull_bar[slot], (n / kRowSlots) & 1);
        tc_fence_after();
        if (p.dbg_flags & 8) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[slot]);
          continue;
        }
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slot * kRowN, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[slot]);
        const int oh = r0 + j;
        if constexpr (kHead) {
          // P[kw*cout+co][pos] -> shared (double buffered), then out[ow] = tanh(bias + sum_kw P[kw, co][pos + kw])
          const int np = 7 * p.cout;
          float* sp = s_scr + (group * 2 + static_cast<int>(grp_rows & 1)) * (28 * 136);
#pragma unroll
          for (int q = 0; q < 28; ++q)
            if (q < np) sp[q * 136 + pos] = __uint_as_float(v[q]);
          if (group == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
          else asm volatile("bar.sync 2, 128;" ::: "memory");
          if (pos < p.strip_valid && ow < p.width) {
            const size_t plane = static_cast<size_t>(p.height) * p.width;
            float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(b) * p.cout * plane +
                       static_cast<size_t>(oh) * p.width + ow;
            for (int co = 0; co < p.cout; ++co) {
              float s = p.bias[co];
#pragma unroll
              for (int kw = 0; kw < 7; ++kw) s += sp[(kw * p.cout + co) * 136 + pos + kw];
              // tanh(s) = 1 - 2 / (exp(2s) + 1); exp2-based, |err| ~ 1e-7, saturates cleanly at +-1
              const float e = __expf(2.f * s);
              o[co * plane] = 1.f - __fdividef(2.f, e + 1.f);
            }
          }
          // this group's next row writes the other buffer; the one after that is ordered behind the next barrier
        } else {
          __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                ((static_cast<size_t>(b) * p.height + oh) * p.width + ow) * p.cout + split * kRowN;
          uint32_t pk[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * t]), __uint_as_float(v[2 * t + 1]));
            pk[t] = *reinterpret_cast<uint32_t*>(&h);
          }
          uint4* dst = reinterpret_cast<uint4*>(orow);
          if (!(p.dbg_flags & 2)) {
#pragma unroll
            for (int t = 0; t < 4; ++t) dst[t] = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
          }
          if (p.dbg_flags & 1) continue;
          // column sums through a per-warp 32x16-word shared transpose: lane l owns column pair (l & 15) over
          // rows 16*(l >> 4) .. +15 (see conv_igemm.cu)
          uint32_t* s_t = reinterpret_cast<uint32_t*>(s_scr) + warp * (32 * 17);
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 16; ++t) s_t[lane * 17 + t] = pk[t];
          __syncwarp();
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
        }
        ++grp_rows;
      }
      if constexpr (!kHead) {
        // one atomic per column per warp per item
        s1a += __shfl_xor_sync(0xffffffffu, s1a, 16);
        s1b += __shfl_xor_sync(0xffffffffu, s1b, 16);
        s2a += __shfl_xor_sync(0xffffffffu, s2a, 16);
        s2b += __shfl_xor_sync(0xffffffffu, s2b, 16);
        if (lane < 16) {
          double* st = p.stats + (static_cast<size_t>(b) * p.cout + split * kRowN + lane * 2) * 2;
          atomicAdd(st + 0, static_cast<double>(s1a));
          atomicAdd(st + 1, static_cast<double>(s2a));
          atomicAdd(st + 2, static_cast<double>(s1b));
          atomicAdd(st + 3, static_cast<double>(s2b));
        }
      }
      n0 += rows;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kMc) cluster_sync_all();  // no CTA leaves while the peer may still arrive on its barriers
  if (warp == kRowMmaWarp) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kRowSlots * kRowN);
  }
}

template <int KB, bool kHead, bool kMc>
static int launch_rowconv(const CUtensorMap& ta, const CUtensorMap& tb, const RowConvParams& p, cudaStream_t stream) {
  using Cfg = RowCfg<KB, kHead>;
  static DeviceOnce configured;  // the attribute is per device: set it on each device this process uses
  if (configured.first_use()) {
    cudaError_t e = cudaFuncSetAttribute(rowconv_kernel<KB, kHead, kMc>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "cudaFuncSetAttribute(rowconv smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
    configured.done();
  }
  const int items = p.batch * p.strips * p.chunks;
  int grid = num_sms();
  if (kHead) {
    if (grid > items) grid = items;
  } else {
    grid -= grid % p.n_splits;
    if (grid > items * p.n_splits) grid = items * p.n_splits;
  }
  if (kMc) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kRowThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, rowconv_kernel<KB, kHead, kMc>, ta, tb, p);
    if (e != cudaSuccess) return fail(JPDSE_ERR_CUDA, "rowconv cluster launch: %s", cudaGetErrorString(e));
    return check_launch("rowconv_kernel");
  }
  rowconv_kernel<KB, kHead, kMc><<<grid, kRowThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  return check_launch("rowconv_kernel");
}

// Host entry used by jpdse_conv_forward for the two row-stationary paths (see conv_geom in conv_igemm.cu).
int rowconv_forward(const jpdse_conv_desc* d, bool head, const void* x, const void* w_packed, const float* bias, void* y,
                    double* stats, cudaStream_t stream) {
  RowConvParams p;
  memset(&p, 0, sizeof(p));
  p.batch = d->batch;
  p.height = d->in_h;
  p.width = d->in_w;
  p.cout = d->cout;
  p.out = y;
  p.stats = stats;
  p.bias = bias;
  {
    static int flags = -1;
    if (flags < 0) {
      const char* e = getenv("JPDSE_DEBUG_FLAGS");
      flags = e ? atoi(e) : 0;
    }
    p.dbg_flags = flags;
  }
  p.chunk_rows = d->in_h < 128 ? d->in_h : 128;
  p.chunks = (d->in_h + p.chunk_rows - 1) / p.chunk_rows;
  const uint64_t C = static_cast<uint64_t>(d->cin), B = static_cast<uint64_t>(d->batch);
  const uint64_t Hp = static_cast<uint64_t>(d->in_h) + 6, Wp = static_cast<uint64_t>(d->in_w) + 6;
  CUtensorMap ta, tb;
  uint64_t dims[4], strides[3];
  uint32_t box[4] = {64, 128, 1, 1};
  strides[0] = C * 2;
  strides[1] = Wp * C * 2;
  strides[2] = Hp * Wp * C * 2;
  dims[2] = Hp;
  dims[3] = B;
  int rc;
  if (head) {
    p.n_splits = 1;
    p.strip_valid = 122;  // 128 positions give 122 outputs (needs positions ow .. ow+6)
    p.strip_step = 122;
    p.strips = (d->in_w + 121) / 122;
    dims[0] = 64;
    dims[1] = Wp;
    rc = make_tmap_bf16(&ta, x, 4, dims, strides, box);
    if (rc != JPDSE_OK) return rc;
    uint64_t bd[2] = {64, 7ull * kRowN}, bs[1] = {128};
    uint32_t bb[2] = {64, 7 * kRowN};
    rc = make_tmap_bf16(&tb, w_packed, 2, bd, bs, bb);
    if (rc != JPDSE_OK) return rc;
    return launch_rowconv<1, true, false>(ta, tb, p, stream);
  }
  p.n_splits = d->cout / kRowN;
  {
    const char* e = getenv("JPDSE_STEM_MULTICAST");  // "0": every CTA loads its own copy of the A boxes
    p.mc = (p.n_splits == 2 && !(e && e[0] == '0')) ? 1 : 0;
  }
  p.strip_valid = 128;
  p.strip_step = 128;
  p.strips = d->in_w / 128;
  dims[0] = 320;  // 7*C = 280 elements under a filter row, padded to 5 k-blocks (zero weights behind them)
  dims[1] = static_cast<uint64_t>(d->in_w);
  rc = make_tmap_bf16(&ta, x, 4, dims, strides, box);
  if (rc != JPDSE_OK) return rc;
  uint64_t bd[2] = {64, static_cast<uint64_t>(p.n_splits) * 7 * 5 * kRowN}, bs[1] = {128};
  uint32_t bb[2] = {64, 7 * kRowN};
  rc = make_tmap_bf16(&tb, w_packed, 2, bd, bs, bb);
  if (rc != JPDSE_OK) return rc;
  return p.mc ? launch_rowconv<5, false, true>(ta, tb, p, stream) : launch_rowconv<5, false, false>(ta, tb, p, stream);
}

}  // namespace jpdse
