// VAE decoder kernels (reference modules/autoencoder.py:203-283 `Decoder`, :22-50 `AttnBlock`,
// :53-94 `ResnetBlock`, :112-123 `Upsample`, run under torch.autocast(bf16) by flux_pipeline.py:423-438).
//
//   conv_igemm_kernel   3x3 / 1x1 convolution as an implicit GEMM on the bf16 tensor cores: activations are NHWC bf16,
//                       the A operand of a 128-pixel tile and one (tap, 64-channel) slice is ONE 4-D TMA box whose
//                       out-of-bounds pixels read as zero (= the convolution's zero padding); weights are pre-packed
//                       [Cout][tap][Cin]; tcgen05.mma.cta_group::2 (256 pixels x BN channels per SM pair), TMEM
//                       double-buffered accumulators, persistent tiles.  The same kernel is the dense bf16 GEMM of the
//                       attention block (1x1 "convolution" over a [rows, K] matrix).
//   gn_stats / gn_apply GroupNorm(32 groups, eps 1e-6, affine) statistics (fp64 accumulation) and normalise (+ swish),
//                       fp32 arithmetic, ONE rounding to bf16 -- what autocast does: group_norm runs in fp32 and the
//                       following convolution casts its input to bf16.
//   upsample2x          nearest-neighbour, NHWC.
//   softmax_rows        fp32 scores -> bf16 probabilities for the single-head d = 512 attention of the mid block.
//   latent_prep         z / scale_factor + shift_factor (fp32, autoencoder.py:331-332) -> NHWC bf16, channels zero-padded to 64.
#include <cuda_bf16.h>

#include <cstdlib>
#include <type_traits>

#include "host_util.h"
#include "ptx.cuh"

namespace fb {

constexpr int kConvThreads = 320;  // warp 0 TMA, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int kConvEpiWarp0 = 2;
constexpr int kConvABytes = 128 * 128;  // 128 pixels x 64 channels x bf16

struct ConvParams {
  CUtensorMap tmap_a;  // activations [B][H][W][Cin]
  CUtensorMap tmap_b;  // weights [N][taps * Cin]
  int B, H, W, Cin, N, taps;  // H, W: OUTPUT rows / columns (= the input's for stride 1)
  int stride;                 // 1 (padding 1 all round) | 2 (Downsample: padding 0 left / top, 1 right / bottom)
  int tw, tw_shift, th;  // pixel tile: th rows x tw columns, th * tw = 128
  int tiles_x, tiles_y, m_tiles, n_tiles, num_tiles;  // num_tiles counts (pair of pixel tiles) x (channel tile)
  int kchunks;                                        // Cin / 64
  const __nv_bfloat16* bias;
  const __nv_bfloat16* residual;
  int64_t ld_res;
  void* out;
  int64_t ldo;
  int out_mode;  // 0 bf16 NHWC, 1 fp32 row-major scaled by alpha, 2 bf16 NCHW
  int wide;      // out_mode 0: 32-byte accesses (N % 16 == 0, 32-byte aligned rows)
  float alpha;
  double* gn_stats;  // out_mode 0, optional: [B][32][2] sums / sums of squares of the stored output (next GroupNorm)
  int gn_cg;         // channels per group of that GroupNorm (N / 32)
};

constexpr int kConvStatsBatch = 4;  // the epilogue statistics are accumulated per CTA for up to this many images

// HALO (3x3, 128-pixel row strips): one pipeline stage holds the 130 pixels [x0 - 1, x0 + 129) of ONE input row and
// 64 channels, and the weights of the three horizontal taps of that row: the dx = -1 / 0 / +1 operands of the MMAs are
// the SAME shared-memory rows read from a start address 0 / 1 / 2 rows into the tile, so a 3x3 convolution fetches each
// input row three times instead of nine (the L2 -> shared-memory stream is what bounds the 128- and 256-channel levels).
constexpr int kHaloRows = 130;
constexpr int kHaloABytes = 17 * 1024;  // 130 x 128 B, rounded up so the weight tiles stay 1024-byte aligned

template <int BN, bool HALO>
struct ConvSmem {
  static constexpr int kBRows = BN / 2;  // this CTA's half of the weight rows
  static constexpr int kA = HALO ? kHaloABytes : kConvABytes;
  static constexpr int kBTile = kBRows * 128;
  static constexpr int kStage = kA + (HALO ? 3 : 1) * kBTile;
  static constexpr int kTxBytes = (HALO ? kHaloRows * 128 : kConvABytes) + (HALO ? 3 : 1) * kBTile;  // per CTA and stage
  static constexpr int kStages = HALO ? (BN == 256 ? 3 : 4) : (BN == 256 ? 6 : 8);
  static constexpr int kBarOff = kStages * kStage;
  static constexpr int kTotal = kBarOff + 512 + 1024;
};

template <int BN, bool HALO>
__global__ void __launch_bounds__(kConvThreads, 1) conv_igemm_kernel(const __grid_constant__ ConvParams P) {
  using S = ConvSmem<BN, HALO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* empty_bar = full_bar + S::kStages;
  uint64_t* tfull_bar = empty_bar + S::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  __shared__ double gn_acc[8][kConvStatsBatch][32][2];  // per epilogue warp: no shared-memory atomics

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int tile0 = blockIdx.x >> 1;
  const int tile_stride = gridDim.x >> 1;
  constexpr int kEpiWarps = (kConvThreads / 32) - kConvEpiWarp0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmap_a);
    tma_prefetch_desc(&P.tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S::kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiWarps * 2);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr, 2 * BN);
    tmem_relinquish_2sm();
  }
  if (P.gn_stats != nullptr)
    for (int i = threadIdx.x; i < 8 * kConvStatsBatch * 64; i += kConvThreads) (&gn_acc[0][0][0][0])[i] = 0.0;
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();
  pdl_launch_dependents();

  const int num_tiles = P.num_tiles;
  const int kiters = (HALO ? 3 : P.taps) * P.kchunks;  // pipeline stages per tile

  // pixel tile `mt` -> (batch, first row, first column); tiles past the end are clamped (their rows are never stored)
  auto tile_origin = [&](int mt, int& b, int& y0, int& x0) {
    if (mt >= P.m_tiles) mt = P.m_tiles - 1;
    const int tx = mt % P.tiles_x;
    const int t2 = mt / P.tiles_x;
    const int ty = t2 % P.tiles_y;
    b = t2 / P.tiles_y;
    y0 = ty * P.th;
    x0 = tx * P.tw;
  };

  if (warp == 0) {
    // ---- TMA producer (both CTAs: own 128 pixels, own half of the weight rows; bytes accounted on the leader) ----
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
      const int n_blk = tile % P.n_tiles;
      const int mt = (tile / P.n_tiles) * 2 + static_cast<int>(rank);
      int b, y0, x0;
      tile_origin(mt, b, y0, x0);
      const int n0 = n_blk * BN + static_cast<int>(rank) * S::kBRows;
      for (int tap = 0; tap < (HALO ? 3 : P.taps); ++tap) {
        const int pad = P.stride == 1 ? 1 : 0;
        const int dy = HALO ? tap - 1 : (P.taps == 9 ? tap / 3 - pad : 0);
        const int dx = HALO ? -1 : (P.taps == 9 ? tap % 3 - pad : 0);  // HALO: the box starts one pixel to the left
        for (int kc = 0; kc < P.kchunks; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * S::kStage;
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kTxBytes);
            tma_load_4d_2sm(sa, &P.tmap_a, &full_bar[stage], kc * 64, x0 * P.stride + dx, y0 * P.stride + dy, b);
            if constexpr (HALO) {
#pragma unroll
              for (int t = 0; t < 3; ++t)
                tma_load_2d_2sm(sa + S::kA + t * S::kBTile, &P.tmap_b, &full_bar[stage], ((tap * 3 + t) * P.kchunks + kc) * 64, n0,
                                kEvictLast);
            } else {
              tma_load_2d_2sm(sa + S::kA, &P.tmap_b, &full_bar[stage], (tap * P.kchunks + kc) * 64, n0, kEvictLast);
            }
          }
          __syncwarp();
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ---- MMA issuer (leader CTA) ----
      constexpr uint32_t idesc = make_idesc(kFmtBF16, kFmtBF16, 256, BN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const uint64_t a_desc0 = make_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t b_desc0 = make_desc_sw128(smem_u32(smem) + S::kA, 16, 1024);
      for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < kiters; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t ad = desc_advance(a_desc0, stage * S::kStage);
          const uint64_t bd = desc_advance(b_desc0, stage * S::kStage);
          if (elect_one()) {
            if constexpr (HALO) {
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                // tap dx = t - 1 reads rows [t, t + 128) of the 130-row tile: the descriptor simply starts t rows (t * 128
                // bytes) into it.  MEASURED: the 128-byte swizzle is a function of the absolute shared-memory address on
                // both sides (TMA write, MMA read), so a start that is not aligned to the 8-row atom needs NO "base offset"
                // in descriptor bits 49-51 -- with the field set to the start row's phase the results are wrong.
                const uint64_t at = desc_advance(ad, t * 128);
                const uint64_t bt = desc_advance(bd, t * S::kBTile);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma_f16_ss_2sm(d_tmem, desc_advance(at, k * 32), desc_advance(bt, k * 32), idesc, (kb | t | k) != 0 ? 1u : 0u);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_f16_ss_2sm(d_tmem, desc_advance(ad, k * 32), desc_advance(bd, k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            tc_commit_2sm(&empty_bar[stage], 3);
            if (kb == kiters - 1) tc_commit_2sm(&tfull_bar[as], 3);
          }
          __syncwarp();
          if (++stage == S::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else {
    // ---- epilogue: 8 warps, warp -> (TMEM lane group, half of the BN columns); thread = one output pixel ----
    const int lg = warp & 3;
    const int part = (warp - kConvEpiWarp0) >> 2;
    constexpr int kPartCols = BN / 2;
    const int r = lg * 32 + lane;
    const int hh = r >> P.tw_shift, ww = r & (P.tw - 1);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = tile0; tile < num_tiles; tile += tile_stride) {
      const int n_blk = tile % P.n_tiles;
      const int mt = (tile / P.n_tiles) * 2 + static_cast<int>(rank);
      int b, y0, x0;
      tile_origin(mt, b, y0, x0);
      const int y = y0 + hh, x = x0 + ww;
      const bool valid = mt < P.m_tiles && y < P.H && x < P.W;
      const int64_t pix = (static_cast<int64_t>(b) * P.H + y) * P.W + x;
      const int col0 = n_blk * BN + part * kPartCols;
      // the residual rows of this tile do not depend on the accumulator: fetch them while the MMAs are still running
      // (issued just in time they put one DRAM round trip per 8 channels in front of every store)
      u32x8 rpre[kPartCols / 16];
      const bool wide_res = P.wide && P.residual != nullptr && P.out_mode == 0 && valid;
      if (wide_res) {
        const __nv_bfloat16* res = P.residual + pix * P.ld_res + col0;
#pragma unroll
        for (int i = 0; i < kPartCols / 16; ++i)
          if (col0 + i * 16 < P.N) rpre[i] = ldg_v8(res + i * 16);
      }
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + as * BN + part * kPartCols;
      // Fused GroupNorm statistics of the STORED output: per thread 16 granules (kG adjacent channels) x {sum, sum of
      // squares} for its pixel, reduced over the warp's 32 pixels once per tile by a transposing butterfly (31 shuffles
      // for all 32 values), then added by single lanes to this warp's private fp64 accumulators.
      constexpr int kG = BN >= 256 ? 8 : 4;
      float gacc[32];
      const bool do_stats = P.gn_stats != nullptr;
      if (do_stats) {
#pragma unroll
        for (int j = 0; j < 32; ++j) gacc[j] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < kPartCols / 32; ++c) {
        uint32_t v[32];
        const int col = col0 + c * 32;
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        if (col >= P.N || !valid) continue;
        if (P.out_mode == 1) {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(P.out) + pix * P.ldo + col);
          if (col + 32 <= P.N && (P.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(P.out) & 31) == 0) {
            // full 32-byte sectors per thread and instruction (rows are ldo * 4 bytes apart: no two lanes share a sector)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t o8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o8[j] = __float_as_uint(__uint_as_float(v[8 * q + j]) * P.alpha);
              stg_v8(reinterpret_cast<float*>(dst) + 8 * q, o8[0], o8[1], o8[2], o8[3], o8[4], o8[5], o8[6], o8[7]);
            }
          } else if (col + 32 <= P.N) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(v[4 * q]) * P.alpha, __uint_as_float(v[4 * q + 1]) * P.alpha,
                                   __uint_as_float(v[4 * q + 2]) * P.alpha, __uint_as_float(v[4 * q + 3]) * P.alpha);
          } else {
            float* d1 = reinterpret_cast<float*>(dst);
            for (int j = 0; j < 32 && col + j < P.N; ++j) d1[j] = __uint_as_float(v[j]) * P.alpha;
          }
        } else if (P.out_mode == 0) {
          // h = bf16(bf16(acc) + bias)  (cuDNN convolution, then the bias add of at::_convolution), out = bf16(residual + h)
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(P.out) + pix * P.ldo + col;
          const __nv_bfloat16* res = P.residual ? P.residual + pix * P.ld_res + col : nullptr;
          uint32_t ow[16];  // the chunk's 32 outputs, packed
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (col + q * 8 >= P.N) break;
            float h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = bf16r(__uint_as_float(v[q * 8 + j]));
            if (P.bias) {
              const uint4 bv = __ldg(reinterpret_cast<const uint4*>(P.bias + col + q * 8));
              const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 bf = unpack_bf16x2(bw[j]);
                h[2 * j] = bf16r(h[2 * j] + bf.x);
                h[2 * j + 1] = bf16r(h[2 * j + 1] + bf.y);
              }
            }
            if (res) {
              uint32_t rw[4];
              if (wide_res) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rw[j] = rpre[c * 2 + (q >> 1)].v[(q & 1) * 4 + j];
              } else {
                const uint4 rv = *reinterpret_cast<const uint4*>(res + q * 8);
                rw[0] = rv.x, rw[1] = rv.y, rw[2] = rv.z, rw[3] = rv.w;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 rf = unpack_bf16x2(rw[j]);
                h[2 * j] += rf.x;
                h[2 * j + 1] += rf.y;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) ow[q * 4 + j] = pack_bf16x2(h[2 * j], h[2 * j + 1]);
            if (P.wide) {  // 32-byte stores: one full sector per thread and instruction
              if (q & 1)
                stg_v8(dst + (q - 1) * 8, ow[q * 4 - 4], ow[q * 4 - 3], ow[q * 4 - 2], ow[q * 4 - 1], ow[q * 4], ow[q * 4 + 1],
                       ow[q * 4 + 2], ow[q * 4 + 3]);
            } else {
              *reinterpret_cast<uint4*>(dst + q * 8) = make_uint4(ow[q * 4], ow[q * 4 + 1], ow[q * 4 + 2], ow[q * 4 + 3]);
            }
            if (do_stats) {
              const float2 f0 = unpack_bf16x2(ow[q * 4]), f1 = unpack_bf16x2(ow[q * 4 + 1]);
              const float2 f2 = unpack_bf16x2(ow[q * 4 + 2]), f3 = unpack_bf16x2(ow[q * 4 + 3]);
              const float slo = (f0.x + f0.y) + (f1.x + f1.y), shi = (f2.x + f2.y) + (f3.x + f3.y);
              const float qlo = (f0.x * f0.x + f0.y * f0.y) + (f1.x * f1.x + f1.y * f1.y);
              const float qhi = (f2.x * f2.x + f2.y * f2.y) + (f3.x * f3.x + f3.y * f3.y);
              if constexpr (kG == 8) {
                gacc[2 * (c * 4 + q)] += slo + shi;
                gacc[2 * (c * 4 + q) + 1] += qlo + qhi;
              } else if constexpr (kPartCols / 4 <= 16) {
                gacc[2 * (c * 8 + q * 2)] += slo;
                gacc[2 * (c * 8 + q * 2) + 1] += qlo;
                gacc[2 * (c * 8 + q * 2 + 1)] += shi;
                gacc[2 * (c * 8 + q * 2 + 1) + 1] += qhi;
              }
            }
          }
        } else {
          // NCHW: out[(b * N + n) * ldo + y * W + x]; a warp's lanes are adjacent pixels of one row
          const int64_t plane = P.ldo;  // channel stride (>= H * W)
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(P.out) + static_cast<int64_t>(b) * P.N * plane +
                               static_cast<int64_t>(y) * P.W + x;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (col + j >= P.N) break;
            float h = bf16r(__uint_as_float(v[j]));
            if (P.bias) h = h + __bfloat162float(P.bias[col + j]);
            dst[(col + j) * plane] = __float2bfloat16(h);
          }
        }
      }
      if (do_stats) {
        __syncwarp();
#pragma unroll
        for (int off = 16, cnt = 16; off >= 1; off >>= 1, cnt >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int j = 0; j < cnt; ++j) {
            const float send = upper ? gacc[j] : gacc[j + cnt];
            const float keep = upper ? gacc[j + cnt] : gacc[j];
            gacc[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        // lane L now owns value L = granule (L >> 1), {sum, squares} (L & 1), summed over the warp's pixels
        float val = gacc[0];
        const int ratio = P.gn_cg / kG;  // granules per group: 1, 2 or 4 ...
        for (int o = 1; o < ratio; o <<= 1) val += __shfl_xor_sync(0xffffffffu, val, o * 2);
        const int granule = lane >> 1;
        const int gcol = col0 + granule * kG;
        if ((granule & (ratio - 1)) == 0 && granule * kG < kPartCols && gcol < P.N) {
          const int bslot = b < kConvStatsBatch ? b : kConvStatsBatch - 1;
          gn_acc[warp - kConvEpiWarp0][bslot][gcol / P.gn_cg][lane & 1] += static_cast<double>(val);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote_relaxed(&tempty_bar[as], 0);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (P.gn_stats != nullptr) {
      named_bar_sync(1, kEpiWarps * 32);  // every epilogue warp of this CTA has added its last tile
      const int nb = P.B < kConvStatsBatch ? P.B : kConvStatsBatch;
      for (int i = threadIdx.x - kConvEpiWarp0 * 32; i < nb * 64; i += kEpiWarps * 32) {
        double v = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += (&gn_acc[w8][0][0][0])[i];
        if (v != 0.0) atomicAdd(P.gn_stats + i, v);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 2 * BN);
  }
}

template <int BN, bool HALO = false>
static int launch_conv(const ConvParams& P, cudaStream_t stream) {
  using S = ConvSmem<BN, HALO>;
  static_assert(S::kTotal + 18 * 1024 <= 227 * 1024, "conv smem budget (dynamic + the static statistics accumulators)");
  static bool attr_set = false;
  auto kern = conv_igemm_kernel<BN, HALO>;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    attr_set = true;
  }
  const int pairs = sm_count() / 2;
  const int grid = (P.num_tiles < pairs ? P.num_tiles : pairs) * 2;
  FB_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(kConvThreads), S::kTotal, stream, 2, P));
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// GroupNorm.  x bf16 NHWC [B, HW, C]; 32 groups of C / 32 adjacent channels; statistics over (HW, C / 32) per (b, group).
// ---------------------------------------------------------------------------------------------------------------------
// stats[b][g] = {sum, sum of squares} (fp64, zeroed by the caller).  Thread = one 8-channel vector of a pixel.  The block's
// partial sums are combined in a FIXED order (no shared-memory float atomics: their arrival order would make the result,
// and with it the whole decode, differ from run to run in the last bit).
__global__ void __launch_bounds__(256) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, double* __restrict__ stats,
                                                       int64_t HW, int C, int pix_per_block) {
  __shared__ float part[256][8];  // [thread][bf16 pair j][sum, squares]
  const int b = blockIdx.y;
  const int vec_per_pix = C >> 3;
  const int cg = C >> 5;  // channels per group: 2, 4, 8, 16 ...
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < HW ? p0 + pix_per_block : HW;
  const int vc = threadIdx.x % vec_per_pix;       // which 8-channel vector (256 % vec_per_pix == 0)
  const int prow = threadIdx.x / vec_per_pix;
  const int rows = 256 / vec_per_pix;
  float sp[4] = {0.f, 0.f, 0.f, 0.f}, qp[4] = {0.f, 0.f, 0.f, 0.f};  // per bf16 pair: cg may be as small as 2
  const __nv_bfloat16* xb = x + static_cast<int64_t>(b) * HW * C;
  for (int64_t p = p0 + prow; p < p1; p += rows) {
    const uint4 v = *reinterpret_cast<const uint4*>(xb + p * C + vc * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(w[j]);
      sp[j] += f.x + f.y;
      qp[j] += f.x * f.x + f.y * f.y;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) part[threadIdx.x][2 * j] = sp[j], part[threadIdx.x][2 * j + 1] = qp[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    // group g = channels [g cg, (g+1) cg) = pairs [g cg/2, (g+1) cg/2); pair p lives in vector p / 4, slot p % 4, of every
    // pixel row r of the block: rows * cg / 2 = 32 terms per group and statistic, summed in index order
    const int g = threadIdx.x >> 1, st = threadIdx.x & 1;
    const int pairs = cg >> 1;
    float acc = 0.f;
    for (int r = 0; r < rows; ++r)
      for (int pp = g * pairs; pp < (g + 1) * pairs; ++pp) acc += part[r * vec_per_pix + (pp >> 2)][2 * (pp & 3) + st];
    atomicAdd(&stats[(static_cast<int64_t>(b) * 32 + g) * 2 + st], static_cast<double>(acc));
  }
}

// x * sigmoid(x) with ONE transcendental: e = 2^(-x log2 e) on the MUFU, 1 / (1 + e) by three Newton steps on the FMA pipe
// from the integer-subtraction seed (relative error 0.12 -> 1.5e-2 -> 2e-4 -> 5e-8).  The normalisation pass is otherwise
// MUFU-bound: two MUFU operations per element (ex2 + rcp) at 16 per clock and SM take longer than streaming the tensor.
__device__ __forceinline__ float swish_f32(float x) {
  const float e = exp2f_approx(fminf(-1.4426950408889634f * x, 120.f));  // clamp: 1 + 2^120 stays finite
  const float d = 1.f + e;
  float r = __int_as_float(0x7EF311C7 - __float_as_int(d));
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  r = r * fmaf(-d, r, 2.f);
  return x * r;
}

// y = bf16( f( x * a_c + b_c ) ), a_c = rstd * gamma_c, b_c = beta_c - mean * a_c (the form torch's CUDA GroupNorm
// evaluates), f = swish (x * sigmoid(x)) or identity.  grid (pixel blocks, B); a thread owns one 8-channel vector
// position and walks the block's pixels, so the per-channel affine lives in 16 registers.
__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const double* __restrict__ stats,
                                                       const __nv_bfloat16* __restrict__ gamma,
                                                       const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                       int64_t HW, int C, float eps, int swish, int pix_per_block) {
  __shared__ float sa[2048], sb[2048];
  const int b = blockIdx.y;
  const int vec_per_pix = C >> 3;
  const int cg = C >> 5;
  const double inv_n = 1.0 / (static_cast<double>(HW) * cg);
  for (int c = threadIdx.x; c < C; c += 256) {
    const double* st = stats + (static_cast<int64_t>(b) * 32 + c / cg) * 2;
    const double mean_d = st[0] * inv_n;
    const double var_d = fmax(st[1] * inv_n - mean_d * mean_d, 0.0);
    const float rstd = static_cast<float>(1.0 / sqrt(var_d + static_cast<double>(eps)));
    const float a = rstd * __bfloat162float(gamma[c]);
    sa[c] = a;
    sb[c] = __bfloat162float(beta[c]) - static_cast<float>(mean_d) * a;
  }
  __syncthreads();
  const int vc = threadIdx.x % vec_per_pix;
  const int prow = threadIdx.x / vec_per_pix;
  const int rows = 256 / vec_per_pix;
  float ca[8], cb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ca[j] = sa[vc * 8 + j], cb[j] = sb[vc * 8 + j];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < HW ? p0 + pix_per_block : HW;
  const __nv_bfloat16* xb = x + static_cast<int64_t>(b) * HW * C + vc * 8;
  __nv_bfloat16* yb = y + static_cast<int64_t>(b) * HW * C + vc * 8;
  // four independent 16-byte loads in flight per thread before any arithmetic (left to the compiler, the swish variant
  // of this loop interleaves load - ~120 instructions - store and keeps barely one load in flight)
  constexpr int kU = 4;
  for (int64_t p = p0 + prow; p < p1; p += kU * rows) {
    uint4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (p + u * rows < p1) v[u] = __ldcs(reinterpret_cast<const uint4*>(xb + (p + u * rows) * C));
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      if (p + u * rows >= p1) break;
      const uint32_t xw[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      uint32_t ow[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xf = unpack_bf16x2(xw[j]);
        float o0 = fmaf(xf.x, ca[2 * j], cb[2 * j]);
        float o1 = fmaf(xf.y, ca[2 * j + 1], cb[2 * j + 1]);
        if (swish) {
          o0 = swish_f32(o0);
          o1 = swish_f32(o1);
        }
        ow[j] = pack_bf16x2(o0, o1);
      }
      *reinterpret_cast<uint4*>(yb + (p + u * rows) * C) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
  }
}

// nearest-neighbour 2x upsampling (F.interpolate(scale_factor=2.0, mode="nearest"), autoencoder.py:121), NHWC, 16-byte vectors
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H, int W,
                                                         int vec_per_pix, int64_t total_out_vec) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < total_out_vec; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int vc = static_cast<int>(i % vec_per_pix);
    int64_t p = i / vec_per_pix;
    const int ox = static_cast<int>(p % (2 * W));
    p /= 2 * W;
    const int oy = static_cast<int>(p % (2 * H));
    const int b = static_cast<int>(p / (2 * H));
    y[i] = x[((static_cast<int64_t>(b) * H + (oy >> 1)) * W + (ox >> 1)) * vec_per_pix + vc];
  }
}

// P[r, :] = bf16( softmax(s[r, :]) ), s fp32 (already scaled); one CTA per row
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int n,
                                                           int64_t lds, int64_t ldp) {
  __shared__ float red[8];
  __shared__ float bc;
  const float* row = s + static_cast<int64_t>(blockIdx.x) * lds;
  __nv_bfloat16* out = p + static_cast<int64_t>(blockIdx.x) * ldp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x * 4; i < n; i += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = red[0];
    for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]);
    bc = t;
  }
  __syncthreads();
  m = bc;
  constexpr float kLog2e = 1.4426950408889634f;
  float sum = 0.f;
  for (int i = threadIdx.x * 4; i < n; i += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    sum += exp2f((v.x - m) * kLog2e) + exp2f((v.y - m) * kLog2e) + exp2f((v.z - m) * kLog2e) + exp2f((v.w - m) * kLog2e);
  }
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    bc = 1.f / t;
  }
  __syncthreads();
  const float inv = bc;
  for (int i = threadIdx.x * 4; i < n; i += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    uint2 o;
    o.x = pack_bf16x2(exp2f((v.x - m) * kLog2e) * inv, exp2f((v.y - m) * kLog2e) * inv);
    o.y = pack_bf16x2(exp2f((v.z - m) * kLog2e) * inv, exp2f((v.w - m) * kLog2e) * inv);
    *reinterpret_cast<uint2*>(out + i) = o;
  }
}

// z fp32 NCHW [B, C, H, W] -> bf16 NHWC [B, H, W, Cpad] with x = z / scale_factor + shift_factor (autoencoder.py:331),
// channels C..Cpad-1 zero
__global__ void __launch_bounds__(256) latent_prep_kernel(const float* __restrict__ z, __nv_bfloat16* __restrict__ y, int B, int C,
                                                          int64_t HW, int Cpad, float inv_scale, float shift_factor) {
  const int64_t total = static_cast<int64_t>(B) * HW * Cpad;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % Cpad);
    const int64_t p = i / Cpad;
    const int64_t b = p / HW, hw = p % HW;
    float v = 0.f;
    // torch's CUDA division by a host scalar multiplies by the fp32 reciprocal; the add is a separate, separately rounded op
    if (c < C) v = __fadd_rn(__fmul_rn(z[(b * C + c) * HW + hw], inv_scale), shift_factor);
    y[i] = __float2bfloat16(v);
  }
}

}  // namespace fb

extern "C" {

int fluxb200_conv2d_nhwc(const fluxb200_conv_args* a, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(a != nullptr && a->x && a->w && a->out, "fluxb200_conv2d_nhwc: null operand");
  FB_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->N > 0, "fluxb200_conv2d_nhwc: bad shape");
  FB_REQUIRE(a->Cin > 0 && a->Cin % 64 == 0, "fluxb200_conv2d_nhwc: Cin=%d must be a multiple of 64 (zero-pad it)", a->Cin);
  FB_REQUIRE(a->taps == 1 || a->taps == 9, "fluxb200_conv2d_nhwc: taps must be 1 (1x1) or 9 (3x3, padding 1)");
  FB_REQUIRE(a->out_mode >= 0 && a->out_mode <= 2, "fluxb200_conv2d_nhwc: unknown out_mode %d", a->out_mode);
  if (a->out_mode == 0) {
    FB_REQUIRE(a->N % 8 == 0 && a->ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
               "fluxb200_conv2d_nhwc(NHWC): N, ldo multiples of 8 and a 16-byte aligned output required");
    FB_REQUIRE(a->bias == nullptr || (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0, "fluxb200_conv2d_nhwc: bias alignment");
    FB_REQUIRE(a->residual == nullptr || (a->ld_res % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0),
               "fluxb200_conv2d_nhwc: residual alignment");
  } else if (a->out_mode == 1) {
    FB_REQUIRE(a->ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 15) == 0, "fluxb200_conv2d_nhwc(fp32): ldo %% 4, 16-byte aligned");
    FB_REQUIRE(a->bias == nullptr && a->residual == nullptr, "fluxb200_conv2d_nhwc(fp32): no bias / residual in this mode");
  } else {
    FB_REQUIRE(a->residual == nullptr, "fluxb200_conv2d_nhwc(NCHW): no residual in this mode");
    FB_REQUIRE(a->stride != 2 && a->ldo >= static_cast<int64_t>(a->H) * a->W,
               "fluxb200_conv2d_nhwc(NCHW): stride 1 only; ldo (channel stride) must be >= H*W");
  }
  const int stride = a->stride == 2 ? 2 : 1;
  FB_REQUIRE(a->stride == 0 || a->stride == 1 || (a->stride == 2 && a->taps == 9 && a->H >= 2 && a->W >= 2),
             "fluxb200_conv2d_nhwc: stride must be 1, or 2 with a 3x3 kernel (Downsample)");
  // Downsample (autoencoder.py:97-110): pad (0, 1, 0, 1) then 3x3 stride 2 padding 0 -> floor((H - 2) / 2) + 1 rows
  const int Ho = stride == 1 ? a->H : (a->H - 2) / 2 + 1, Wo = stride == 1 ? a->W : (a->W - 2) / 2 + 1;
  ConvParams P{};
  P.B = a->B, P.H = Ho, P.W = Wo, P.Cin = a->Cin, P.N = a->N, P.taps = a->taps, P.stride = stride;
  // pixel tile th x tw = 128 pixels: the power-of-two strip width that wastes the fewest out-of-image pixels
  // (ties -> the wider strip: longer contiguous runs per TMA box row)
  int tw = 128;
  int64_t best = -1;
  for (int t = 128; t >= 8; t /= 2) {
    const int h = 128 / t;
    const int64_t area = static_cast<int64_t>((Wo + t - 1) / t) * t * ((Ho + h - 1) / h) * h;
    if (best < 0 || area < best) best = area, tw = t;
  }
  P.tw = tw, P.th = 128 / tw;
  P.tw_shift = 0;
  while ((1 << P.tw_shift) < tw) ++P.tw_shift;
  P.tiles_x = (Wo + P.tw - 1) / P.tw;
  P.tiles_y = (Ho + P.th - 1) / P.th;
  P.m_tiles = a->B * P.tiles_x * P.tiles_y;
  int bn = a->N > 128 ? 256 : (a->N > 64 ? 128 : 64);
  // a launch that cannot even fill the machine once with 256-wide tiles (dense layers over a few hundred rows: the text
  // encoders' o / wo projections) takes 128-wide tiles: twice the CTAs, same bytes per flop from L2
  // ... and, for the few-wave launches of the dense layers (M of a few hundred rows), whichever of 256 / 128 leaves the
  // smaller idle tail in its last wave (e.g. T5's q | k | v projection: 96 tiles = 1.3 waves of 74 pairs at 256, 192 tiles
  // = 2.6 waves at 128)
  if (bn == 256 && a->N % 128 == 0 && a->taps == 1) {
    const int64_t pairs = sm_count() / 2, mp = (P.m_tiles + 1) / 2;
    const int64_t t256 = mp * ((a->N + 255) / 256), t128 = mp * (a->N / 128);
    const double e256 = static_cast<double>(t256) / (((t256 + pairs - 1) / pairs) * pairs);
    const double e128 = static_cast<double>(t128) / (((t128 + pairs - 1) / pairs) * pairs);
    if (t256 < pairs || (t256 < 4 * pairs && e128 > e256 + 0.1)) bn = 128;
  }
  P.n_tiles = (a->N + bn - 1) / bn;
  P.num_tiles = ((P.m_tiles + 1) / 2) * P.n_tiles;
  P.kchunks = a->Cin / 64;
  P.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
  P.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
  P.ld_res = a->ld_res;
  P.out = a->out, P.ldo = a->ldo, P.out_mode = a->out_mode, P.alpha = a->alpha;
  P.wide = a->out_mode == 0 && a->N % 16 == 0 && a->ldo % 16 == 0 && (reinterpret_cast<uintptr_t>(a->out) & 31) == 0 &&
           (a->residual == nullptr || (a->ld_res % 16 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 31) == 0));
  P.gn_stats = a->gn_stats;
  P.gn_cg = a->N / 32;
  if (a->gn_stats != nullptr) {
    FB_REQUIRE(a->out_mode == 0 && (a->N == 128 || a->N == 256 || a->N == 512 || a->N == 1024) &&
                   a->B <= kConvStatsBatch,
               "fluxb200_conv2d_nhwc: fused GroupNorm statistics need out_mode 0, N in {128, 256, 512, 1024}, B <= %d",
               kConvStatsBatch);
    FB_CUDA_OK(cudaMemsetAsync(a->gn_stats, 0, sizeof(double) * 64 * a->B, stream));
  }
  const int64_t ldx = a->ldx > 0 ? a->ldx : a->Cin;  // channel stride of a pixel (elements)
  FB_REQUIRE(ldx % 8 == 0, "fluxb200_conv2d_nhwc: ldx must be a multiple of 8");
  const uint64_t dims[4] = {static_cast<uint64_t>(a->Cin), static_cast<uint64_t>(a->W), static_cast<uint64_t>(a->H),
                            static_cast<uint64_t>(a->B)};
  const uint64_t strides[3] = {static_cast<uint64_t>(ldx) * 2, static_cast<uint64_t>(ldx) * 2 * a->W,
                               static_cast<uint64_t>(ldx) * 2 * a->W * a->H};
  // 3x3 on 128-pixel row strips: the halo form (one 130-pixel row box serves the three horizontal taps)
  static const bool halo_on = [] { const char* e = getenv("FLUXB200_CONV_HALO"); return e == nullptr || atoi(e) != 0; }();
  const bool halo = halo_on && a->taps == 9 && P.tw == 128 && stride == 1;
  // stride 2: the box spans 2 tw x 2 th input pixels and the TMA unit traverses it with element stride 2 in W and H, so the
  // tile that lands in shared memory is again th x tw pixels x 64 channels (out-of-bounds pixels, here the one-pixel padding
  // on the right / bottom, still read as zero)
  const uint32_t box[4] = {64, static_cast<uint32_t>((halo ? kHaloRows : P.tw) * stride), static_cast<uint32_t>(P.th * stride), 1};
  const uint32_t estr[4] = {1, static_cast<uint32_t>(stride), static_cast<uint32_t>(stride), 1};
  int rc = make_tmap_4d(&P.tmap_a, a->x, 2, dims, strides, box, estr);
  if (rc) return rc;
  const int64_t K = static_cast<int64_t>(a->taps) * a->Cin;
  const int64_t ldw = a->ldw > 0 ? a->ldw : K;
  rc = make_tmap_2d(&P.tmap_b, a->w, 2, a->N, K, ldw * 2, bn / 2, 64);
  if (rc) return rc;
  if (halo)
    return bn == 256 ? launch_conv<256, true>(P, stream) : (bn == 128 ? launch_conv<128, true>(P, stream) : launch_conv<64, true>(P, stream));
  if (bn == 256) return launch_conv<256>(P, stream);
  if (bn == 128) return launch_conv<128>(P, stream);
  return launch_conv<64>(P, stream);
}

int fluxb200_group_norm_nhwc(const void* x_bf16, const void* gamma_bf16, const void* beta_bf16, void* y_bf16, double* stats_ws,
                             int stats_ready, int B, int64_t HW, int C, float eps, int swish, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x_bf16 && gamma_bf16 && beta_bf16 && y_bf16 && stats_ws, "fluxb200_group_norm_nhwc: null operand");
  FB_REQUIRE(B > 0 && HW > 0 && C >= 64 && C % 64 == 0 && 2048 % C == 0, "fluxb200_group_norm_nhwc: C=%d must be 64 .. 2048, a power-of-two multiple of 64", C);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  if (!stats_ready) {
    FB_CUDA_OK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 64 * B, stream));
    const int spb = 512;
    dim3 sgrid(static_cast<unsigned>((HW + spb - 1) / spb), B);
    gn_stats_kernel<<<sgrid, 256, 0, stream>>>(xp, stats_ws, HW, C, spb);
  }
  const int pix_per_block = HW >= (1 << 18) ? 1024 : 256;
  dim3 grid(static_cast<unsigned>((HW + pix_per_block - 1) / pix_per_block), B);
  gn_apply_kernel<<<grid, 256, 0, stream>>>(xp, stats_ws, reinterpret_cast<const __nv_bfloat16*>(gamma_bf16),
                                            reinterpret_cast<const __nv_bfloat16*>(beta_bf16),
                                            reinterpret_cast<__nv_bfloat16*>(y_bf16), HW, C, eps, swish, pix_per_block);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

int fluxb200_upsample2x_nhwc(const void* x_bf16, void* y_bf16, int B, int H, int W, int C, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x_bf16 && y_bf16 && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "fluxb200_upsample2x_nhwc: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * 4 * H * W * (C / 8);
  const int64_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want < sm_count() * 16 ? want : sm_count() * 16);
  upsample2x_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x_bf16), reinterpret_cast<uint4*>(y_bf16), B, H, W,
                                                C / 8, total);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

int fluxb200_softmax_rows(const float* scores, int64_t lds, void* p_bf16, int64_t ldp, int rows, int n, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(scores && p_bf16 && rows > 0 && n > 0 && n % 4 == 0 && lds % 4 == 0 && ldp % 4 == 0,
             "fluxb200_softmax_rows: n, lds, ldp must be multiples of 4");
  softmax_rows_kernel<<<rows, 256, 0, stream>>>(scores, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, lds, ldp);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

int fluxb200_vae_latent_prep(const float* z_nchw, void* y_nhwc_bf16, int B, int C, int64_t HW, int Cpad, float scale_factor,
                             float shift_factor, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(z_nchw && y_nhwc_bf16 && B > 0 && C > 0 && HW > 0 && Cpad >= C, "fluxb200_vae_latent_prep: bad arguments");
  const int64_t total = static_cast<int64_t>(B) * HW * Cpad;
  const int64_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want < sm_count() * 16 ? want : sm_count() * 16);
  latent_prep_kernel<<<blocks, 256, 0, stream>>>(z_nchw, reinterpret_cast<__nv_bfloat16*>(y_nhwc_bf16), B, C, HW, Cpad,
                                                 1.0f / scale_factor, shift_factor);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
