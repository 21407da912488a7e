// Joint image+text attention for the Flux blocks: softmax(Q K^T / sqrt(128)) V, non-causal, no mask,
// head_dim 128, bf16 operands, fp32 softmax / accumulation.  Replaces attention()'s
// F.scaled_dot_product_attention + transpose/reshape (modules/flux_model.py:41-45); RoPE is applied by the
// producer of q,k (the QKV GEMM epilogue).  The output is written straight into [B,S,H*128] and can be
// quantised for the next F8Linear (float8_quantize.py:274-276) so no bf16 copy of it ever reaches HBM.
//
// One CTA = NQ query tiles of 128 rows of one (sample, head); loop over KV tiles of 128 rows.
//   warp 0        TMA: Q once, K/V tiles through mbarrier rings (SWIZZLE_128B boxes of 64 columns)
//   warp 1        tcgen05.mma issuer (kind::f16, bf16): S = Q K^T into TMEM, O += P V into TMEM
//   warp 2        TMEM allocator
//   warps 4..     one softmax warpgroup per query tile: thread == query row, S row read with tcgen05.ld,
//                 exp2-domain online softmax with lazy rescaling of O (only when the running max grows by
//                 more than 2^8), P written back as bf16 either into TMEM over S (TS MMA) or into
//                 swizzled shared memory (SS MMA)
// TMEM columns: S slot i at i*128 (fp32), O accumulator of query tile g at 256 + g*128.
#include <cuda.h>

#include "flux_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace fb {

constexpr int kD = 128;    // head dim
constexpr int kBQ = 128;   // query rows per tile
constexpr int kBKV = 128;  // kv rows per tile
constexpr int kTileBytes = kBQ * kD * 2;       // 32 KB: [2 column-chunks][128 rows][128 B]
constexpr int kChunkBytes = kBQ * 128;         // 16 KB
constexpr float kRescaleThreshold = 8.0f;      // log2 units

template <int NQ, bool TS>
struct AttnCfg {
  static constexpr int kStages = NQ == 1 ? (TS ? 3 : 2) : (TS ? 2 : 1);
  static constexpr int kPBufs = TS ? 0 : 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kQOff + NQ * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  static constexpr int kPOff = kVOff + kStages * kTileBytes;
  static constexpr int kBarOff = kPOff + kPBufs * kTileBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static constexpr int kThreads = 128 + 128 * NQ;
};

struct AttnParams {
  CUtensorMap tmap_q, tmap_k, tmap_v;
  fluxb200_attention_args a;
  int num_kv_tiles;
  float scale_log2;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// volatile: stays between the ping-pong barriers that bracket the exp phase
__device__ __forceinline__ float fast_exp2_pinned(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int REGS>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}

template <int NQ, bool TS>
__global__ void __launch_bounds__(AttnCfg<NQ, TS>::kThreads, 1) attention_kernel(const __grid_constant__ AttnParams P) {
  using C = AttnCfg<NQ, TS>;
  constexpr int KS = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = q_full + 1;           // KS
  uint64_t* k_empty = k_full + KS;         // KS
  uint64_t* v_full = k_empty + KS;         // KS
  uint64_t* v_empty = v_full + KS;         // KS
  uint64_t* s_ready = v_empty + KS;        // 2 (per S slot)
  uint64_t* p_ready = s_ready + 2;         // NQ
  uint64_t* o_done = p_ready + NQ;         // NQ
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + NQ);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const fluxb200_attention_args& a = P.a;
  const int q0 = blockIdx.x * (NQ * kBQ);
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.H + h;
  const int n = P.num_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmap_q);
    tma_prefetch_desc(&P.tmap_k);
    tma_prefetch_desc(&P.tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(&s_ready[0], 1);
    mbar_init(&s_ready[1], 1);
    for (int g = 0; g < NQ; ++g) {
      mbar_init(&p_ready[g], 4);  // one arrive per softmax warp
      mbar_init(&o_done[g], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < 4) {
    if constexpr (NQ == 2) reg_dec<56>();
    if (warp == 0 && lane == 0) {
      // ---------------- TMA producer ----------------
      mbar_arrive_expect_tx(q_full, NQ * kTileBytes);
      for (int g = 0; g < NQ; ++g) {
        uint8_t* dst = smem + C::kQOff + g * kTileBytes;
        tma_load_3d(dst, &P.tmap_q, q_full, 0, q0 + g * kBQ, bh, kEvictFirst);
        tma_load_3d(dst + kChunkBytes, &P.tmap_q, q_full, 64, q0 + g * kBQ, bh, kEvictFirst);
      }
      for (int j = 0; j < n; ++j) {
        const int st = j % KS;
        const uint32_t ph = (j / KS) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], kTileBytes);
        uint8_t* kd = smem + C::kKOff + st * kTileBytes;
        tma_load_3d(kd, &P.tmap_k, &k_full[st], 0, j * kBKV, bh, kEvictLast);
        tma_load_3d(kd + kChunkBytes, &P.tmap_k, &k_full[st], 64, j * kBKV, bh, kEvictLast);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], kTileBytes);
        uint8_t* vd = smem + C::kVOff + st * kTileBytes;
        tma_load_3d(vd, &P.tmap_v, &v_full[st], 0, j * kBKV, bh, kEvictLast);
        tma_load_3d(vd + kChunkBytes, &P.tmap_v, &v_full[st], 64, j * kBKV, bh, kEvictLast);
      }
    } else if (warp == 1 && lane == 0) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kBQ, kD, 0, 1);  // V is MN-major
      const uint32_t q_addr = smem_u32(smem + C::kQOff);
      const uint32_t k_addr = smem_u32(smem + C::kKOff);
      const uint32_t v_addr = smem_u32(smem + C::kVOff);
      const uint32_t p_addr = smem_u32(smem + C::kPOff);

      auto issue_qk = [&](int g, int slot, int st) {
        const uint32_t d = tmem_base + slot * 128;
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
          uint64_t ad = make_desc_sw128(q_addr + g * kTileBytes + off, 16, 1024);
          uint64_t bd = make_desc_sw128(k_addr + st * kTileBytes + off, 16, 1024);
          mma_f16_ss(d, ad, bd, idesc_qk, kk != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int g, int pslot, int st, bool first) {
        const uint32_t d = tmem_base + 256 + g * 128;
#pragma unroll
        for (int kk = 0; kk < kBKV / 16; ++kk) {
          // V tile: [2 d-chunks][kv rows][128 B]; 16 kv rows per MMA = 2048 B; next d-chunk at kChunkBytes
          uint64_t bd = make_desc_sw128(v_addr + st * kTileBytes + kk * 2048, kChunkBytes, 1024);
          const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
          if constexpr (TS) {
            mma_f16_ts(d, tmem_base + pslot * 128 + kk * 8, bd, idesc_pv, acc);
          } else {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
            uint64_t ad = make_desc_sw128(p_addr + pslot * kTileBytes + off, 16, 1024);
            mma_f16_ss(d, ad, bd, idesc_pv, acc);
          }
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if constexpr (NQ == 2) {
        for (int g = 0; g < 2; ++g) {
          issue_qk(g, g, 0);
          tc_commit(&s_ready[g]);
        }
        tc_commit(&k_empty[0]);
        for (int j = 0; j < n; ++j) {
          const int st = j % KS;
          mbar_wait(&v_full[st], (j / KS) & 1);
          for (int g = 0; g < 2; ++g) {
            mbar_wait(&p_ready[g], j & 1);
            tc_fence_after();
            issue_pv(g, g, st, j == 0);
            tc_commit(&o_done[g]);
            if (g == 1) tc_commit(&v_empty[st]);
            if (j + 1 < n) {
              const int st1 = (j + 1) % KS;
              if (g == 0) {
                mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
                tc_fence_after();
              }
              issue_qk(g, g, st1);
              tc_commit(&s_ready[g]);
              if (g == 1) tc_commit(&k_empty[st1]);
            }
          }
        }
      } else {
        issue_qk(0, 0, 0);
        tc_commit(&s_ready[0]);
        tc_commit(&k_empty[0]);
        for (int j = 0; j < n; ++j) {
          const int st = j % KS;
          if (j + 1 < n) {
            const int st1 = (j + 1) % KS;
            mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
            tc_fence_after();
            issue_qk(0, (j + 1) & 1, st1);
            tc_commit(&s_ready[(j + 1) & 1]);
            tc_commit(&k_empty[st1]);
          }
          mbar_wait(&v_full[st], (j / KS) & 1);
          mbar_wait(&p_ready[0], j & 1);
          tc_fence_after();
          issue_pv(0, j & 1, st, j == 0);
          tc_commit(&o_done[0]);
          tc_commit(&v_empty[st]);
        }
      }
    }
  } else {
    // ---------------- softmax warpgroups ----------------
    if constexpr (NQ == 2) reg_inc<208>();
    const int g = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;  // row within the query tile
    const int qrow = q0 + g * kBQ + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t o_taddr = lane_base + 256 + g * 128;
    const float sl2 = P.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;
    // Ping-pong token between the two softmax warpgroups (named barriers 1 and 2): only one of them is in its
    // MUFU-bound exp phase at a time, so while one exponentiates the tensor pipe works on the other's tiles.
    // Without it both run in lock-step and the exp phases and the MMAs serialise (profiles/r1_attention.md).
    if constexpr (NQ == 2) {
      if (g == 1) named_bar_arrive(1, 256);  // hand the first turn to warpgroup 0
    }

    for (int j = 0; j < n; ++j) {
      const int slot = NQ == 2 ? g : (j & 1);
      const uint32_t sph = NQ == 2 ? (j & 1) : ((j >> 1) & 1);
      mbar_wait(&s_ready[slot], sph);
      tc_fence_after();
      uint32_t sv[128];
      {
        uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(lane_base + slot * 128 + 0, sv4[0]);
        tmem_ld32(lane_base + slot * 128 + 32, sv4[1]);
        tmem_ld32(lane_base + slot * 128 + 64, sv4[2]);
        tmem_ld32(lane_base + slot * 128 + 96, sv4[3]);
        tmem_ld_wait();
      }
      const int kv_left = a.S - j * kBKV;  // columns >= kv_left are out of range (last tile only)
      if (kv_left < kBKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      float mx = __uint_as_float(sv[0]);
#pragma unroll
      for (int i = 1; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      const float m_cand = mx * sl2;
      // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one.
      const bool grow = m_cand > m_used + kRescaleThreshold;
      const bool warp_grow = __any_sync(0xffffffffu, grow);
      float alpha = 1.f;
      if (warp_grow) {
        const float m_new = fmaxf(m_used, m_cand);
        alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
        m_used = m_new;
        l *= alpha;
      }
      float rs = 0.f;
      const float neg_m = -m_used;
      if constexpr (NQ == 2) named_bar_sync(1 + g, 256);  // wait for our turn
#pragma unroll
      for (int i = 0; i < 128; ++i) {
        float p = fast_exp2_pinned(fmaf(__uint_as_float(sv[i]), sl2, neg_m));
        rs += p;
        sv[i] = __float_as_uint(p);
      }
      if constexpr (NQ == 2) named_bar_arrive(1 + (g ^ 1), 256);  // pass the turn
      l += rs;

      if (j > 0) {
        mbar_wait(&o_done[g], (j - 1) & 1);  // PV(j-1) finished: O stable, P buffer reusable
        tc_fence_after();
        if (warp_grow) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t ov[32];
            tmem_ld32(o_taddr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(o_taddr + c * 32, ov);
          }
        }
      }
      if constexpr (TS) {
        // P (bf16 pairs) over the S slot: column c holds kv (2c, 2c+1) of this row
        uint32_t pk[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(sv[half * 64 + 2 * i]), __uint_as_float(sv[half * 64 + 2 * i + 1]));
          tmem_st32(lane_base + slot * 128 + half * 32, pk);
        }
        tmem_st_wait();
      } else {
        // P into swizzled smem: [2 kv-chunks][128 rows][128 B], 16-byte unit u of row r at u ^ (r & 7)
        uint8_t* pb = smem + C::kPOff + (NQ == 2 ? g : (j & 1)) * kTileBytes + r * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int i0 = c * 64 + u * 8;
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(sv[i0 + 0]), __uint_as_float(sv[i0 + 1]));
            o.y = pack_bf16x2(__uint_as_float(sv[i0 + 2]), __uint_as_float(sv[i0 + 3]));
            o.z = pack_bf16x2(__uint_as_float(sv[i0 + 4]), __uint_as_float(sv[i0 + 5]));
            o.w = pack_bf16x2(__uint_as_float(sv[i0 + 6]), __uint_as_float(sv[i0 + 7]));
            *reinterpret_cast<uint4*>(pb + c * kChunkBytes + ((u ^ (r & 7)) << 4)) = o;
          }
        }
        tmem_st_wait();  // the O rescale stores, if any
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);
    }

    // ---------------- epilogue: O / l -> bf16 -> (optional) fp8 ----------------
    mbar_wait(&o_done[g], (n - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const bool valid = qrow < a.S;
    const bool second = a.out1 != nullptr && qrow >= a.split_row;
    void* const outp = second ? a.out1 : a.out;
    const int64_t obase = second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                       static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h * kD
                                 : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h * kD;
    float oscale = 1.f;
    if (a.out_kind == 1) oscale = __ldg(qrow < a.split_row ? a.out_scale0 : a.out_scale1);
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t ov[32];
      tmem_ld32(o_taddr + c * 32, ov);
      tmem_ld_wait();
      if (!valid) continue;
      if (a.out_kind == 0) {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(outp) + obase + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(ov[q * 8 + 0]) * inv_l, __uint_as_float(ov[q * 8 + 1]) * inv_l);
          o.y = pack_bf16x2(__uint_as_float(ov[q * 8 + 2]) * inv_l, __uint_as_float(ov[q * 8 + 3]) * inv_l);
          o.z = pack_bf16x2(__uint_as_float(ov[q * 8 + 4]) * inv_l, __uint_as_float(ov[q * 8 + 5]) * inv_l);
          o.w = pack_bf16x2(__uint_as_float(ov[q * 8 + 6]) * inv_l, __uint_as_float(ov[q * 8 + 7]) * inv_l);
          dst[q] = o;
        }
      } else {
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(outp) + obase + c * 32);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float o = bf16r(__uint_as_float(ov[q * 16 + t * 4 + e]) * inv_l);
              f[e] = a.out_fmt == FLUXB200_E5M2 ? quant_pre<1>(o, oscale) : quant_pre<0>(o, oscale);
            }
            if (a.out_fmt == FLUXB200_E5M2)
              w[t] = to_fp8x2<1>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<1>(f[2], f[3])) << 16);
            else
              w[t] = to_fp8x2<0>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<0>(f[2], f[3])) << 16);
          }
          dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int NQ, bool TS>
static int launch_attention(const AttnParams& P, cudaStream_t stream) {
  using C = AttnCfg<NQ, TS>;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  auto kern = attention_kernel<NQ, TS>;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + NQ * kBQ - 1) / (NQ * kBQ), a.H, a.B);
  kern<<<grid, C::kThreads, C::kTotal, stream>>>(P);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace fb

extern "C" int fluxb200_attention(const fluxb200_attention_args* args, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(args != nullptr, "fluxb200_attention: args is NULL");
  const fluxb200_attention_args& a = *args;
  FB_REQUIRE(a.q && a.k && a.v && a.out, "fluxb200_attention: null operand");
  FB_REQUIRE(a.B > 0 && a.H > 0 && a.S > 0, "fluxb200_attention: bad shape B=%d H=%d S=%d", a.B, a.H, a.S);
  FB_REQUIRE(a.out_kind == 0 || a.out_kind == 1, "fluxb200_attention: bad out_kind");
  if (a.out_kind == 0)
    FB_REQUIRE(a.ldo % 8 == 0 && a.out_batch_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0,
               "fluxb200_attention: bf16 out needs 16-byte aligned rows");
  else
    FB_REQUIRE(a.ldo % 16 == 0 && a.out_batch_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 &&
                   a.out_scale0 && a.out_scale1 && (a.out_fmt == 0 || a.out_fmt == 1),
               "fluxb200_attention: fp8 out needs 16-byte aligned rows and both scales");
  FB_REQUIRE(a.ldo >= static_cast<int64_t>(a.H) * kD, "fluxb200_attention: ldo < H*128");
  if (a.out1 != nullptr)
    FB_REQUIRE(a.ldo1 >= static_cast<int64_t>(a.H) * kD && a.ldo1 % 16 == 0 && a.out1_batch_stride % 16 == 0 &&
                   (reinterpret_cast<uintptr_t>(a.out1) & 15) == 0 && a.split_row >= 0,
               "fluxb200_attention: bad second destination");

  AttnParams P;
  P.a = a;
  P.num_kv_tiles = (a.S + kBKV - 1) / kBKV;
  P.scale_log2 = a.softmax_scale * 1.4426950408889634f;
  const uint64_t bhn = static_cast<uint64_t>(a.B) * a.H;
  const uint64_t row_bytes = kD * 2;
  int rc;
  if ((rc = make_tmap_3d(&P.tmap_q, a.q, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBQ, 1))) return rc;
  if ((rc = make_tmap_3d(&P.tmap_k, a.k, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBKV, 1))) return rc;
  if ((rc = make_tmap_3d(&P.tmap_v, a.v, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBKV, 1))) return rc;

  switch (a.variant) {
    case 0:
    case 1: return launch_attention<2, true>(P, stream);   // 2 query tiles, P through TMEM
    case 2: return launch_attention<1, false>(P, stream);  // 1 query tile, P through smem (SS MMA)
    case 3: return launch_attention<1, true>(P, stream);   // 1 query tile, P through TMEM
    case 4: return launch_attention<2, false>(P, stream);  // 2 query tiles, P through smem
    default: return set_error(FLUXB200_ERR_INVALID, "fluxb200_attention: unknown variant %d", a.variant);
  }
}
