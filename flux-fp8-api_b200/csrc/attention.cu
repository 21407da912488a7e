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

#include <cstdlib>

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
  int debug;  // timing experiments only (env FLUXB200_ATTN_DEBUG): 1 = MMA/TMA pipeline alone, softmax skipped
};

// Phase timing (cycles) of CTA (0,0,0): [0..7] softmax warpgroup 0 / warp 4 lane 0, [8..15] MMA issuer.
// Read back with fluxb200_debug_counters(); negligible cost (one predicated thread per role).
__device__ unsigned long long g_attn_dbg[16];
#ifdef FLUXB200_ATTN_PROBE
// Event trace of CTA (0,0,0) of the step-interleaved kernel: [role 0 issuer, 1 / 2 softmax warpgroups][step < 24][event < 8]
__device__ unsigned long long g_attn_trace[3 * 24 * 8];
#define FB_TRACE(on, role, j, ev) \
  do { if ((on) && (j) < 24) g_attn_trace[((role) * 24 + (j)) * 8 + (ev)] = clock64(); } while (0)
#else
#define FB_TRACE(on, role, j, ev) do { } while (0)
#endif
__device__ __forceinline__ unsigned long long clk() { return clock64(); }

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

// exp2 on the FMA/ALU pipes (no MUFU): round-to-nearest range reduction x = n + f, f in [-0.5, 0.5], cubic minimax
// 2^f (max relative error 1.0e-4, a twentieth of a bf16 ulp of P), exponent patched in with an integer add.
// Used for a quarter of the elements so the MUFU-bound exp phase shortens (the trick FlashAttention-4 uses).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float r = x + 12582912.f;  // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (r - 12582912.f);
  const float p = fmaf(fmaf(fmaf(0.05500892f, f, 0.24221096f), f, 0.69328293f), f, 1.f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

template <int REGS>
__device__ __forceinline__ void reg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS));
}
template <int REGS>
__device__ __forceinline__ void reg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS));
}

// SPLIT = 64 or 96 (with FAST && TS && NQ == 2; "CHUNK"): P is handed to the MMA issuer in two pieces, the first
// SPLIT KV columns and the rest.  The PV product is
// K-split over the KV columns (MMA kk reads P[:, 16kk .. 16kk+16)), so PV over the first half runs on the tensor
// pipe while the softmax warps still exponentiate the second half: ~350 cycles of MMA and the first P store leave
// the strictly serial S -> softmax -> P -> PV -> next S chain that bounds this kernel.
// POLY (chunked path): how many of every 8 exponentials go to the FMA-pipe polynomial instead of the MUFU (2, 4 or 6).
// EARLY (chunked path): the next step's QK product of a tile is issued as two N = 64 halves.  P(j) (bf16 pairs) only
// occupies TMEM columns [0, 64) of the tile's S slot, so the half that lands in columns [64, 128) -- KV rows 64..127 of
// K(j+1) -- is issued as soon as the softmax warps have pulled S(j) into registers (barrier s_cons), i.e. it runs
// UNDER the exponentials; only the N = 64 half for columns [0, 64) still waits for PV(j) to retire.  That takes 4 of
// the 12 N = 128 MMA-equivalents out of the strictly serial S -> softmax -> P -> PV -> QK -> S chain of a tile.
template <int NQ, bool TS, bool FAST = false, int SPLIT = 0, int POLY = 2, bool EARLY = false>
__global__ void __launch_bounds__(AttnCfg<NQ, TS>::kThreads, 1) attention_kernel(const __grid_constant__ AttnParams P) {
  constexpr bool CHUNK = SPLIT != 0;
  static_assert(!EARLY || (CHUNK && NQ == 2 && TS && FAST), "EARLY is a variant of the chunked two-tile kernel");
  static_assert(SPLIT == 0 || SPLIT == 64 || SPLIT == 96, "first hand-off: 64 or 96 KV columns");
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
  uint64_t* p_lo = o_done + NQ;            // NQ (CHUNK: first half of P stored)
  uint64_t* s_cons = p_lo + NQ;            // NQ (EARLY: S(j) is in the softmax warps' registers)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_cons + NQ);

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
      mbar_init(&p_lo[g], 4);
      mbar_init(&s_cons[g], 4);
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
  pdl_wait();  // q, k, v come from the preceding QKV GEMMs

  if (warp < 4) {
    if constexpr (NQ == 2) reg_dec<88>();  // 128 x 88 + 256 x 208 = 64512 = the whole launch allocation (384 x 168)
    if (warp == 0) {
      // ---------------- TMA producer (warp-uniform loop, one elected lane issues) ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, NQ * kTileBytes);
        for (int g = 0; g < NQ; ++g) {
          uint8_t* dst = smem + C::kQOff + g * kTileBytes;
          tma_load_3d(dst, &P.tmap_q, q_full, 0, q0 + g * kBQ, bh, kEvictFirst);
          tma_load_3d(dst + kChunkBytes, &P.tmap_q, q_full, 64, q0 + g * kBQ, bh, kEvictFirst);
        }
      }
      __syncwarp();
      for (int j = 0; j < n; ++j) {
        const int st = j % KS;
        const uint32_t ph = (j / KS) & 1;
        uint8_t* kd = smem + C::kKOff + st * kTileBytes;
        uint8_t* vd = smem + C::kVOff + st * kTileBytes;
        mbar_wait(&k_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], kTileBytes);
          tma_load_3d(kd, &P.tmap_k, &k_full[st], 0, j * kBKV, bh, kEvictLast);
          tma_load_3d(kd + kChunkBytes, &P.tmap_k, &k_full[st], 64, j * kBKV, bh, kEvictLast);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], kTileBytes);
          tma_load_3d(vd, &P.tmap_v, &v_full[st], 0, j * kBKV, bh, kEvictLast);
          tma_load_3d(vd + kChunkBytes, &P.tmap_v, &v_full[st], 64, j * kBKV, bh, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---------------- MMA issuer ----------------
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kBQ, kD, 0, 1);  // V is MN-major
      const uint32_t q_addr = smem_u32(smem + C::kQOff);
      const uint32_t k_addr = smem_u32(smem + C::kKOff);
      const uint32_t v_addr = smem_u32(smem + C::kVOff);
      const uint32_t p_addr = smem_u32(smem + C::kPOff);

      // descriptors of stage / tile i are base + i * tile bytes: derived arithmetically (arrays indexed by a run-time
      // stage number live in local memory, and the issuer would fetch every descriptor through LDL)
      const uint64_t q_desc0 = make_desc_sw128(q_addr, 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(k_addr, 16, 1024);
      const uint64_t v_desc0 = make_desc_sw128(v_addr, kChunkBytes, 1024);
      const uint64_t p_desc0 = make_desc_sw128(p_addr, 16, 1024);
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
        __syncwarp();
      };
      auto issue_qk = [&](int g, int slot, int st) {
        const uint32_t d = tmem_base + slot * 128;
        const uint64_t ad0 = desc_advance(q_desc0, g * kTileBytes), bd0 = desc_advance(k_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
#ifdef FLUXB200_ATTN_EXPERIMENT_TS_QK  // timing experiment only: A operand (Q) from TMEM instead of shared memory
            mma_f16_ts(d, tmem_base + 256 + g * 128 + kk * 8, desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
#else
            mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
#endif
          }
        }
        __syncwarp();
      };
      // N = 64 half of QK: KV rows [64 half, 64 half + 64) of the K tile -> TMEM columns [64 half, +64) of the slot
      constexpr uint32_t idesc_qk64 = make_idesc(kFmtBF16, kFmtBF16, kBQ, kBKV / 2, 0, 0);
      auto issue_qk_half = [&](int g, int slot, int st, int half) {
        const uint32_t d = tmem_base + slot * 128 + half * 64;
        const uint64_t ad0 = desc_advance(q_desc0, g * kTileBytes);
        const uint64_t bd0 = desc_advance(k_desc0, st * kTileBytes + half * (kBKV / 2) * 128);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
            mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk64, kk != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int pslot, int st, bool first, int kk0 = 0, int kk1 = kBKV / 16) {
        const uint32_t d = tmem_base + 256 + g * 128;
        const uint64_t bd0 = desc_advance(v_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk) {
            // V tile: [2 d-chunks][kv rows][128 B]; 16 kv rows per MMA = 2048 B; next d-chunk at kChunkBytes
            const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
            if constexpr (TS) {
              mma_f16_ts(d, tmem_base + pslot * 128 + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, acc);
            } else {
              const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
              mma_f16_ss(d, desc_advance(desc_advance(p_desc0, pslot * kTileBytes), off), desc_advance(bd0, kk * 2048), idesc_pv, acc);
            }
          }
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      if constexpr (NQ == 2) {
        for (int g = 0; g < 2; ++g) {
          issue_qk(g, g, 0);
          commit(&s_ready[g]);
        }
        commit(&k_empty[0]);
        for (int j = 0; j < n; ++j) {
          const int st = j % KS;
          mbar_wait(&v_full[st], (j / KS) & 1);
          for (int g = 0; g < 2; ++g) {
            if constexpr (EARLY) {
              if (j + 1 < n) {
                const int st1 = (j + 1) % KS;
                if (g == 0) {
                  mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
                  tc_fence_after();
                }
                if (P.debug != 1) mbar_wait(&s_cons[g], j & 1);
                tc_fence_after();
                issue_qk_half(g, g, st1, 1);
              }
            }
            if constexpr (CHUNK) {
              if (P.debug != 1) mbar_wait(&p_lo[g], j & 1);
              tc_fence_after();
              issue_pv(g, g, st, j == 0, 0, SPLIT / 16);
              if (P.debug != 1) mbar_wait(&p_ready[g], j & 1);
              tc_fence_after();
              issue_pv(g, g, st, j == 0, SPLIT / 16, kBKV / 16);
            } else {
              if (P.debug != 1) mbar_wait(&p_ready[g], j & 1);
              tc_fence_after();
              issue_pv(g, g, st, j == 0);
            }
            commit(&o_done[g]);
            if (g == 1) commit(&v_empty[st]);
            if (j + 1 < n) {
              const int st1 = (j + 1) % KS;
              if constexpr (EARLY) {
                issue_qk_half(g, g, st1, 0);
              } else {
                if (g == 0) {
                  mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
                  tc_fence_after();
                }
                issue_qk(g, g, st1);
              }
              commit(&s_ready[g]);
              if (g == 1) commit(&k_empty[st1]);
            }
          }
        }
        if (P.debug == 1) {  // drain the tensor pipe before teardown
          commit(q_full);
          mbar_wait(q_full, 1);
        }
      } else {
        issue_qk(0, 0, 0);
        commit(&s_ready[0]);
        commit(&k_empty[0]);
        for (int j = 0; j < n; ++j) {
          const int st = j % KS;
          if (j + 1 < n) {
            const int st1 = (j + 1) % KS;
            mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
            tc_fence_after();
            issue_qk(0, (j + 1) & 1, st1);
            commit(&s_ready[(j + 1) & 1]);
            commit(&k_empty[st1]);
          }
          mbar_wait(&v_full[st], (j / KS) & 1);
          mbar_wait(&p_ready[0], j & 1);
          tc_fence_after();
          issue_pv(0, j & 1, st, j == 0);
          commit(&o_done[0]);
          commit(&v_empty[st]);
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
    if (NQ == 2 && (CHUNK || P.debug != 4)) {
      if (g == 1) named_bar_arrive(1, 256);  // hand the first turn to warpgroup 0
    }

#ifdef FLUXB200_ATTN_PROBE
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
#else
    constexpr bool dbg = false;  // phase timers compiled out (build with -DFLUXB200_ATTN_PROBE to enable)
#endif
    unsigned long long d_wait_s = 0, d_ld = 0, d_max = 0, d_exp = 0, d_wait_o = 0, d_st = 0, tA = 0, tB = 0;
    for (int j = 0; j < (P.debug == 1 ? 0 : n); ++j) {
      const int slot = NQ == 2 ? g : (j & 1);
      const uint32_t sph = NQ == 2 ? (j & 1) : ((j >> 1) & 1);
      if (dbg) tA = clk();
      if constexpr (FAST && TS) {
        // PV(j-1) was issued before QK(j): by the time S(j) can be ready this wait has long been satisfied, so taking
        // it here (instead of between the exponentials and the P store) removes a barrier poll from the S -> P chain
        if (j > 0) mbar_wait(&o_done[g], (j - 1) & 1);  // O stable, P buffer reusable
      }
      mbar_wait(&s_ready[slot], sph);
      tc_fence_after();
      if (dbg) { tB = clk(); d_wait_s += tB - tA; tA = tB; }
      uint32_t sv[128];
      {
        uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(lane_base + slot * 128 + 0, sv4[0]);
        tmem_ld32(lane_base + slot * 128 + 32, sv4[1]);
        tmem_ld32(lane_base + slot * 128 + 64, sv4[2]);
        tmem_ld32(lane_base + slot * 128 + 96, sv4[3]);
        tmem_ld_wait();
      }
      if constexpr (EARLY) {
        // S(j) is in registers: columns [64, 128) of the slot may receive half of S(j+1) while we exponentiate
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_cons[g]);
      }
      if (dbg) { tB = clk(); d_ld += tB - tA; tA = tB; }
      const int kv_left = a.S - j * kBKV;  // columns >= kv_left are out of range (last tile only)
      if (kv_left < kBKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      float alpha = 1.f;
      bool warp_grow = false;
      float rs = 0.f;
      if constexpr (FAST && TS && CHUNK) {
        // classic order (max -> lazy rescale -> exponentials), the exponentials in two halves with a P hand-off each
        float m0 = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
        float m1 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
        float m2 = fmaxf(__uint_as_float(sv[6]), __uint_as_float(sv[7]));
        float m3 = -INFINITY;
#pragma unroll
        for (int i = 8; i < 128; i += 8) {
          m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
          m2 = fmax3(m2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
          m3 = fmax3(m3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
        }
        const float m_cand = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
        const bool grow = m_cand > m_used + kRescaleThreshold;
        warp_grow = __any_sync(0xffffffffu, grow);
        if (warp_grow) {
          const float m_new = fmaxf(m_used, m_cand);
          alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
          m_used = m_new;
          l *= alpha;
          if (j > 0) {  // o_done(j-1) was awaited at the top of the step
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t ov[32];
              tmem_ld32(o_taddr + c * 32, ov);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
              tmem_st32(o_taddr + c * 32, ov);
            }
            tmem_st_wait();
          }
        }
        if (dbg) { tB = clk(); d_max += tB - tA; tA = tB; }
        // UNCONDITIONAL on purpose: ptxas sinks a predicated BAR.SYNC below the first ~67 exponentials of the pass (it moves
        // MUFU freely across barriers), the two warpgroups' passes then overlap by half, and that costs 1-3.5 %
        // (profiles/r2_attention.md, experiment 5)
        named_bar_sync(1 + g, 256);  // wait for our turn
        const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(-m_used, -m_used);
        const float2 magic = make_float2(12582912.f, 12582912.f);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
          for (int i = hh * SPLIT; i < (hh == 0 ? SPLIT : 128); i += 8) {
            const float2 t01 = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, negmv);
            const float2 t23 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, negmv);
            const float2 t45 = ffma2(make_float2(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])), sl2v, negmv);
            const float2 t67 = ffma2(make_float2(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])), sl2v, negmv);
            // 2^x for a pair on the FMA pipe (poly_exp2 on packed fp32)
            auto poly2 = [&](float2 tt) {
              const float2 x = make_float2(fmaxf(tt.x, -126.f), fmaxf(tt.y, -126.f));
              const float2 rr = fadd2(x, magic);
              const float2 f = fsub2(x, fsub2(rr, magic));
              float2 pp = ffma2(make_float2(0.05500892f, 0.05500892f), f, make_float2(0.24221096f, 0.24221096f));
              pp = ffma2(pp, f, make_float2(0.69328293f, 0.69328293f));
              pp = ffma2(pp, f, make_float2(1.f, 1.f));
              return make_float2(__int_as_float(__float_as_int(pp.x) + (__float_as_int(rr.x) << 23)),
                                 __int_as_float(__float_as_int(pp.y) + (__float_as_int(rr.y) << 23)));
            };
            float p0, p1, p2, p3, p4, p5, p6, p7;
            p0 = fast_exp2_pinned(t01.x), p1 = fast_exp2_pinned(t01.y);
            if constexpr (POLY == 0) {
              p2 = fast_exp2_pinned(t23.x), p3 = fast_exp2_pinned(t23.y), p4 = fast_exp2_pinned(t45.x);
              p5 = fast_exp2_pinned(t45.y), p6 = fast_exp2_pinned(t67.x), p7 = fast_exp2_pinned(t67.y);
            } else if constexpr (POLY == 2) {
              p2 = fast_exp2_pinned(t23.x), p4 = fast_exp2_pinned(t45.x), p5 = fast_exp2_pinned(t45.y);
              p6 = fast_exp2_pinned(t67.x);
              const float2 q = poly2(make_float2(t23.y, t67.y));
              p3 = q.x, p7 = q.y;
            } else if constexpr (POLY == 4) {
              p4 = fast_exp2_pinned(t45.x), p5 = fast_exp2_pinned(t45.y);
              const float2 q0 = poly2(t23), q1 = poly2(t67);
              p2 = q0.x, p3 = q0.y, p6 = q1.x, p7 = q1.y;
            } else {
              const float2 q0 = poly2(t23), q1 = poly2(t45), q2 = poly2(t67);
              p2 = q0.x, p3 = q0.y, p4 = q1.x, p5 = q1.y, p6 = q2.x, p7 = q2.y;
            }
            acc0 = fadd2(acc0, make_float2(p0, p1));
            acc1 = fadd2(acc1, make_float2(p2, p3));
            acc0 = fadd2(acc0, make_float2(p4, p5));
            acc1 = fadd2(acc1, make_float2(p6, p7));
            sv[i >> 1] = pack_bf16x2(p0, p1);
            sv[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            sv[(i >> 1) + 2] = pack_bf16x2(p4, p5);
            sv[(i >> 1) + 3] = pack_bf16x2(p6, p7);
          }
          // hand this piece of P (packed: 2 KV columns per TMEM column) to the issuer
          uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
          if constexpr (SPLIT == 64) {
            tmem_st32(lane_base + slot * 128 + hh * 32, pk[hh]);
          } else if (hh == 0) {
            tmem_st32(lane_base + slot * 128, pk[0]);
            tmem_st16(lane_base + slot * 128 + 32, sv + 32);
          } else {
            tmem_st16(lane_base + slot * 128 + 48, sv + 48);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(hh == 0 ? &p_lo[g] : &p_ready[g]);
        }
        named_bar_arrive(1 + (g ^ 1), 256);  // pass the turn
        const float2 acc = fadd2(acc0, acc1);
        l += acc.x + acc.y;
        if (dbg) { tB = clk(); d_exp += tB - tA; tA = tB; }
        continue;
      } else if constexpr (FAST && TS) {
        // One pass: 3 of 4 exponentials on the MUFU, 1 of 4 on the FMA pipe; bf16 pairs packed in place.  The fp32 work
        // runs on packed pairs (FFMA2 / FADD2: two IEEE operations per issue slot).
        // The pass exponentiates against the STALE running max and tracks the maximum exponent it met; only if
        // some row of the warp ran more than 2^kRescaleThreshold above its stale max (or on the first tile, which
        // has none) is the tile redone the classic way: max first, rescale, exponentiate.  The separate max pass
        // was 9 % of the per-tile S -> P -> PV -> S chain that bounds the kernel (profiles/r1_attention_variants.md).
        auto exp_pass = [&](float neg_m, float& tmax_out) {
          const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(neg_m, neg_m);
          const float2 magic = make_float2(12582912.f, 12582912.f);
          float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
          float tm0 = -INFINITY, tm1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            const float2 t01 = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, negmv);
            const float2 t23 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, negmv);
            const float2 t45 = ffma2(make_float2(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])), sl2v, negmv);
            const float2 t67 = ffma2(make_float2(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])), sl2v, negmv);
            tm0 = fmaxf(tm0, fmaxf(t01.x, t01.y));
            tm1 = fmaxf(tm1, fmaxf(t23.x, t23.y));
            tm0 = fmaxf(tm0, fmaxf(t45.x, t45.y));
            tm1 = fmaxf(tm1, fmaxf(t67.x, t67.y));
            const float p0 = fast_exp2_pinned(t01.x), p1 = fast_exp2_pinned(t01.y), p2 = fast_exp2_pinned(t23.x);
            const float p4 = fast_exp2_pinned(t45.x), p5 = fast_exp2_pinned(t45.y), p6 = fast_exp2_pinned(t67.x);
            // poly_exp2 on the pair (t23.y, t67.y)
            const float2 x = make_float2(fmaxf(t23.y, -126.f), fmaxf(t67.y, -126.f));
            const float2 rr = fadd2(x, magic);
            const float2 f = fsub2(x, fsub2(rr, magic));
            float2 pp = ffma2(make_float2(0.05500892f, 0.05500892f), f, make_float2(0.24221096f, 0.24221096f));
            pp = ffma2(pp, f, make_float2(0.69328293f, 0.69328293f));
            pp = ffma2(pp, f, make_float2(1.f, 1.f));
            const float p3 = __int_as_float(__float_as_int(pp.x) + (__float_as_int(rr.x) << 23));
            const float p7 = __int_as_float(__float_as_int(pp.y) + (__float_as_int(rr.y) << 23));
            acc0 = fadd2(acc0, make_float2(p0, p1));
            acc1 = fadd2(acc1, make_float2(p2, p3));
            acc0 = fadd2(acc0, make_float2(p4, p5));
            acc1 = fadd2(acc1, make_float2(p6, p7));
            sv[i >> 1] = pack_bf16x2(p0, p1);
            sv[(i >> 1) + 1] = pack_bf16x2(p2, p3);
            sv[(i >> 1) + 2] = pack_bf16x2(p4, p5);
            sv[(i >> 1) + 3] = pack_bf16x2(p6, p7);
          }
          const float2 acc = fadd2(acc0, acc1);
          tmax_out = fmaxf(tm0, tm1);
          return acc.x + acc.y;
        };
        if (dbg) { tB = clk(); d_max += tB - tA; tA = tB; }
        if (NQ == 2 && P.debug != 4) named_bar_sync(1 + g, 256);  // wait for our turn
        bool redo = j == 0;
        if (!redo) {
          float tmax;
          rs = exp_pass(-m_used, tmax);
          redo = __any_sync(0xffffffffu, tmax > kRescaleThreshold);
          if (redo) {  // rare: S is still intact in TMEM (P has not been stored yet)
            uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
            tmem_ld32(lane_base + slot * 128 + 0, sv4[0]);
            tmem_ld32(lane_base + slot * 128 + 32, sv4[1]);
            tmem_ld32(lane_base + slot * 128 + 64, sv4[2]);
            tmem_ld32(lane_base + slot * 128 + 96, sv4[3]);
            tmem_ld_wait();
            if (kv_left < kBKV) {
#pragma unroll
              for (int i = 0; i < 128; ++i)
                if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
            }
          }
        }
        if (redo) {
          float m0 = __uint_as_float(sv[0]), m1 = __uint_as_float(sv[1]);
          float m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
#pragma unroll
          for (int i = 4; i < 128; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(sv[i]));
            m1 = fmaxf(m1, __uint_as_float(sv[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(sv[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(sv[i + 3]));
          }
          const float m_cand = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
          // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one.
          const bool grow = m_cand > m_used + kRescaleThreshold;
          warp_grow = __any_sync(0xffffffffu, grow);
          if (warp_grow) {
            const float m_new = fmaxf(m_used, m_cand);
            alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
            m_used = m_new;
            l *= alpha;
          }
          float tmax;
          rs = exp_pass(-m_used, tmax);
        }
      } else {
      float mx;
      if constexpr (FAST) {  // four independent chains instead of one 128-deep dependency chain
        float m0 = __uint_as_float(sv[0]), m1 = __uint_as_float(sv[1]);
        float m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(sv[i]));
          m1 = fmaxf(m1, __uint_as_float(sv[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(sv[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(sv[i + 3]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      } else {
        mx = __uint_as_float(sv[0]);
#pragma unroll
        for (int i = 1; i < 128; ++i) mx = fmaxf(mx, __uint_as_float(sv[i]));
      }
      const float m_cand = mx * sl2;
      // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one.
      const bool grow = m_cand > m_used + kRescaleThreshold;
      warp_grow = __any_sync(0xffffffffu, grow);
      if (warp_grow) {
        const float m_new = fmaxf(m_used, m_cand);
        alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
        m_used = m_new;
        l *= alpha;
      }
      const float neg_m = -m_used;
      if (dbg) { tB = clk(); d_max += tB - tA; tA = tB; }
      if (NQ == 2 && P.debug != 4) named_bar_sync(1 + g, 256);  // wait for our turn
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          float p = fast_exp2_pinned(fmaf(__uint_as_float(sv[i]), sl2, neg_m));
          rs += p;
          sv[i] = __float_as_uint(p);
        }
      }
      if (NQ == 2 && P.debug != 4) named_bar_arrive(1 + (g ^ 1), 256);  // pass the turn
      l += rs;
      if (dbg) { tB = clk(); d_exp += tB - tA; tA = tB; }

      if (j > 0) {
        if constexpr (!(FAST && TS)) {
          mbar_wait(&o_done[g], (j - 1) & 1);  // PV(j-1) finished: O stable, P buffer reusable
          tc_fence_after();
        }
        if (dbg) { tB = clk(); d_wait_o += tB - tA; tA = tB; }
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
      if constexpr (FAST && TS) {
        uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);  // packed in the exp loop
        tmem_st32(lane_base + slot * 128, pk[0]);
        tmem_st32(lane_base + slot * 128 + 32, pk[1]);
        tmem_st_wait();
      } else if constexpr (TS) {
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
      if (dbg) { tB = clk(); d_st += tB - tA; tA = tB; }
    }
    if (dbg) {
      g_attn_dbg[0] = d_wait_s, g_attn_dbg[1] = d_ld, g_attn_dbg[2] = d_max, g_attn_dbg[3] = d_exp;
      g_attn_dbg[4] = d_wait_o, g_attn_dbg[5] = d_st, g_attn_dbg[6] = n;
      g_attn_dbg[8] = g_attn_dbg[9] = g_attn_dbg[10] = 0;
    }

    // ---------------- epilogue: O / l -> bf16 -> (optional) fp8 ----------------
    if (P.debug != 1) mbar_wait(&o_done[g], (n - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const bool valid = qrow < a.S && P.debug != 1;
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

  pdl_launch_dependents();  // multi-wave grid: let the next kernel in only when this CTA is done
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// =====================================================================================================
// cta_group::2 form of the default kernel: a CLUSTER OF TWO CTAs works on four query tiles of one head and shares
// every K / V tile.  Each CTA keeps the single-CTA structure (two 128-row query tiles, S / P / O in its own TMEM, two
// softmax warpgroups with the exponential token), but the MMAs are M = 256 instructions issued by the leader CTA over
// the pair: tile g of CTA 0 and tile g of CTA 1 are the two M halves, and the B operand is split between the CTAs --
// CTA r stages KV rows [64 r, 64 r + 64) of the K tile (N halves of QK) and head-dim columns [64 r, 64 r + 64) of the V
// tile (N halves of PV).  Per MMA an SM reads 4 + 2 KB (QK) / 2 KB (PV) of shared memory instead of 8 / 4 KB -- the
// single-CTA kernel's MMAs are bound by exactly that operand bandwidth (87-100 cycles per instruction against 64
// nominal; splitting QK into N = 64 halves to shorten the S -> P -> PV -> S chain made the kernel 12 % SLOWER,
// profiles/r2_attention.md) -- and the TMA fill per SM and KV tile halves (32 KB), so the ring is 3 deep.
// Cross-CTA traffic: the leader's issuer needs P from both CTAs (remote mbarrier arrives), every commit is multicast.
// =====================================================================================================
#ifndef FLUXB200_ATTN_PAIR_STAGES
#define FLUXB200_ATTN_PAIR_STAGES 3
#endif
struct AttnPairCfg {
  static constexpr int kStages = FLUXB200_ATTN_PAIR_STAGES;
  static constexpr int kHalfBytes = kTileBytes / 2;  // one CTA's half of a K or V tile: 16 KB
  static constexpr int kQOff = 0;
  static constexpr int kKOff = 2 * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kHalfBytes;
  static constexpr int kBarOff = kVOff + kStages * kHalfBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static constexpr int kThreads = 384;
};

__global__ void __launch_bounds__(AttnPairCfg::kThreads, 1) attention_kernel_pair(const __grid_constant__ AttnParams P,
                                                                                 const __grid_constant__ CUtensorMap tmap_k64) {
  using C = AttnPairCfg;
  constexpr int KS = C::kStages;
  constexpr int SPLIT = 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                 // leader: both CTAs' Q bytes
  uint64_t* k_full = q_full + 1;           // KS, leader
  uint64_t* k_empty = k_full + KS;         // KS, each CTA (multicast commit)
  uint64_t* v_full = k_empty + KS;         // KS, leader
  uint64_t* v_empty = v_full + KS;         // KS, each CTA
  uint64_t* s_ready = v_empty + KS;        // 2, each CTA (multicast commit)
  uint64_t* p_ready = s_ready + 2;         // 2, leader: 4 softmax warps x 2 CTAs
  uint64_t* o_done = p_ready + 2;          // 2, each CTA
  uint64_t* p_lo = o_done + 2;             // 2, leader
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(p_lo + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const fluxb200_attention_args& a = P.a;
  const int q0 = (blockIdx.x >> 1) * (4 * kBQ) + rank * (2 * kBQ);  // this CTA's 256 query rows
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.H + h;
  const int n = P.num_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmap_q);
    tma_prefetch_desc(&tmap_k64);
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
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_ready[g], 1);
      mbar_init(&p_ready[g], 8);  // one arrive per softmax warp of both CTAs
      mbar_init(&o_done[g], 1);
      mbar_init(&p_lo[g], 8);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_ptr, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();  // q, k, v come from the preceding QKV GEMMs

  if (warp < 4) {
    reg_dec<88>();
    if (warp == 0) {
      // ---------------- TMA producer (both CTAs; bytes are accounted on the leader's barriers) ----------------
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * 2 * kTileBytes);
        for (int g = 0; g < 2; ++g) {
          uint8_t* dst = smem + C::kQOff + g * kTileBytes;
          tma_load_3d_2sm(dst, &P.tmap_q, q_full, 0, q0 + g * kBQ, bh, kEvictFirst);
          tma_load_3d_2sm(dst + kChunkBytes, &P.tmap_q, q_full, 64, q0 + g * kBQ, bh, kEvictFirst);
        }
      }
      __syncwarp();
      for (int j = 0; j < n; ++j) {
        const int st = j % KS;
        const uint32_t ph = (j / KS) & 1;
        uint8_t* kd = smem + C::kKOff + st * C::kHalfBytes;   // [2 d-chunks][64 kv rows][128 B]
        uint8_t* vd = smem + C::kVOff + st * C::kHalfBytes;   // [128 kv rows][128 B] = d columns [64 rank, +64)
        mbar_wait(&k_empty[st], ph ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&k_full[st], kTileBytes);
          tma_load_3d_2sm(kd, &tmap_k64, &k_full[st], 0, j * kBKV + rank * 64, bh, kEvictLast);
          tma_load_3d_2sm(kd + C::kHalfBytes / 2, &tmap_k64, &k_full[st], 64, j * kBKV + rank * 64, bh, kEvictLast);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&v_full[st], kTileBytes);
          tma_load_3d_2sm(vd, &P.tmap_v, &v_full[st], rank * 64, j * kBKV, bh, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 1 && rank == 0) {
      // ---------------- MMA issuer (leader CTA only) ----------------
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, 2 * kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, 2 * kBQ, kD, 0, 1);  // V is MN-major
      const uint64_t q_desc0 = make_desc_sw128(smem_u32(smem + C::kQOff), 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(smem_u32(smem + C::kKOff), 16, 1024);
      const uint64_t v_desc0 = make_desc_sw128(smem_u32(smem + C::kVOff), kChunkBytes, 1024);
      auto commit = [&](uint64_t* bar) {  // both CTAs
        if (elect_one()) tc_commit_2sm(bar, 3);
        __syncwarp();
      };
      auto issue_qk = [&](int g, int st) {
        const uint32_t d = tmem_base + g * 128;
        const uint64_t ad0 = desc_advance(q_desc0, g * kTileBytes), bd0 = desc_advance(k_desc0, st * C::kHalfBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t aoff = (kk >> 2) * kChunkBytes + (kk & 3) * 32;          // Q: [2 chunks][128 rows][128 B]
            const uint32_t boff = (kk >> 2) * (C::kHalfBytes / 2) + (kk & 3) * 32;  // K half: [2 chunks][64 rows][128 B]
            mma_f16_ss_2sm(d, desc_advance(ad0, aoff), desc_advance(bd0, boff), idesc_qk, kk != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int st, bool first, int kk0, int kk1) {
        const uint32_t d = tmem_base + 256 + g * 128;
        const uint64_t bd0 = desc_advance(v_desc0, st * C::kHalfBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk)
            mma_f16_ts_2sm(d, tmem_base + g * 128 + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv,
                           (!first || kk != 0) ? 1u : 0u);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int g = 0; g < 2; ++g) {
        issue_qk(g, 0);
        commit(&s_ready[g]);
      }
      commit(&k_empty[0]);
      for (int j = 0; j < n; ++j) {
        const int st = j % KS;
        mbar_wait(&v_full[st], (j / KS) & 1);
        for (int g = 0; g < 2; ++g) {
          mbar_wait(&p_lo[g], j & 1);
          tc_fence_after();
          issue_pv(g, st, j == 0, 0, SPLIT / 16);
          mbar_wait(&p_ready[g], j & 1);
          tc_fence_after();
          issue_pv(g, st, j == 0, SPLIT / 16, kBKV / 16);
          commit(&o_done[g]);
          if (g == 1) commit(&v_empty[st]);
          if (j + 1 < n) {
            const int st1 = (j + 1) % KS;
            if (g == 0) {
              mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
              tc_fence_after();
            }
            issue_qk(g, st1);
            commit(&s_ready[g]);
            if (g == 1) commit(&k_empty[st1]);
          }
        }
      }
    }
  } else {
    // ---------------- softmax warpgroups (identical in both CTAs; P-ready arrives go to the leader) ----------------
    reg_inc<208>();
    const int g = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t o_taddr = lane_base + 256 + g * 128;
    const float sl2 = P.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;
    if (g == 1) named_bar_arrive(1, 256);  // hand the first exponential turn to warpgroup 0
    for (int j = 0; j < n; ++j) {
      if (j > 0) mbar_wait(&o_done[g], (j - 1) & 1);  // O stable, P region reusable
      mbar_wait(&s_ready[g], j & 1);
      tc_fence_after();
      uint32_t sv[128];
      {
        uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(lane_base + g * 128 + 0, sv4[0]);
        tmem_ld32(lane_base + g * 128 + 32, sv4[1]);
        tmem_ld32(lane_base + g * 128 + 64, sv4[2]);
        tmem_ld32(lane_base + g * 128 + 96, sv4[3]);
        tmem_ld_wait();
      }
      const int kv_left = a.S - j * kBKV;
      if (kv_left < kBKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      float m0 = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float m1 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
      float m2 = fmaxf(__uint_as_float(sv[6]), __uint_as_float(sv[7]));
      float m3 = -INFINITY;
#pragma unroll
      for (int i = 8; i < 128; i += 8) {
        m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        m2 = fmax3(m2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
        m3 = fmax3(m3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
      }
      const float m_cand = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
      const bool grow = m_cand > m_used + kRescaleThreshold;
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = fmaxf(m_used, m_cand);
        const float alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
        m_used = m_new;
        l *= alpha;
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t ov[32];
            tmem_ld32(o_taddr + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(o_taddr + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      named_bar_sync(1 + g, 256);  // wait for our turn
      const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(-m_used, -m_used);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int i = hh * SPLIT; i < (hh == 0 ? SPLIT : 128); i += 8) {
          const float2 t01 = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, negmv);
          const float2 t23 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, negmv);
          const float2 t45 = ffma2(make_float2(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])), sl2v, negmv);
          const float2 t67 = ffma2(make_float2(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])), sl2v, negmv);
          const float p0 = fast_exp2_pinned(t01.x), p1 = fast_exp2_pinned(t01.y), p2 = fast_exp2_pinned(t23.x);
          const float p3 = fast_exp2_pinned(t23.y), p4 = fast_exp2_pinned(t45.x), p5 = fast_exp2_pinned(t45.y);
          const float p6 = fast_exp2_pinned(t67.x), p7 = fast_exp2_pinned(t67.y);
          acc0 = fadd2(acc0, make_float2(p0, p1));
          acc1 = fadd2(acc1, make_float2(p2, p3));
          acc0 = fadd2(acc0, make_float2(p4, p5));
          acc1 = fadd2(acc1, make_float2(p6, p7));
          sv[i >> 1] = pack_bf16x2(p0, p1);
          sv[(i >> 1) + 1] = pack_bf16x2(p2, p3);
          sv[(i >> 1) + 2] = pack_bf16x2(p4, p5);
          sv[(i >> 1) + 3] = pack_bf16x2(p6, p7);
        }
        uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_st32(lane_base + g * 128 + hh * 32, pk[hh]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(hh == 0 ? &p_lo[g] : &p_ready[g], 0);  // the leader's barrier
      }
      named_bar_arrive(1 + (g ^ 1), 256);  // pass the turn
      const float2 acc = fadd2(acc0, acc1);
      l += acc.x + acc.y;
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

  pdl_launch_dependents();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

static int launch_attention_pair(const AttnParams& P, cudaStream_t stream) {
  using C = AttnPairCfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  CUtensorMap tmap_k64;
  const uint64_t bhn = static_cast<uint64_t>(a.B) * a.H;
  const uint64_t row_bytes = kD * 2;
  int rc = make_tmap_3d(&tmap_k64, a.k, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBKV / 2, 1);
  if (rc) return rc;
  dim3 grid(2 * ((a.S + 4 * kBQ - 1) / (4 * kBQ)), a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_pair, grid, dim3(C::kThreads), C::kTotal, stream, 2, P, tmap_k64));
  return 0;
}

#ifdef FLUXB200_ATTN_EXPERIMENTS
#include "experiments/attention_variants.cuh"
#endif

template <int NQ, bool TS, bool FAST = false, int SPLIT = 0, int POLY = 2, bool EARLY = false>
static int launch_attention(const AttnParams& P, cudaStream_t stream) {
  using C = AttnCfg<NQ, TS>;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  auto kern = attention_kernel<NQ, TS, FAST, SPLIT, POLY, EARLY>;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + NQ * kBQ - 1) / (NQ * kBQ), a.H, a.B);
  FB_CUDA_OK(launch_kernel(kern, grid, dim3(C::kThreads), C::kTotal, stream, 1, P));
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
  static const int dbg = [] {
    const char* e = getenv("FLUXB200_ATTN_DEBUG");
    return e ? atoi(e) : 0;
  }();
  P.debug = dbg;
  const uint64_t bhn = static_cast<uint64_t>(a.B) * a.H;
  const uint64_t row_bytes = kD * 2;
  int rc;
  if ((rc = make_tmap_3d(&P.tmap_q, a.q, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBQ, 1))) return rc;
  if ((rc = make_tmap_3d(&P.tmap_k, a.k, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBKV, 1))) return rc;
  if ((rc = make_tmap_3d(&P.tmap_v, a.v, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBKV, 1))) return rc;

  // The product library ships ONE attention kernel: 2 query tiles per CTA, P through TMEM, packed-fp32 softmax with
  // every exponential on the MUFU, P handed to the issuer in two 64-column halves.
  // variant 0 = the library's choice between the two product kernels (same arithmetic, bit-identical results):
  //   single  one CTA = two query tiles                                     (variant 17 forces it)
  //   pair    a cluster of two CTAs = four query tiles sharing K / V tiles   (variant 16 forces it)
  // The pair form wins once a head has enough KV tiles to amortise its cluster set-up (measured: -1.5 % at S = 4608,
  // -3.5 % at S = 9728, +2 % at S = 2816; profiles/r2_attention.md).  FLUXB200_ATTN_PAIR=0|1 overrides.
  if (a.variant == 0 || a.variant == 16 || a.variant == 17) {
    static const int forced = [] {
      const char* e = getenv("FLUXB200_ATTN_PAIR");
      return e ? atoi(e) : -1;
    }();
    bool pair = a.S >= 4096;
    if (forced >= 0) pair = forced != 0;
    if (a.variant == 16) pair = true;
    if (a.variant == 17) pair = false;
    return pair ? launch_attention_pair(P, stream) : launch_attention<2, true, true, 64, 0>(P, stream);
  }
#ifdef FLUXB200_ATTN_EXPERIMENTS
  if (a.variant == 18) return launch_attention_step<true>(P, stream);  // step-interleaved pair kernel
  if (a.variant == 19) return launch_attention_step<false>(P, stream);  // ... without the exclusive exp turns
  if (a.variant == 15) return launch_attention<2, true, true, 64, 0, true>(P, stream);  // + early half-QK: 12 % slower
#endif
#ifdef FLUXB200_ATTN_EXPERIMENTS
  switch (a.variant) {
    case 7: return launch_attention<2, true, true>(P, stream);  // one whole-tile P hand-off, max fused into the exp pass
    case 13: return launch_attention<2, true, true, 64, 2>(P, stream);  // 2 of 8 exponentials on the FMA pipe
    case 10: return launch_attention<2, true, true, 64, 4>(P, stream);  // 1/2 of the exponentials on the FMA pipe
    case 11: return launch_attention<2, true, true, 64, 6>(P, stream);  // ... 3/4
    case 14: return launch_attention_coop(P, stream);               // both tiles' softmax split over all 8 softmax warps
    case 12: return launch_attention_one(P, stream);                // 1 query tile, 8 softmax warps, S double-buffered
    case 9: return launch_attention_wide(P, stream);                // 8 softmax warps per query tile (640 threads)
    case 8: return launch_attention<2, true, true, 96>(P, stream);  // 3/4 + 1/4 split of the hand-off
    case 1: return launch_attention<2, true>(P, stream);   // 2 query tiles, whole KV tiles, P through TMEM
    case 5: return launch_attention_halves(P, stream);     // 2 query tiles x 2 KV halves in flight, explicit PV->QK waits
    case 6: return launch_attention_events(P, stream);     // event-driven issuer + fused single-pass softmax
    case 2: return launch_attention<1, false>(P, stream);  // 1 query tile, P through smem (SS MMA)
    case 3: return launch_attention<1, true>(P, stream);   // 1 query tile, P through TMEM
    case 4: return launch_attention<2, false>(P, stream);  // 2 query tiles, P through smem
    default: break;
  }
#endif
  return set_error(FLUXB200_ERR_INVALID, "fluxb200_attention: variant %d is not built into this library (0 is the product "
                   "kernel; others need -DFLUXB200_ATTN_EXPERIMENTS)", a.variant);
}

// Diagnostics: copies the 16 phase counters of the last attention launch (CTA 0) to host memory.
#ifdef FLUXB200_ATTN_PROBE
extern "C" int fluxb200_debug_trace(unsigned long long* host_out576) {
  FB_CUDA_OK(cudaDeviceSynchronize());
  FB_CUDA_OK(cudaMemcpyFromSymbol(host_out576, fb::g_attn_trace, sizeof(unsigned long long) * 576));
  return 0;
}
#endif
extern "C" int fluxb200_debug_counters(unsigned long long* host_out16) {
  FB_CUDA_OK(cudaDeviceSynchronize());
  FB_CUDA_OK(cudaMemcpyFromSymbol(host_out16, fb::g_attn_dbg, sizeof(unsigned long long) * 16));
  return 0;
}
