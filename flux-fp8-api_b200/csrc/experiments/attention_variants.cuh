// Attention tilings that LOST to the default kernel (attention.cu, variant 0) in round 1 -- kept for A/B measurements
// only (profiles/r1_attention_variants.md has the numbers).  Not part of the product library: this file is compiled
// into libflux_b200.so only when attention.cu is built with -DFLUXB200_ATTN_EXPERIMENTS (make EXTRA=-DFLUXB200_ATTN_EXPERIMENTS).
// Included inside namespace fb, after the default kernel and its helpers.
//   5  attention_kernel_halves   2 query tiles x 2 KV halves in flight, explicit PV->QK waits
//   6  attention_kernel_events   event-driven issuer + fused single-pass softmax
//   9  attention_kernel_wide     8 softmax warps per query tile (640 threads)
//   12 attention_kernel_one      1 query tile per CTA, 8 softmax warps, S double-buffered
//   14 attention_kernel_coop     both tiles' softmax split over all 8 softmax warps
//   18 attention_kernel_step     (round 2, end of file) cta_group::2, ONE query tile per CTA, the two softmax warpgroups
//                                take alternate KV steps, S / P double-buffered separately, two MMA-issuing threads;
//                                19 = the same without the exclusive exponential turns (profiles/r2_attention.md)
#pragma once

// =====================================================================================================
// Half-tile pipelined variant (default).  Same roles and TMEM budget as attention_kernel<2, true>, but every
// 128-row KV tile is processed as two independent 64-column halves: S_g[h] (64 fp32 columns) -> softmax ->
// P_g[h] (32 bf16-pair columns over the same TMEM) -> PV_g(h), and QK for half h of the NEXT tile is issued
// right after PV of half h of this one.  While a softmax warpgroup exponentiates one half, the tensor pipe is
// already producing the other half of its next scores, so each query tile has two dependency chains in flight
// (four per SM) instead of one: the MMA round trip (commit -> mbarrier -> tcgen05.ld -> max) that left both
// MUFU and tensor pipes ~50 % idle in the whole-tile version (profiles/r1_attention_lockstep_ncu_full.txt) is
// hidden behind the other half's exp phase.
//   TMEM: S_g[h] at g*128 + h*64, O_g at 256 + g*128.  smem: Q 64 KB + 2 stages x (K 32 KB + V 32 KB).
// =====================================================================================================
struct AttnHCfg {
  static constexpr int kStages = 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = 2 * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kStages * kTileBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static constexpr int kThreads = 384;
};

__global__ void __launch_bounds__(AttnHCfg::kThreads, 1) attention_kernel_halves(const __grid_constant__ AttnParams P) {
  using C = AttnHCfg;
  constexpr int KS = C::kStages;
  constexpr int kHalf = kBKV / 2;  // 64 kv rows
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = q_full + 1;      // KS
  uint64_t* k_empty = k_full + KS;    // KS
  uint64_t* v_full = k_empty + KS;    // KS
  uint64_t* v_empty = v_full + KS;    // KS
  uint64_t* s_ready = v_empty + KS;   // [g][h] = 4
  uint64_t* p_ready = s_ready + 4;    // [g][h] = 4
  uint64_t* o_done = p_ready + 4;     // [g][h] = 4: PV of half h of query tile g retired
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const fluxb200_attention_args& a = P.a;
  const int q0 = blockIdx.x * (2 * kBQ);
  const int h_idx = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.H + h_idx;
  const int n = P.num_kv_tiles;
  const int nt = 2 * n;  // half-steps

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
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);  // one arrive per softmax warp
    }
    for (int i = 0; i < 4; ++i) mbar_init(&o_done[i], 1);
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
    // The producer and issuer warps run their loops warp-uniformly and predicate only the TMA / MMA / commit
    // instructions on one elected lane: operands then live in uniform registers.  (Running the whole loop under
    // `lane == 0` makes ptxas wrap every UTCHMMA in a vector->uniform "waterfall" loop, ~80 cycles per MMA.)
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        for (int g = 0; g < 2; ++g) {
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
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kBQ, kHalf, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kBQ, kD, 0, 1);  // V is MN-major
      const uint32_t q_addr = smem_u32(smem + C::kQOff);
      const uint32_t k_addr = smem_u32(smem + C::kKOff);
      const uint32_t v_addr = smem_u32(smem + C::kVOff);

      // Descriptors are built once; every MMA then costs one 64-bit add per operand (a single thread issues
      // ~50 MMAs per KV tile, so the issue stream itself is on the critical path).
      uint64_t q_desc[2], k_desc[KS], v_desc[KS];
      for (int g = 0; g < 2; ++g) q_desc[g] = make_desc_sw128(q_addr + g * kTileBytes, 16, 1024);
      for (int i = 0; i < KS; ++i) {
        k_desc[i] = make_desc_sw128(k_addr + i * kTileBytes, 16, 1024);
        v_desc[i] = make_desc_sw128(v_addr + i * kTileBytes, kChunkBytes, 1024);
      }
      // S_g[h] = Q_g . K[tile t>>1, rows (t&1)*64 .. +64]^T
      auto issue_qk = [&](int g, int t) {
        const int hh = t & 1, st = (t >> 1) % KS;
        const uint32_t d = tmem_base + g * 128 + hh * kHalf;
        const uint64_t ad0 = q_desc[g];
        const uint64_t bd0 = desc_advance(k_desc[st], hh * (kHalf * 128));
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
            mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
          }
          tc_commit(&s_ready[g * 2 + hh]);
        }
        __syncwarp();
      };
      // O_g += P_g[h] . V[tile t>>1, rows (t&1)*64 .. +64]
      auto issue_pv = [&](int g, int t) {
        const int hh = t & 1, st = (t >> 1) % KS;
        const uint32_t d = tmem_base + 256 + g * 128;
        const uint64_t bd0 = desc_advance(v_desc[st], hh * kHalf * 128);
        const uint32_t a0 = tmem_base + g * 128 + hh * kHalf;
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kHalf / 16; ++kk)
            mma_f16_ts(d, a0 + kk * 8, desc_advance(bd0, kk * 16 * 128), idesc_pv, (t != 0 || kk != 0) ? 1u : 0u);
          tc_commit(&o_done[g * 2 + hh]);
          if (hh == 1 && g == 1) tc_commit(&v_empty[st]);
        }
        __syncwarp();
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < 2; ++t)
        for (int g = 0; g < 2; ++g) issue_qk(g, t);
      if (elect_one()) tc_commit(&k_empty[0]);
      __syncwarp();
      const bool dbg = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
      unsigned long long d_wp = 0, d_wkv = 0, d_iss = 0, tA = clk(), tB = 0;
      for (int t = 0; t < nt; ++t) {
        const int j = t >> 1, hh = t & 1, st = j % KS;
        if (hh == 0) mbar_wait(&v_full[st], (j / KS) & 1);
        if (dbg) { tB = clk(); d_wkv += tB - tA; tA = tB; }
        for (int g = 0; g < 2; ++g) {
          if (P.debug != 1) mbar_wait(&p_ready[g * 2 + hh], j & 1);
          tc_fence_after();
          if (dbg) { tB = clk(); d_wp += tB - tA; tA = tB; }
          issue_pv(g, t);
          if (dbg) { tB = clk(); d_iss += tB - tA; tA = tB; }
        }
        if (t + 2 < nt) {
          const int j1 = (t + 2) >> 1, st1 = j1 % KS;
          if (hh == 0) mbar_wait(&k_full[st1], (j1 / KS) & 1);
          for (int g = 0; g < 2; ++g) {
            // QK(t+2) overwrites the TMEM columns PV(t) reads P from.  MMAs into different accumulators are not
            // ordered with respect to each other, so make the dependency explicit: PV_g(t) must have retired.
            mbar_wait(&o_done[g * 2 + hh], j & 1);
            tc_fence_after();
            if (dbg) { tB = clk(); d_wkv += tB - tA; tA = tB; }
            issue_qk(g, t + 2);
            if (dbg) { tB = clk(); d_iss += tB - tA; tA = tB; }
          }
          if (hh == 1) {
            if (elect_one()) tc_commit(&k_empty[st1]);
            __syncwarp();
          }
        }
      }
      if (dbg) g_attn_dbg[8] = d_wp, g_attn_dbg[9] = d_wkv, g_attn_dbg[10] = d_iss;
      if (P.debug == 1) {  // drain the tensor pipe before the CTA tears down
        if (elect_one()) tc_commit(q_full);
        __syncwarp();
        mbar_wait(q_full, 1);
      }
    }
  } else {
    // ---------------- softmax warpgroups ----------------
    const int g = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t o_taddr = lane_base + 256 + g * 128;
    const float sl2 = P.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;

    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
    unsigned long long d_wait_s = 0, d_ld = 0, d_max = 0, d_exp = 0, d_wait_o = 0, d_st = 0, tA = 0, tB = 0;
    for (int t = 0; t < (P.debug == 1 ? 0 : nt); ++t) {
      const int j = t >> 1, hh = t & 1;
      const uint32_t s_taddr = lane_base + g * 128 + hh * kHalf;
      if (dbg) tA = clk();
      mbar_wait(&s_ready[g * 2 + hh], j & 1);
      tc_fence_after();
      if (dbg) { tB = clk(); d_wait_s += tB - tA; tA = tB; }
      uint32_t sv[kHalf];
      {
        uint32_t(*sv2)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(s_taddr, sv2[0]);
        tmem_ld32(s_taddr + 32, sv2[1]);
        tmem_ld_wait();
      }
      if (dbg) { tB = clk(); d_ld += tB - tA; tA = tB; }
      const int kv_left = a.S - j * kBKV - hh * kHalf;  // columns >= kv_left are out of range (last tile only)
      if (kv_left < kHalf) {
#pragma unroll
        for (int i = 0; i < kHalf; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      float mx0 = __uint_as_float(sv[0]), mx1 = __uint_as_float(sv[1]);
      float mx2 = __uint_as_float(sv[2]), mx3 = __uint_as_float(sv[3]);
#pragma unroll
      for (int i = 4; i < kHalf; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(sv[i]));
        mx1 = fmaxf(mx1, __uint_as_float(sv[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(sv[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(sv[i + 3]));
      }
      const float m_cand = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
      // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one.  A fully masked half
      // (kv_left <= 0 on the last tile) has m_cand = -inf and contributes zeros.
      const bool grow = m_cand > m_used + kRescaleThreshold;
      const bool warp_grow = __any_sync(0xffffffffu, grow);
      float alpha = 1.f;
      if (warp_grow) {
        const float m_new = fmaxf(m_used, m_cand);
        alpha = (m_new == -INFINITY) ? 1.f : fast_exp2(m_used - m_new);
        m_used = m_new;
        l *= alpha;
      }
      float rs0 = 0.f, rs1 = 0.f;
      const float neg_m = (m_used == -INFINITY) ? 0.f : -m_used;
      if (dbg) { tB = clk(); d_max += tB - tA; tA = tB; }
#pragma unroll
      for (int i = 0; i < kHalf; i += 2) {
        float p0 = fast_exp2(fmaf(__uint_as_float(sv[i]), sl2, neg_m));
        float p1 = fast_exp2(fmaf(__uint_as_float(sv[i + 1]), sl2, neg_m));
        rs0 += p0;
        rs1 += p1;
        sv[i >> 1] = pack_bf16x2(p0, p1);  // P as bf16 pairs, in place
      }
      l += rs0 + rs1;
      if (dbg) { tB = clk(); d_exp += tB - tA; tA = tB; }

      // The S slot we are about to overwrite with P(t) last held P(t-2): wait for PV(t-2) (normally long retired;
      // this barrier is waited every time its half comes round, so its phase parity is always unambiguous).
      if (t >= 2) {
        mbar_wait(&o_done[g * 2 + hh], (j - 1) & 1);
        tc_fence_after();
      }
      // The rare O rescale additionally needs the previous half's PV (t-1).  Conditional wait on the other half's
      // barrier is safe: that phase is waited again (unconditionally) at step t+1 before anything can advance it.
      if (t > 0 && (warp_grow || P.debug == 2)) {
        const int tp = t - 1;
        mbar_wait(&o_done[g * 2 + (tp & 1)], (tp >> 1) & 1);
        tc_fence_after();
        if (dbg) { tB = clk(); d_wait_o += tB - tA; tA = tB; }
      }
      if (t > 0 && warp_grow) {
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
      {
        uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_st32(s_taddr, pk[0]);  // 32 columns of bf16 pairs = 64 kv
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g * 2 + hh]);
      if (dbg) { tB = clk(); d_st += tB - tA; tA = tB; }
    }
    if (dbg) {
      g_attn_dbg[0] = d_wait_s, g_attn_dbg[1] = d_ld, g_attn_dbg[2] = d_max, g_attn_dbg[3] = d_exp;
      g_attn_dbg[4] = d_wait_o, g_attn_dbg[5] = d_st, g_attn_dbg[6] = nt;
    }

    // ---------------- epilogue: O / l -> bf16 -> (optional) fp8 ----------------
    if (P.debug != 1) {
      mbar_wait(&o_done[g * 2 + 0], (n - 1) & 1);  // last PV of each half
      mbar_wait(&o_done[g * 2 + 1], (n - 1) & 1);
    }
    tc_fence_after();
    const float inv_l = 1.f / l;
    const bool valid = qrow < a.S && P.debug != 1;
    const bool second = a.out1 != nullptr && qrow >= a.split_row;
    void* const outp = second ? a.out1 : a.out;
    const int64_t obase = second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                       static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h_idx * kD
                                 : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h_idx * kD;
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
          for (int tt = 0; tt < 4; ++tt) {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float o = bf16r(__uint_as_float(ov[q * 16 + tt * 4 + e]) * inv_l);
              f[e] = a.out_fmt == FLUXB200_E5M2 ? quant_pre<1>(o, oscale) : quant_pre<0>(o, oscale);
            }
            if (a.out_fmt == FLUXB200_E5M2)
              w[tt] = to_fp8x2<1>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<1>(f[2], f[3])) << 16);
            else
              w[tt] = to_fp8x2<0>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<0>(f[2], f[3])) << 16);
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

static int launch_attention_halves(const AttnParams& P, cudaStream_t stream) {
  using C = AttnHCfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_halves, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + 2 * kBQ - 1) / (2 * kBQ), a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_halves, grid, dim3(C::kThreads), C::kTotal, stream, 1, P));
  return 0;
}


// =====================================================================================================
// Event-driven variant ("v5", default).  Two query tiles per CTA, whole 128-row KV tiles, P over S in TMEM.
//  * MMA issuer = small state machine polling the barriers of both streams (P_g ready -> issue PV_g; PV_g retired
//    and K loaded -> issue QK_g of the next tile), so neither stream ever waits behind the other's softmax, and
//    the TMEM hazard "QK_g(j+1) overwrites the columns PV_g(j) reads P from" is an explicit wait on PV_g(j)
//    (MMAs into different accumulators are not ordered among themselves).
//  * softmax = one fused pass per tile: p = exp2(s*c - m) against the running (stale) max while the tile max is
//    tracked on the side; the row max no longer sits on the critical path in front of the MUFU work.  The first
//    tile (no running max yet) and the rare tile whose max outgrows the running max by more than 2^8 take an exact
//    two-step path (the latter rescales O).  Packing to bf16 pairs happens in the same loop.
// =====================================================================================================
struct AttnECfg {
  static constexpr int kStages = 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = 2 * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kStages * kTileBytes;
  static constexpr int kTotal = kBarOff + 256 + 1024;
  static constexpr int kThreads = 384;
};

__global__ void __launch_bounds__(AttnECfg::kThreads, 1) attention_kernel_events(const __grid_constant__ AttnParams P) {
  using C = AttnECfg;
  constexpr int KS = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = q_full + 1;      // KS
  uint64_t* k_empty = k_full + KS;    // KS
  uint64_t* v_full = k_empty + KS;    // KS
  uint64_t* v_empty = v_full + KS;    // KS
  uint64_t* s_ready = v_empty + KS;   // [g]
  uint64_t* p_ready = s_ready + 2;    // [g]
  uint64_t* o_done = p_ready + 2;     // [g]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const fluxb200_attention_args& a = P.a;
  const int q0 = blockIdx.x * (2 * kBQ);
  const int h_idx = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.H + h_idx;
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
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_ready[g], 1);
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
  pdl_wait();  // q, k, v come from the preceding QKV GEMMs

  if (warp < 4) {
    reg_dec<80>();
    if (warp == 0) {
      // ---------------- TMA producer (warp-uniform loop, one elected lane issues) ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        for (int g = 0; g < 2; ++g) {
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
      // ---------------- MMA issuer: event-driven over the two query-tile streams ----------------
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, kBQ, kBKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, kBQ, kD, 0, 1);  // V is MN-major
      const uint32_t q_addr = smem_u32(smem + C::kQOff);
      const uint32_t k_addr = smem_u32(smem + C::kKOff);
      const uint32_t v_addr = smem_u32(smem + C::kVOff);
      // (no runtime-indexed arrays in this loop: they would live in local memory)
      const uint64_t q_desc0 = make_desc_sw128(q_addr, 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(k_addr, 16, 1024);
      const uint64_t v_desc0 = make_desc_sw128(v_addr, kChunkBytes, 1024);
      auto ready = [&](uint64_t* bar, uint32_t parity) { return __all_sync(0xffffffffu, mbar_try_wait(bar, parity)); };

      mbar_wait(q_full, 0);
      int jq0 = 0, jq1 = 0;    // next tile whose QK stream 0 / 1 issues
      int jp0 = 0, jp1 = 0;    // next tile whose PV stream 0 / 1 issues
      uint32_t k_uses = 0, v_uses = 0;  // bit st: one of the two streams has already used K / V stage st
      int remaining = 2 * n * 2;  // QK + PV events of both streams
      int idle = 0;
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          int& jq_g = g == 0 ? jq0 : jq1;
          int& jp_g = g == 0 ? jp0 : jp1;
          if (jq_g == jp_g) {
            // next event of this stream: QK(jq).  Needs K(jq) in smem and, because S(jq) overwrites the columns
            // PV(jq-1) reads P from, PV(jq-1) retired.
            const int j = jq_g;
            if (j < n) {
              const int st = j % KS;
              if ((j == 0 || ready(&o_done[g], (j - 1) & 1)) && ready(&k_full[st], (j / KS) & 1)) {
                tc_fence_after();
                const uint32_t d = tmem_base + g * 128;
                const uint64_t ad0 = desc_advance(q_desc0, g * kTileBytes), bd0 = desc_advance(k_desc0, st * kTileBytes);
                if (elect_one()) {
#pragma unroll
                  for (int kk = 0; kk < kD / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
                    mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
                  }
                  tc_commit(&s_ready[g]);
                  if ((k_uses >> st) & 1) tc_commit(&k_empty[st]);  // second stream done with this K tile
                }
                __syncwarp();
                k_uses ^= 1u << st;
                ++jq_g;
                --remaining;
                progressed = true;
              }
            }
          } else {
            // next event: PV(jp).  Needs P(jp) from the softmax warpgroup and V(jp) in smem.
            const int j = jp_g;
            const int st = j % KS;
            if (ready(&p_ready[g], j & 1) && ready(&v_full[st], (j / KS) & 1)) {
              tc_fence_after();
              const uint32_t d = tmem_base + 256 + g * 128;
              const uint64_t bd0 = desc_advance(v_desc0, st * kTileBytes);
              const uint32_t a0 = tmem_base + g * 128;
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < kBKV / 16; ++kk)
                  mma_f16_ts(d, a0 + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, (j != 0 || kk != 0) ? 1u : 0u);
                tc_commit(&o_done[g]);
                if ((v_uses >> st) & 1) tc_commit(&v_empty[st]);
              }
              __syncwarp();
              v_uses ^= 1u << st;
              ++jp_g;
              --remaining;
              progressed = true;
            }
          }
        }
#if FLUXB200_HANG_TRAP_NS
        if (progressed) {
          idle = 0;
        } else if (++idle > (1 << 24)) {
          if (lane == 0) printf("fluxb200: attention MMA issuer stalled (block %d,%d,%d jq %d,%d jp %d,%d)\n", blockIdx.x,
                                blockIdx.y, blockIdx.z, jq0, jq1, jp0, jp1);
          __trap();
        }
#endif
      }
    }
  } else {
    // ---------------- softmax warpgroups ----------------
    reg_inc<208>();
    const int g = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int qrow = q0 + g * kBQ + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t s_taddr = lane_base + g * 128;
    const uint32_t o_taddr = lane_base + 256 + g * 128;
    const float sl2 = P.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;
#ifdef FLUXB200_ATTN_PROBE
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
#else
    constexpr bool dbg = false;  // phase timers compiled out (build with -DFLUXB200_ATTN_PROBE to enable)
#endif
    unsigned long long d_wait_s = 0, d_ld = 0, d_max = 0, d_exp = 0, d_wait_o = 0, d_st = 0, tA = 0, tB = 0;

    for (int j = 0; j < n; ++j) {
      if (dbg) tA = clk();
      mbar_wait(&s_ready[g], j & 1);
      tc_fence_after();
      if (dbg) { tB = clk(); d_wait_s += tB - tA; tA = tB; }
      uint32_t sv[128];
      {
        uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(s_taddr + 0, sv4[0]);
        tmem_ld32(s_taddr + 32, sv4[1]);
        tmem_ld32(s_taddr + 64, sv4[2]);
        tmem_ld32(s_taddr + 96, sv4[3]);
        tmem_ld_wait();
      }
      if (dbg) { tB = clk(); d_ld += tB - tA; tA = tB; }
      const int kv_left = a.S - j * kBKV;  // columns >= kv_left are out of range (last tile only)
      if (kv_left < kBKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      if (j == 0) {
        // no running max yet: exact row max first
        float m0 = __uint_as_float(sv[0]), m1 = __uint_as_float(sv[1]);
        float m2 = __uint_as_float(sv[2]), m3 = __uint_as_float(sv[3]);
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          m0 = fmaxf(m0, __uint_as_float(sv[i]));
          m1 = fmaxf(m1, __uint_as_float(sv[i + 1]));
          m2 = fmaxf(m2, __uint_as_float(sv[i + 2]));
          m3 = fmaxf(m3, __uint_as_float(sv[i + 3]));
        }
        m_used = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
      }
      if (dbg) { tB = clk(); d_max += tB - tA; tA = tB; }
      // fused pass: exponentiate against the running max, track this tile's max on the side
      float mx0 = -INFINITY, mx1 = -INFINITY, rs0 = 0.f, rs1 = 0.f;
      const float neg_m = -m_used;
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        const float s0 = __uint_as_float(sv[i]), s1 = __uint_as_float(sv[i + 1]);
        mx0 = fmaxf(mx0, s0);
        mx1 = fmaxf(mx1, s1);
        const float p0 = fast_exp2(fmaf(s0, sl2, neg_m));
        const float p1 = fast_exp2(fmaf(s1, sl2, neg_m));
        rs0 += p0;
        rs1 += p1;
        sv[i] = __float_as_uint(p0);
        sv[i + 1] = __float_as_uint(p1);
      }
      float rs = rs0 + rs1;
      const float m_cand = fmaxf(mx0, mx1) * sl2;
      // Lazy rescale: the stale max stands unless this tile's max outgrew it by more than 2^8 (rare after the first
      // tiles).  Then every p of this tile, the running sum and O are multiplied by 2^(m_old - m_new).
      const bool grow = m_cand > m_used + kRescaleThreshold;
      const bool warp_grow = __any_sync(0xffffffffu, grow);
      if (dbg) { tB = clk(); d_exp += tB - tA; tA = tB; }
      if (warp_grow) {
        const float m_new = fmaxf(m_used, m_cand);
        const float alpha = fast_exp2(m_used - m_new);  // <= 1; exactly 1 for rows that did not grow
        // Growth beyond 2^64 could have overflowed exp2 against the stale max: redo this tile exactly from the
        // scores still sitting in TMEM (never taken for RMS-normalised q, k: |s*c| is bounded by ~25).
        const bool redo = __any_sync(0xffffffffu, m_cand > m_used + 64.f);
        m_used = m_new;
        l *= alpha;
        if (redo) {
          uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
          tmem_ld32(s_taddr + 0, sv4[0]);
          tmem_ld32(s_taddr + 32, sv4[1]);
          tmem_ld32(s_taddr + 64, sv4[2]);
          tmem_ld32(s_taddr + 96, sv4[3]);
          tmem_ld_wait();
          rs = 0.f;
#pragma unroll
          for (int i = 0; i < 128; ++i) {
            const float p = (i < kv_left) ? fast_exp2(fmaf(__uint_as_float(sv[i]), sl2, -m_new)) : 0.f;
            rs += p;
            sv[i] = __float_as_uint(p);
          }
        } else {
          rs *= alpha;
#pragma unroll
          for (int i = 0; i < 128; ++i) sv[i] = __float_as_uint(__uint_as_float(sv[i]) * alpha);
        }
        if (j > 0) {
          mbar_wait(&o_done[g], (j - 1) & 1);  // O stable (already retired: QK(j) was only issued after it)
          tc_fence_after();
          if (dbg) { tB = clk(); d_wait_o += tB - tA; tA = tB; }
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
      l += rs;
      // P (bf16 pairs) over the S columns: column c holds kv (2c, 2c+1) of this row
      {
        uint32_t pk[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            pk[i] = pack_bf16x2(__uint_as_float(sv[half * 64 + 2 * i]), __uint_as_float(sv[half * 64 + 2 * i + 1]));
          tmem_st32(s_taddr + half * 32, pk);
        }
        tmem_st_wait();
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
    mbar_wait(&o_done[g], (n - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l;
    const bool valid = qrow < a.S;
    const bool second = a.out1 != nullptr && qrow >= a.split_row;
    void* const outp = second ? a.out1 : a.out;
    const int64_t obase = second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                       static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h_idx * kD
                                 : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h_idx * kD;
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
          for (int tt = 0; tt < 4; ++tt) {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float o = bf16r(__uint_as_float(ov[q * 16 + tt * 4 + e]) * inv_l);
              f[e] = a.out_fmt == FLUXB200_E5M2 ? quant_pre<1>(o, oscale) : quant_pre<0>(o, oscale);
            }
            if (a.out_fmt == FLUXB200_E5M2)
              w[tt] = to_fp8x2<1>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<1>(f[2], f[3])) << 16);
            else
              w[tt] = to_fp8x2<0>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<0>(f[2], f[3])) << 16);
          }
          dst[q] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
  }

  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_attention_events(const AttnParams& P, cudaStream_t stream) {
  using C = AttnECfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_events, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + 2 * kBQ - 1) / (2 * kBQ), a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_events, grid, dim3(C::kThreads), C::kTotal, stream, 1, P));
  return 0;
}

// =====================================================================================================
// Wide-softmax variant: two query tiles per CTA as above, but EIGHT softmax warps per tile (640 threads):
// the warps w and w+4 of a tile share the tile's TMEM lanes (rows) and own the KV columns [0,64) and [64,128).
// Why: with four warps per tile each SM sub-partition holds exactly one warp of a tile, so during the tile's
// exponential phase that sub-partition issues from ONE warp and every dependency stall is exposed -- the pass
// took ~1540 cycles for ~620 issue slots / 768 MUFU cycles, and the two tiles' passes (serialised by the MUFU
// token) summed to the whole loop.  Two warps per sub-partition halve the pass; the kernel then leans on the
// tensor pipe (2 x (PV + QK) per loop) instead of on softmax latency.
// The two half-row threads agree on the row max through shared memory + a 64-thread named barrier, keep separate
// partial row sums (added once at the end) and each rescale / write out their own 64 columns of O.
// =====================================================================================================
struct AttnWCfg {
  static constexpr int kStages = 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = 2 * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kStages * kTileBytes;
  static constexpr int kXchOff = kBarOff + 256;                 // [2 tiles][2 buffers][2 halves][128 rows] fp32
  static constexpr int kTotal = kXchOff + 2 * 2 * 2 * kBQ * 4 + 1024;
  static constexpr int kThreads = 128 + 2 * 256;
};

__global__ void __launch_bounds__(AttnWCfg::kThreads, 1) attention_kernel_wide(const __grid_constant__ AttnParams P) {
  using C = AttnWCfg;
  constexpr int KS = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = q_full + 1;           // KS
  uint64_t* k_empty = k_full + KS;         // KS
  uint64_t* v_full = k_empty + KS;         // KS
  uint64_t* v_empty = v_full + KS;         // KS
  uint64_t* s_ready = v_empty + KS;        // 2
  uint64_t* p_lo = s_ready + 2;            // 2: columns [0,64) of P stored (4 warps)
  uint64_t* p_hi = p_lo + 2;               // 2: columns [64,128)
  uint64_t* o_done = p_hi + 2;             // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);
  const uint32_t xch_saddr = smem_u32(smem + C::kXchOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const fluxb200_attention_args& a = P.a;
  const int q0 = blockIdx.x * (2 * kBQ);
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
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_ready[g], 1);
      mbar_init(&p_lo[g], 4);
      mbar_init(&p_hi[g], 4);
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
  pdl_wait();  // q, k, v come from the preceding QKV GEMMs

  if (warp < 4) {
    reg_dec<56>();
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        for (int g = 0; g < 2; ++g) {
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
      const uint64_t q_desc0 = make_desc_sw128(q_addr, 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(k_addr, 16, 1024);
      const uint64_t v_desc0 = make_desc_sw128(v_addr, kChunkBytes, 1024);
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
        __syncwarp();
      };
      auto issue_qk = [&](int g, int st) {
        const uint32_t d = tmem_base + g * 128;
        const uint64_t ad0 = desc_advance(q_desc0, g * kTileBytes), bd0 = desc_advance(k_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
            mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int st, bool first, int kk0, int kk1) {
        const uint32_t d = tmem_base + 256 + g * 128;
        const uint64_t bd0 = desc_advance(v_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk) {
            const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
            mma_f16_ts(d, tmem_base + g * 128 + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, acc);
          }
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
          issue_pv(g, st, j == 0, 0, 4);
          mbar_wait(&p_hi[g], j & 1);
          tc_fence_after();
          issue_pv(g, st, j == 0, 4, 8);
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
    // ---------------- softmax: 8 warps per query tile ----------------
    reg_inc<104>();  // (88 / 96 instead spills 500 bytes per softmax thread: 593 us)
    const int g = (warp - 4) >> 3;          // query tile
    const int hc = ((warp - 4) >> 2) & 1;   // column half of the KV tile: [64*hc, 64*hc + 64)
    const int lg = warp & 3;                // TMEM lane group
    const int r = lg * 32 + lane;           // row within the query tile
    const int qrow = q0 + g * kBQ + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t s_taddr = lane_base + g * 128 + hc * 64;       // my 64 S columns
    const uint32_t p_taddr = lane_base + g * 128 + hc * 32;       // my 32 packed P columns
    const uint32_t o_taddr = lane_base + 256 + g * 128 + hc * 64; // my 64 O columns
    const uint32_t pair_bar = 3 + g * 4 + lg;                     // named barrier of the two warps sharing my rows
    const float sl2 = P.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;
    if (g == 1 && P.debug != 4) named_bar_arrive(1, 512);  // hand the first MUFU turn to tile 0

    for (int j = 0; j < n; ++j) {
      if (j > 0) mbar_wait(&o_done[g], (j - 1) & 1);  // PV(j-1) finished: O stable, P columns reusable
      mbar_wait(&s_ready[g], j & 1);
      tc_fence_after();
      uint32_t sv[64];
      {
        uint32_t(*sv2)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(s_taddr, sv2[0]);
        tmem_ld32(s_taddr + 32, sv2[1]);
        tmem_ld_wait();
      }
      const int kv_left = a.S - j * kBKV - hc * 64;  // my columns >= kv_left are out of range (last tile only)
      if (kv_left < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      float m0 = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float m1 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
      float m2 = fmaxf(__uint_as_float(sv[6]), __uint_as_float(sv[7]));
      float m3 = -INFINITY;
#pragma unroll
      for (int i = 8; i < 64; i += 8) {
        m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        m2 = fmax3(m2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
        m3 = fmax3(m3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
      }
      const float mx_half = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      // row max = max over the two half-row threads (exchange buffer double-buffered by step parity)
      const uint32_t xb = xch_saddr + ((g * 2 + (j & 1)) * 2) * kBQ * 4;
      sts_f32(xb + (hc * kBQ + r) * 4, mx_half);
      named_bar_sync(pair_bar, 64);
      const float m_cand = fmaxf(mx_half, lds_f32(xb + ((hc ^ 1) * kBQ + r) * 4)) * sl2;
      // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one (identical decision in both
      // half-row threads: same rows, same m_used, same m_cand).
      const bool grow = m_cand > m_used + kRescaleThreshold;
      const bool warp_grow = __any_sync(0xffffffffu, grow);
      if (warp_grow) {
        const float m_new = fmaxf(m_used, m_cand);
        const float alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
        m_used = m_new;
        l *= alpha;
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
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
      if (P.debug != 4) named_bar_sync(1 + g, 512);  // wait for this tile's turn on the MUFU
      const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(-m_used, -m_used);
      const float2 magic = make_float2(12582912.f, 12582912.f);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        const float2 t01 = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, negmv);
        const float2 t23 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, negmv);
        const float2 t45 = ffma2(make_float2(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])), sl2v, negmv);
        const float2 t67 = ffma2(make_float2(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])), sl2v, negmv);
        const float p0 = fast_exp2_pinned(t01.x), p1 = fast_exp2_pinned(t01.y), p2 = fast_exp2_pinned(t23.x);
        const float p4 = fast_exp2_pinned(t45.x), p5 = fast_exp2_pinned(t45.y), p6 = fast_exp2_pinned(t67.x);
        // poly_exp2 on the pair (t23.y, t67.y): a quarter of the exponentials on the FMA pipe
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
      if (P.debug != 4) named_bar_arrive(1 + (g ^ 1), 512);  // pass the turn
      {
        uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_st32(p_taddr, pk[0]);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(hc == 0 ? &p_lo[g] : &p_hi[g]);
      const float2 acc = fadd2(acc0, acc1);
      l += acc.x + acc.y;
    }

    // ---------------- epilogue: O / l -> bf16 -> (optional) fp8; this thread writes its 64 columns ----------------
    mbar_wait(&o_done[g], (n - 1) & 1);
    tc_fence_after();
    // row sum = the two half-row partial sums (same stale-max history in both threads)
    const uint32_t xb = xch_saddr + ((g * 2 + (n & 1)) * 2) * kBQ * 4;
    sts_f32(xb + (hc * kBQ + r) * 4, l);
    named_bar_sync(pair_bar, 64);
    const float l_other = lds_f32(xb + ((hc ^ 1) * kBQ + r) * 4);
    const float inv_l = 1.f / (hc == 0 ? l + l_other : l_other + l);
    const bool valid = qrow < a.S;
    const bool second = a.out1 != nullptr && qrow >= a.split_row;
    void* const outp = second ? a.out1 : a.out;
    const int64_t obase = (second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                        static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h * kD
                                  : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h * kD) +
                          hc * 64;
    float oscale = 1.f;
    if (a.out_kind == 1) oscale = __ldg(qrow < a.split_row ? a.out_scale0 : a.out_scale1);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
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

static int launch_attention_wide(const AttnParams& P, cudaStream_t stream) {
  using C = AttnWCfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + 2 * kBQ - 1) / (2 * kBQ), a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_wide, grid, dim3(C::kThreads), C::kTotal, stream, 1, P));
  return 0;
}


// =====================================================================================================
// Cooperative-halves variant: the layout of the default kernel (two query tiles, 384 threads, K/V shared by 256 query
// rows) with the softmax work of EVERY tile-step split over all eight softmax warps (see the comment at the softmax
// section).  The per-tile chain S -> softmax -> P -> PV -> next S keeps its MMA half, its softmax half is cut in two.
// =====================================================================================================
struct AttnCCfg {
  static constexpr int kStages = 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = 2 * kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kStages * kTileBytes;
  static constexpr int kXchOff = kBarOff + 256;                 // [2 tiles][2 buffers][2 halves][128 rows] fp32
  static constexpr int kTotal = kXchOff + 2 * 2 * 2 * kBQ * 4 + 1024;
  static constexpr int kThreads = 128 + 256;
};

__global__ void __launch_bounds__(AttnCCfg::kThreads, 1) attention_kernel_coop(const __grid_constant__ AttnParams P) {
  using C = AttnCCfg;
  constexpr int KS = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = q_full + 1;           // KS
  uint64_t* k_empty = k_full + KS;         // KS
  uint64_t* v_full = k_empty + KS;         // KS
  uint64_t* v_empty = v_full + KS;         // KS
  uint64_t* s_ready = v_empty + KS;        // 2
  uint64_t* p_lo = s_ready + 2;            // 2: columns [0,64) of P stored (4 warps)
  uint64_t* p_hi = p_lo + 2;               // 2: columns [64,128)
  uint64_t* o_done = p_hi + 2;             // 2
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);
  const uint32_t xch_saddr = smem_u32(smem + C::kXchOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const fluxb200_attention_args& a = P.a;
  const int q0 = blockIdx.x * (2 * kBQ);
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
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_ready[g], 1);
      mbar_init(&p_lo[g], 4);
      mbar_init(&p_hi[g], 4);
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
  pdl_wait();  // q, k, v come from the preceding QKV GEMMs

  if (warp < 4) {
    reg_dec<88>();
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        for (int g = 0; g < 2; ++g) {
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
      const uint64_t q_desc0 = make_desc_sw128(q_addr, 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(k_addr, 16, 1024);
      const uint64_t v_desc0 = make_desc_sw128(v_addr, kChunkBytes, 1024);
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
        __syncwarp();
      };
      auto issue_qk = [&](int g, int st) {
        const uint32_t d = tmem_base + g * 128;
        const uint64_t ad0 = desc_advance(q_desc0, g * kTileBytes), bd0 = desc_advance(k_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
            mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int g, int st, bool first, int kk0, int kk1) {
        const uint32_t d = tmem_base + 256 + g * 128;
        const uint64_t bd0 = desc_advance(v_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk) {
            const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
            mma_f16_ts(d, tmem_base + g * 128 + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, acc);
          }
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
          issue_pv(g, st, j == 0, 0, 4);
          mbar_wait(&p_hi[g], j & 1);
          tc_fence_after();
          issue_pv(g, st, j == 0, 4, 8);
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
    // ---------------- softmax: 8 warps, every one of them works on BOTH query tiles ----------------
    // Warps 4-7 own KV columns [0,64) and warps 8-11 columns [64,128) of whichever tile's scores are ready; warp w and
    // w+4 have the same w % 4, i.e. the same TMEM lanes (rows) AND the same SM sub-partition, so a tile's exponential
    // pass runs as two concurrent warps per sub-partition (their FMA / ALU / MUFU work overlaps) and takes half as
    // long, while the tensor pipe works on the other tile.  No extra threads or registers compared with one
    // warpgroup per tile; a thread carries the running max / partial row sum of its row in both tiles.
    reg_inc<208>();  // 128 x 88 + 256 x 208 = 64512 = 384 x 168
    const int hc = (warp - 4) >> 2;         // column half of the KV tile: [64*hc, 64*hc + 64)
    const int lg = warp & 3;                // TMEM lane group == SM sub-partition
    const int r = lg * 32 + lane;           // row within either query tile
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t pair_bar = 3 + lg;       // named barrier of the two warps sharing my rows
    const float sl2 = P.scale_log2;
    float m_used[2] = {-INFINITY, -INFINITY};
    float l[2] = {0.f, 0.f};

    for (int j = 0; j < n; ++j) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const uint32_t s_taddr = lane_base + g * 128 + hc * 64;        // my 64 S columns of tile g
        const uint32_t p_taddr = lane_base + g * 128 + hc * 32;        // my 32 packed P columns
        const uint32_t o_taddr = lane_base + 256 + g * 128 + hc * 64;  // my 64 O columns
        if (j > 0) mbar_wait(&o_done[g], (j - 1) & 1);  // PV(j-1) finished: O stable, P columns reusable
        mbar_wait(&s_ready[g], j & 1);
        tc_fence_after();
        uint32_t sv[64];
        {
          uint32_t(*sv2)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
          tmem_ld32(s_taddr, sv2[0]);
          tmem_ld32(s_taddr + 32, sv2[1]);
          tmem_ld_wait();
        }
        const int kv_left = a.S - j * kBKV - hc * 64;  // my columns >= kv_left are out of range (last tile only)
        if (kv_left < 64) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
        }
        float m0 = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
        float m1 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
        float m2 = fmaxf(__uint_as_float(sv[6]), __uint_as_float(sv[7]));
        float m3 = -INFINITY;
#pragma unroll
        for (int i = 8; i < 64; i += 8) {
          m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
          m2 = fmax3(m2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
          m3 = fmax3(m3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
        }
        const float mx_half = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        // row max = max over the two half-row threads (exchange buffer double-buffered by step parity)
        const uint32_t xb = xch_saddr + ((g * 2 + (j & 1)) * 2) * kBQ * 4;
        sts_f32(xb + (hc * kBQ + r) * 4, mx_half);
        named_bar_sync(pair_bar, 64);
        const float m_cand = fmaxf(mx_half, lds_f32(xb + ((hc ^ 1) * kBQ + r) * 4)) * sl2;
        // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one (identical decision in both
        // half-row threads: same rows, same m_used, same m_cand).
        const bool grow = m_cand > m_used[g] + kRescaleThreshold;
        const bool warp_grow = __any_sync(0xffffffffu, grow);
        if (warp_grow) {
          const float m_new = fmaxf(m_used[g], m_cand);
          const float alpha = fast_exp2(m_used[g] - m_new);  // exp2(-inf) = 0 on the first tile
          m_used[g] = m_new;
          l[g] *= alpha;
          if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
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
        const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(-m_used[g], -m_used[g]);
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
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
        {
          uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
          tmem_st32(p_taddr, pk[0]);
          tmem_st_wait();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(hc == 0 ? &p_lo[g] : &p_hi[g]);
        const float2 acc = fadd2(acc0, acc1);
        l[g] += acc.x + acc.y;
      }
    }

    // ---------------- epilogue: O / l -> bf16 -> (optional) fp8; this thread writes its 64 columns of both tiles ----------------
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      const uint32_t o_taddr = lane_base + 256 + g * 128 + hc * 64;
      const int qrow = q0 + g * kBQ + r;
      mbar_wait(&o_done[g], (n - 1) & 1);
      tc_fence_after();
      // row sum = the two half-row partial sums (same stale-max history in both threads)
      const uint32_t xb = xch_saddr + ((g * 2 + (n & 1)) * 2) * kBQ * 4;
      sts_f32(xb + (hc * kBQ + r) * 4, l[g]);
      named_bar_sync(pair_bar, 64);
      const float l_other = lds_f32(xb + ((hc ^ 1) * kBQ + r) * 4);
      const float inv_l = 1.f / (hc == 0 ? l[g] + l_other : l_other + l[g]);
      const bool valid = qrow < a.S;
      const bool second = a.out1 != nullptr && qrow >= a.split_row;
      void* const outp = second ? a.out1 : a.out;
      const int64_t obase = (second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                          static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h * kD
                                    : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h * kD) +
                            hc * 64;
      float oscale = 1.f;
      if (a.out_kind == 1) oscale = __ldg(qrow < a.split_row ? a.out_scale0 : a.out_scale1);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
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
  }

  pdl_launch_dependents();  // multi-wave grid: let the next kernel in only when this CTA is done
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_attention_coop(const AttnParams& P, cudaStream_t stream) {
  using C = AttnCCfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + 2 * kBQ - 1) / (2 * kBQ), a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_coop, grid, dim3(C::kThreads), C::kTotal, stream, 1, P));
  return 0;
}


// =====================================================================================================
// One query tile per CTA, EIGHT softmax warps on it (384 threads, 208 registers each for the softmax warps), S
// double-buffered in TMEM: QK(j+1) is issued before the issuer waits for P(j), so the tensor pipe computes the next
// scores while the softmax warps work -- the MMA and softmax halves of the per-tile chain of the two-tile kernels
// overlap within ONE tile, and two warps per SM sub-partition share the exponential pass.  Costs: K/V are fetched
// from L2 once per 128 query rows instead of once per 256.
// =====================================================================================================
struct AttnOCfg {
  static constexpr int kStages = 2;
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kTileBytes;
  static constexpr int kVOff = kKOff + kStages * kTileBytes;
  static constexpr int kBarOff = kVOff + kStages * kTileBytes;
  static constexpr int kXchOff = kBarOff + 256;                 // [2 buffers][2 halves][128 rows] fp32
  static constexpr int kTotal = kXchOff + 2 * 2 * kBQ * 4 + 1024;
  static constexpr int kThreads = 128 + 256;
};

__global__ void __launch_bounds__(AttnOCfg::kThreads, 1) attention_kernel_one(const __grid_constant__ AttnParams P) {
  using C = AttnOCfg;
  constexpr int KS = C::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = q_full + 1;           // KS
  uint64_t* k_empty = k_full + KS;         // KS
  uint64_t* v_full = k_empty + KS;         // KS
  uint64_t* v_empty = v_full + KS;         // KS
  uint64_t* s_ready = v_empty + KS;        // 2
  uint64_t* p_lo = s_ready + 2;            // 1 (+1 unused): columns [0,64) of P stored (4 warps)
  uint64_t* p_hi = p_lo + 2;               // 1 (+1 unused): columns [64,128)
  uint64_t* o_done = p_hi + 2;             // 1 (+1 unused)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_done + 2);
  const uint32_t xch_saddr = smem_u32(smem + C::kXchOff);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const fluxb200_attention_args& a = P.a;
  const int q0 = blockIdx.x * kBQ;
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
    for (int g = 0; g < 2; ++g) mbar_init(&s_ready[g], 1);  // one per S slot
    mbar_init(&p_lo[0], 4);
    mbar_init(&p_hi[0], 4);
    mbar_init(&o_done[0], 1);
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
    reg_dec<88>();
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        mbar_arrive_expect_tx(q_full, kTileBytes);
        uint8_t* dst = smem + C::kQOff;
        tma_load_3d(dst, &P.tmap_q, q_full, 0, q0, bh, kEvictFirst);
        tma_load_3d(dst + kChunkBytes, &P.tmap_q, q_full, 64, q0, bh, kEvictFirst);
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
      const uint64_t q_desc0 = make_desc_sw128(q_addr, 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(k_addr, 16, 1024);
      const uint64_t v_desc0 = make_desc_sw128(v_addr, kChunkBytes, 1024);
      auto commit = [&](uint64_t* bar) {
        if (elect_one()) tc_commit(bar);
        __syncwarp();
      };
      auto issue_qk = [&](int slot, int st) {
        const uint32_t d = tmem_base + slot * 128;
        const uint64_t ad0 = q_desc0, bd0 = desc_advance(k_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < kD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
            mma_f16_ss(d, desc_advance(ad0, off), desc_advance(bd0, off), idesc_qk, kk != 0 ? 1u : 0u);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int slot, int st, bool first, int kk0, int kk1) {
        const uint32_t d = tmem_base + 256;
        const uint64_t bd0 = desc_advance(v_desc0, st * kTileBytes);
        if (elect_one()) {
#pragma unroll
          for (int kk = kk0; kk < kk1; ++kk) {
            const uint32_t acc = (!first || kk != 0) ? 1u : 0u;
            mma_f16_ts(d, tmem_base + slot * 128 + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, acc);
          }
        }
        __syncwarp();
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      commit(&s_ready[0]);
      commit(&k_empty[0]);
      for (int j = 0; j < n; ++j) {
        const int st = j % KS;
        if (j + 1 < n) {
          // S(j+1) goes to the other slot while the softmax warps work on S(j): the tensor pipe never waits for them.
          // (That slot held P(j-1); PV(j-1) was issued before this QK and the pipe executes one issuer's MMAs in order.)
          const int st1 = (j + 1) % KS;
          mbar_wait(&k_full[st1], ((j + 1) / KS) & 1);
          tc_fence_after();
          issue_qk((j + 1) & 1, st1);
          commit(&s_ready[(j + 1) & 1]);
          commit(&k_empty[st1]);
        }
        mbar_wait(&v_full[st], (j / KS) & 1);
        mbar_wait(&p_lo[0], j & 1);
        tc_fence_after();
        issue_pv(j & 1, st, j == 0, 0, 4);
        mbar_wait(&p_hi[0], j & 1);
        tc_fence_after();
        issue_pv(j & 1, st, j == 0, 4, 8);
        commit(&o_done[0]);
        commit(&v_empty[st]);
      }
    }
  } else {
    // ---------------- softmax: 8 warps per query tile ----------------
    reg_inc<208>();  // 128 x 88 + 256 x 208 = 64512 = 384 x 168
    constexpr int g = 0;
    const int hc = ((warp - 4) >> 2) & 1;   // column half of the KV tile: [64*hc, 64*hc + 64)
    const int lg = warp & 3;                // TMEM lane group
    const int r = lg * 32 + lane;           // row within the query tile
    const int qrow = q0 + g * kBQ + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t o_taddr = lane_base + 256 + hc * 64;           // my 64 O columns
    const uint32_t pair_bar = 3 + lg;                             // named barrier of the two warps sharing my rows
    const float sl2 = P.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;

    for (int j = 0; j < n; ++j) {
      const int slot = j & 1;
      const uint32_t s_taddr = lane_base + slot * 128 + hc * 64;  // my 64 S columns
      const uint32_t p_taddr = lane_base + slot * 128 + hc * 32;  // my 32 packed P columns
      mbar_wait(&s_ready[slot], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[64];
      {
        uint32_t(*sv2)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(s_taddr, sv2[0]);
        tmem_ld32(s_taddr + 32, sv2[1]);
        tmem_ld_wait();
      }
      const int kv_left = a.S - j * kBKV - hc * 64;  // my columns >= kv_left are out of range (last tile only)
      if (kv_left < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= kv_left) sv[i] = __float_as_uint(-INFINITY);
      }
      float m0 = fmax3(__uint_as_float(sv[0]), __uint_as_float(sv[1]), __uint_as_float(sv[2]));
      float m1 = fmax3(__uint_as_float(sv[3]), __uint_as_float(sv[4]), __uint_as_float(sv[5]));
      float m2 = fmaxf(__uint_as_float(sv[6]), __uint_as_float(sv[7]));
      float m3 = -INFINITY;
#pragma unroll
      for (int i = 8; i < 64; i += 8) {
        m0 = fmax3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        m1 = fmax3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        m2 = fmax3(m2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
        m3 = fmax3(m3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
      }
      const float mx_half = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      // row max = max over the two half-row threads (exchange buffer double-buffered by step parity)
      const uint32_t xb = xch_saddr + ((g * 2 + (j & 1)) * 2) * kBQ * 4;
      sts_f32(xb + (hc * kBQ + r) * 4, mx_half);
      named_bar_sync(pair_bar, 64);
      const float m_cand = fmaxf(mx_half, lds_f32(xb + ((hc ^ 1) * kBQ + r) * 4)) * sl2;
      // Lazy rescale: keep the stale max unless it is more than 2^8 below the new one (identical decision in both
      // half-row threads: same rows, same m_used, same m_cand).
      const bool grow = m_cand > m_used + kRescaleThreshold;
      const bool warp_grow = __any_sync(0xffffffffu, grow);
      if (warp_grow) {
        const float m_new = fmaxf(m_used, m_cand);
        const float alpha = fast_exp2(m_used - m_new);  // exp2(-inf) = 0 on the first tile
        m_used = m_new;
        l *= alpha;
        if (j > 0) {
          mbar_wait(&o_done[0], (j - 1) & 1);  // PV(j-1) finished: O stable (only needed on this rare path)
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
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
      const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(-m_used, -m_used);
      const float2 magic = make_float2(12582912.f, 12582912.f);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        const float2 t01 = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, negmv);
        const float2 t23 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, negmv);
        const float2 t45 = ffma2(make_float2(__uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5])), sl2v, negmv);
        const float2 t67 = ffma2(make_float2(__uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7])), sl2v, negmv);
        const float p0 = fast_exp2_pinned(t01.x), p1 = fast_exp2_pinned(t01.y), p2 = fast_exp2_pinned(t23.x);
        const float p4 = fast_exp2_pinned(t45.x), p5 = fast_exp2_pinned(t45.y), p6 = fast_exp2_pinned(t67.x);
        // poly_exp2 on the pair (t23.y, t67.y): a quarter of the exponentials on the FMA pipe
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
      {
        uint32_t(*pk)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_st32(p_taddr, pk[0]);
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(hc == 0 ? &p_lo[g] : &p_hi[g]);
      const float2 acc = fadd2(acc0, acc1);
      l += acc.x + acc.y;
    }

    // ---------------- epilogue: O / l -> bf16 -> (optional) fp8; this thread writes its 64 columns ----------------
    mbar_wait(&o_done[g], (n - 1) & 1);
    tc_fence_after();
    // row sum = the two half-row partial sums (same stale-max history in both threads)
    const uint32_t xb = xch_saddr + ((g * 2 + (n & 1)) * 2) * kBQ * 4;
    sts_f32(xb + (hc * kBQ + r) * 4, l);
    named_bar_sync(pair_bar, 64);
    const float l_other = lds_f32(xb + ((hc ^ 1) * kBQ + r) * 4);
    const float inv_l = 1.f / (hc == 0 ? l + l_other : l_other + l);
    const bool valid = qrow < a.S;
    const bool second = a.out1 != nullptr && qrow >= a.split_row;
    void* const outp = second ? a.out1 : a.out;
    const int64_t obase = (second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                        static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h * kD
                                  : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h * kD) +
                          hc * 64;
    float oscale = 1.f;
    if (a.out_kind == 1) oscale = __ldg(qrow < a.split_row ? a.out_scale0 : a.out_scale1);
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
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

static int launch_attention_one(const AttnParams& P, cudaStream_t stream) {
  using C = AttnOCfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_one, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  dim3 grid((a.S + kBQ - 1) / kBQ, a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_one, grid, dim3(C::kThreads), C::kTotal, stream, 1, P));
  return 0;
}

// =====================================================================================================
// "Step-interleaved" cta_group::2 kernel: ONE 128-row query tile per CTA (a cluster of two CTAs = 256 query rows of one
// head sharing every K / V tile), and the two softmax warpgroups of a CTA work on ALTERNATE KV STEPS of that tile
// instead of on two different tiles.  TMEM then has room for S and P double-buffered SEPARATELY:
//     S[0] S[1] (2 x 128 columns, step parity)   P[0] P[1] (2 x 64)   O (128)   = 512 columns
// so P(j) no longer lives on top of S(j) and the serial chain  S -> softmax -> P -> PV -> QK -> S  of the two-tile kernels
// disappears: QK(j+2) is issued right after PV(j) into the slot whose S(j) is already in registers, S(j+1) has been
// waiting since step j-1, and warpgroup (j+1)&1 exponentiates step j+1 while warpgroup j&1 is still storing P(j).  The MUFU
// sees two warps per sub-partition in different phases (no exclusive turns), the tensor pipe is only ever waiting for P.
// Online-softmax state crosses warpgroups: step j needs the reference maximum of step j-1, published per row through
// shared memory (m_pub) with a one-directional named-barrier hand-off; each warpgroup keeps its own partial row sum in
// the units of the maximum it last used; whoever raises the maximum rescales O (after PV(j-1) has retired and before
// it releases P(j)).  The epilogue splits the head dimension between the warpgroups.
// =====================================================================================================
struct AttnStepCfg {
  static constexpr int kKStages = 6;  // K runs three steps ahead of V (QK(j+3) is issued before PV(j))
  static constexpr int kVStages = 4;
  static constexpr int kHalfBytes = kTileBytes / 2;  // one CTA's half of a K or V tile: 16 KB
  static constexpr int kQOff = 0;
  static constexpr int kKOff = kTileBytes;
  static constexpr int kVOff = kKOff + kKStages * kHalfBytes;
  static constexpr int kBarOff = kVOff + kVStages * kHalfBytes;
  static constexpr int kPubOff = kBarOff + 512;                 // m_pub[2][128] + l_pub[2][128] + mfin[2][128] fp32
  static constexpr int kTotal = kPubOff + 6 * kBQ * 4 + 1024;
  static constexpr int kThreads = 384;
};

template <bool TURN>
__global__ void __launch_bounds__(AttnStepCfg::kThreads, 1) attention_kernel_step(const __grid_constant__ AttnParams P,
                                                                                 const __grid_constant__ CUtensorMap tmap_k64) {
  using C = AttnStepCfg;
  constexpr int KK = C::kKStages, KV = C::kVStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* q_full = bars;                 // leader
  uint64_t* k_full = q_full + 1;           // KK, leader
  uint64_t* k_empty = k_full + KK;         // KK, each CTA
  uint64_t* v_full = k_empty + KK;         // KV, leader
  uint64_t* v_empty = v_full + KV;         // KV, each CTA
  uint64_t* s_ready = v_empty + KV;        // 2 (S slot = step parity), each CTA
  uint64_t* p_ready = s_ready + 2;         // 2, leader: 4 warps x 2 CTAs
  uint64_t* p_lo = p_ready + 2;            // 2, leader
  uint64_t* o_done = p_lo + 2;             // 2 (step parity), each CTA
  uint64_t* s_free = o_done + 2;           // 2, leader: S slot read into registers by 4 warps x 2 CTAs
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_free + 2);
  float* m_pub = reinterpret_cast<float*>(smem + C::kPubOff);  // [2][128]
  float* l_pub = m_pub + 2 * kBQ;                              // [2][128] (final exchange)
  float* mf_pub = l_pub + 2 * kBQ;                             // [2][128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const fluxb200_attention_args& a = P.a;
  const int q0 = (blockIdx.x >> 1) * (2 * kBQ) + rank * kBQ;  // this CTA's 128 query rows
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int bh = b * a.H + h;
  const int n = P.num_kv_tiles;
  constexpr uint32_t kSCol = 0, kPCol = 256, kOCol = 384;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tmap_q);
    tma_prefetch_desc(&tmap_k64);
    tma_prefetch_desc(&P.tmap_v);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KK; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < KV; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int w = 0; w < 2; ++w) {
      mbar_init(&s_ready[w], 1);
      mbar_init(&p_ready[w], 8);
      mbar_init(&p_lo[w], 8);
      mbar_init(&o_done[w], 1);
      mbar_init(&s_free[w], 8);
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
  pdl_wait();

  if (warp < 4) {
    reg_dec<88>();
    if (warp == 0) {
      // ---------------- TMA producer ----------------
      if (elect_one()) {
        if (rank == 0) mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
        uint8_t* dst = smem + C::kQOff;
        tma_load_3d_2sm(dst, &P.tmap_q, q_full, 0, q0, bh, kEvictFirst);
        tma_load_3d_2sm(dst + kChunkBytes, &P.tmap_q, q_full, 64, q0, bh, kEvictFirst);
      }
      __syncwarp();
      // load order = the issuer's consumption order: K(0) K(1) K(2), then K(j+3) V(j) for every step j
      auto load_k = [&](int i) {
        const int st = i % KK;
        uint8_t* kd = smem + C::kKOff + st * C::kHalfBytes;
        mbar_wait(&k_empty[st], ((i / KK) & 1) ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&k_full[st], kTileBytes);
          tma_load_3d_2sm(kd, &tmap_k64, &k_full[st], 0, i * kBKV + rank * 64, bh, kEvictLast);
          tma_load_3d_2sm(kd + C::kHalfBytes / 2, &tmap_k64, &k_full[st], 64, i * kBKV + rank * 64, bh, kEvictLast);
        }
        __syncwarp();
      };
      for (int i = 0; i < 3 && i < n; ++i) load_k(i);
      for (int j = 0; j < n; ++j) {
        if (j + 3 < n) load_k(j + 3);
        const int st = j % KV;
        uint8_t* vd = smem + C::kVOff + st * C::kHalfBytes;
        mbar_wait(&v_empty[st], ((j / KV) & 1) ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&v_full[st], kTileBytes);
          tma_load_3d_2sm(vd, &P.tmap_v, &v_full[st], rank * 64, j * kBKV, bh, kEvictLast);
        }
        __syncwarp();
      }
    } else if (warp == 1 && rank == 0 && lane == 0) {
      // ---------------- QK issuer (leader CTA, one thread): S(i) = Q K(i)^T into slot i & 1 ----------------
      // Two issuing threads (this one and the PV issuer below) so that neither the ~45 cycles an MMA costs its issuing
      // thread nor the barrier round trips of one product sit in front of the other product's MMAs.
      constexpr uint32_t idesc_qk = make_idesc(kFmtBF16, kFmtBF16, 2 * kBQ, kBKV, 0, 0);
      const uint64_t q_desc0 = make_desc_sw128(smem_u32(smem + C::kQOff), 16, 1024);
      const uint64_t k_desc0 = make_desc_sw128(smem_u32(smem + C::kKOff), 16, 1024);
#ifdef FLUXB200_ATTN_PROBE
      const bool itr = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#endif
      mbar_wait(q_full, 0);
      for (int i = 0; i < n; ++i) {
        const int st = i % KK, slot = i & 1;
        FB_TRACE(itr, 0, i, 0);
        if (i >= 2) mbar_wait(&s_free[slot], ((i - 2) >> 1) & 1);  // S(i-2) has been pulled into registers
        FB_TRACE(itr, 0, i, 1);
        mbar_wait(&k_full[st], (i / KK) & 1);
        tc_fence_after();
        FB_TRACE(itr, 0, i, 2);
        const uint32_t d = tmem_base + kSCol + slot * 128;
        const uint64_t bd0 = desc_advance(k_desc0, st * C::kHalfBytes);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk) {
          const uint32_t aoff = (kk >> 2) * kChunkBytes + (kk & 3) * 32;
          const uint32_t boff = (kk >> 2) * (C::kHalfBytes / 2) + (kk & 3) * 32;
          mma_f16_ss_2sm(d, desc_advance(q_desc0, aoff), desc_advance(bd0, boff), idesc_qk, kk != 0 ? 1u : 0u);
        }
        tc_commit_2sm(&s_ready[slot], 3);
        tc_commit_2sm(&k_empty[st], 3);
        FB_TRACE(itr, 0, i, 3);
      }
    } else if (warp == 3 && rank == 0 && lane == 0) {
      // ---------------- PV issuer (leader CTA, one thread): O += P(j) V(j), in two K halves ----------------
      constexpr uint32_t idesc_pv = make_idesc(kFmtBF16, kFmtBF16, 2 * kBQ, kD, 0, 1);
      const uint64_t v_desc0 = make_desc_sw128(smem_u32(smem + C::kVOff), kChunkBytes, 1024);
#ifdef FLUXB200_ATTN_PROBE
      const bool itr = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#endif
      for (int j = 0; j < n; ++j) {
        const int w = j & 1, st = j % KV;
        const uint32_t par = (j >> 1) & 1;
        const uint64_t bd0 = desc_advance(v_desc0, st * C::kHalfBytes);
        const uint32_t pcol = tmem_base + kPCol + w * 64;
        mbar_wait(&v_full[st], (j / KV) & 1);
        mbar_wait(&p_lo[w], par);
        tc_fence_after();
        FB_TRACE(itr, 0, j, 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          mma_f16_ts_2sm(tmem_base + kOCol, pcol + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, (j != 0 || kk != 0) ? 1u : 0u);
        FB_TRACE(itr, 0, j, 5);
        mbar_wait(&p_ready[w], par);
        tc_fence_after();
        FB_TRACE(itr, 0, j, 6);
#pragma unroll
        for (int kk = 4; kk < 8; ++kk)
          mma_f16_ts_2sm(tmem_base + kOCol, pcol + kk * 8, desc_advance(bd0, kk * 2048), idesc_pv, 1u);
        tc_commit_2sm(&o_done[w], 3);
        tc_commit_2sm(&v_empty[st], 3);
        FB_TRACE(itr, 0, j, 7);
      }
    }
  } else {
    // ---------------- softmax warpgroups: warpgroup w takes the KV steps j = w, w + 2, ... ----------------
    reg_inc<208>();
    const int w = (warp - 4) >> 2;
    const int lg = warp & 3;
    const int r = lg * 32 + lane;
    const int qrow = q0 + r;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);
    const uint32_t o_taddr = lane_base + kOCol;
    const float sl2 = P.scale_log2;
#ifdef FLUXB200_ATTN_PROBE
    const bool dbg = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
    const bool tr = blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lg == 0 && lane == 0;
#else
    [[maybe_unused]] constexpr bool tr = false;
    constexpr bool dbg = false;
#endif
    unsigned long long d_wait_s = 0, d_ld = 0, d_max = 0, d_exp = 0, d_wait_o = 0, d_st = 0, tA = 0, tB = 0;
    int nsteps = 0;
    if (TURN && w == 1) named_bar_arrive(4, 256);  // the first turn is warpgroup 0's
    float m_mine = -INFINITY;  // reference maximum (log2 units) this warpgroup's row sum is expressed in
    float l = 0.f;
    for (int j = w; j < n; j += 2) {
      const uint32_t par = (j >> 1) & 1;
      if (dbg) tA = clk();
      FB_TRACE(tr, 1 + w, j, 0);
      mbar_wait(&s_ready[w], par);
      tc_fence_after();
      FB_TRACE(tr, 1 + w, j, 1);
      if (dbg) { tB = clk(); d_wait_s += tB - tA; tA = tB; ++nsteps; }
      uint32_t sv[128];
      {
        uint32_t(*sv4)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        tmem_ld32(lane_base + kSCol + w * 128 + 0, sv4[0]);
        tmem_ld32(lane_base + kSCol + w * 128 + 32, sv4[1]);
        tmem_ld32(lane_base + kSCol + w * 128 + 64, sv4[2]);
        tmem_ld32(lane_base + kSCol + w * 128 + 96, sv4[3]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(&s_free[w], 0);  // QK(j+2) may overwrite the slot from here on
      }
      if (dbg) { tB = clk(); d_ld += tB - tA; tA = tB; }
      FB_TRACE(tr, 1 + w, j, 2);
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
      // reference maximum of the previous step (the other warpgroup's), then ours, then hand it on
      float m_prev = -INFINITY;
      if (dbg) { tB = clk(); d_max += tB - tA; tA = tB; }
      FB_TRACE(tr, 1 + w, j, 3);
      if (j > 0) {
        named_bar_sync(1 + ((j - 1) & 1), 256);  // the other warpgroup has published m_ref(j-1)
        m_prev = m_pub[((j - 1) & 1) * kBQ + r];
      }
      if (dbg) { tB = clk(); d_st += tB - tA; tA = tB; }  // wait for the other warpgroup's maximum
      FB_TRACE(tr, 1 + w, j, 4);
      const bool grow = __any_sync(0xffffffffu, m_cand > m_prev + kRescaleThreshold);
      const float m_ref = grow ? fmaxf(m_prev, m_cand) : m_prev;
      if (j + 1 < n) {
        m_pub[(j & 1) * kBQ + r] = m_ref;
        named_bar_arrive(1 + (j & 1), 256);
      }
      if (m_ref != m_mine) {
        l *= fast_exp2(m_mine - m_ref);  // exp2(-inf) = 0 (first step: l = 0 anyway)
        m_mine = m_ref;
      }
      if (j >= 2) {
        // PV(j-2) retired: P[w] may be overwritten.  Also keeps the parity wait on the OTHER slot's barrier below
        // sound: PV(j-3) has retired too (in order), so that barrier is at most one phase behind what we ask for.
        mbar_wait(&o_done[w], par ^ 1);
        tc_fence_after();
      }
      if (grow && j > 0) {
        // O (which holds PV(0..j-1)) moves to the new reference: wait for PV(j-1) to retire, rescale, THEN release P(j)
        mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        tc_fence_after();
        const float alpha = fast_exp2(m_prev - m_ref);
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
      if (dbg) { tB = clk(); d_wait_o += tB - tA; tA = tB; }
      FB_TRACE(tr, 1 + w, j, 5);
      // x = s * scale * log2(e) - m for the whole row BEFORE the turn: FMA-pipe work that runs under the other
      // warpgroup's exponentials instead of inside this warpgroup's exclusive MUFU window (the empty volatile asm pins
      // the results in front of the barrier; ptxas otherwise sinks them between the first MUFU instructions).
      {
        const float2 sl2v = make_float2(sl2, sl2), negmv = make_float2(-m_ref, -m_ref);
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          const float2 t = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, negmv);
          sv[i] = __float_as_uint(t.x);
          sv[i + 1] = __float_as_uint(t.y);
        }
#pragma unroll
        for (int i = 0; i < 128; i += 16)
          asm volatile("" : "+r"(sv[i]), "+r"(sv[i + 1]), "+r"(sv[i + 2]), "+r"(sv[i + 3]), "+r"(sv[i + 4]), "+r"(sv[i + 5]),
                            "+r"(sv[i + 6]), "+r"(sv[i + 7]), "+r"(sv[i + 8]), "+r"(sv[i + 9]), "+r"(sv[i + 10]),
                            "+r"(sv[i + 11]), "+r"(sv[i + 12]), "+r"(sv[i + 13]), "+r"(sv[i + 14]), "+r"(sv[i + 15]));
      }
      // The exponentials of the two warpgroups take turns on the MUFU (named barriers 4 / 5): run concurrently both
      // take twice as long and the warpgroups fall into phase (TMEM load + max of both, then both exponentiate).
      if constexpr (TURN) named_bar_sync(4 + w, 256);
      float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
        for (int i = hh * 64; i < hh * 64 + 64; i += 8) {
          const float p0 = fast_exp2_pinned(__uint_as_float(sv[i])), p1 = fast_exp2_pinned(__uint_as_float(sv[i + 1]));
          const float p2 = fast_exp2_pinned(__uint_as_float(sv[i + 2])), p3 = fast_exp2_pinned(__uint_as_float(sv[i + 3]));
          const float p4 = fast_exp2_pinned(__uint_as_float(sv[i + 4])), p5 = fast_exp2_pinned(__uint_as_float(sv[i + 5]));
          const float p6 = fast_exp2_pinned(__uint_as_float(sv[i + 6])), p7 = fast_exp2_pinned(__uint_as_float(sv[i + 7]));
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
        tmem_st32(lane_base + kPCol + w * 64 + hh * 32, pk[hh]);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote_relaxed(hh == 0 ? &p_lo[w] : &p_ready[w], 0);
        FB_TRACE(tr, 1 + w, j, 6 + hh);
      }
      if constexpr (TURN) named_bar_arrive(4 + (w ^ 1), 256);
      const float2 acc = fadd2(acc0, acc1);
      l += acc.x + acc.y;
      if (dbg) { tB = clk(); d_exp += tB - tA; tA = tB; }
    }
    if (dbg) {
      g_attn_dbg[0] = d_wait_s, g_attn_dbg[1] = d_ld, g_attn_dbg[2] = d_max, g_attn_dbg[3] = d_exp;
      g_attn_dbg[4] = d_wait_o, g_attn_dbg[5] = d_st, g_attn_dbg[6] = nsteps;
    }
    // ---------------- combine the two partial row sums ----------------
    l_pub[w * kBQ + r] = l;
    mf_pub[w * kBQ + r] = m_mine;
    named_bar_sync(3, 256);
    const float l_o = l_pub[(w ^ 1) * kBQ + r], m_o = mf_pub[(w ^ 1) * kBQ + r];
    const float m_fin = fmaxf(m_mine, m_o);  // = the reference maximum of the last step (it never decreases)
    float l_tot = 0.f;
    if (m_mine != -INFINITY) l_tot += l * fast_exp2(m_mine - m_fin);
    if (m_o != -INFINITY) l_tot += l_o * fast_exp2(m_o - m_fin);

    // ---------------- epilogue: warpgroup w writes head-dim columns [64 w, 64 w + 64) ----------------
    mbar_wait(&o_done[(n - 1) & 1], ((n - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_tot;
    const bool valid = qrow < a.S;
    const bool second = a.out1 != nullptr && qrow >= a.split_row;
    void* const outp = second ? a.out1 : a.out;
    const int64_t obase = second ? static_cast<int64_t>(b) * a.out1_batch_stride +
                                       static_cast<int64_t>(qrow - a.split_row) * a.ldo1 + h * kD
                                 : static_cast<int64_t>(b) * a.out_batch_stride + static_cast<int64_t>(qrow) * a.ldo + h * kD;
    float oscale = 1.f;
    if (a.out_kind == 1) oscale = __ldg(qrow < a.split_row ? a.out_scale0 : a.out_scale1);
#pragma unroll 1
    for (int c = w * 2; c < w * 2 + 2; ++c) {
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
          uint32_t wd[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float o = bf16r(__uint_as_float(ov[q * 16 + t * 4 + e]) * inv_l);
              f[e] = a.out_fmt == FLUXB200_E5M2 ? quant_pre<1>(o, oscale) : quant_pre<0>(o, oscale);
            }
            if (a.out_fmt == FLUXB200_E5M2)
              wd[t] = to_fp8x2<1>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<1>(f[2], f[3])) << 16);
            else
              wd[t] = to_fp8x2<0>(f[0], f[1]) | (static_cast<uint32_t>(to_fp8x2<0>(f[2], f[3])) << 16);
          }
          dst[q] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
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

template <bool TURN>
static int launch_attention_step(const AttnParams& P, cudaStream_t stream) {
  using C = AttnStepCfg;
  static_assert(C::kTotal <= 227 * 1024, "attention smem budget");
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA_OK(cudaFuncSetAttribute(attention_kernel_step<TURN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kTotal));
    attr_set = true;
  }
  const fluxb200_attention_args& a = P.a;
  CUtensorMap tmap_k64;
  const uint64_t bhn = static_cast<uint64_t>(a.B) * a.H;
  const uint64_t row_bytes = kD * 2;
  int rc = make_tmap_3d(&tmap_k64, a.k, 2, kD, a.S, bhn, row_bytes, row_bytes * a.S, 64, kBKV / 2, 1);
  if (rc) return rc;
  dim3 grid(2 * ((a.S + 2 * kBQ - 1) / (2 * kBQ)), a.H, a.B);
  FB_CUDA_OK(launch_kernel(attention_kernel_step<TURN>, grid, dim3(C::kThreads), C::kTotal, stream, 2, P, tmap_k64));
  return 0;
}

