// LayerNorm (no affine) -> (1 + scale) * x + shift -> fp8 quantise of ONE row by ONE warp: shared by the stand-alone
// kernel (elementwise.cu, fluxb200_ln_mod_quant*) and the prologue phase of the fused LN -> GEMM launch (f8_gemm.cu,
// fluxb200_f8_gemm_ln).   modules/flux_model.py:367-368 (and 374-375, 389, 395, 469-470) + float8_quantize.py:274-276
#pragma once
#include <type_traits>

#include "ptx.cuh"

namespace fb {

constexpr int kLnMaxIter = 16;  // D <= 16*256 = 4096

// NI = D/256 when known at compile time (12 for hidden 3072: the row lives in 48 registers and four 256-thread
// blocks fit per SM, so the 4608-row launch is a single wave); NI = 0 is the generic predicated form.
struct LnSeg {
  const __nv_bfloat16* x;
  const __nv_bfloat16* shift;
  const __nv_bfloat16* scale;
  uint8_t* yq;
  __nv_bfloat16* yb;
  const float* in_scale;
  int64_t ldx, ldy, ldyb, mod_stride;
  int rows, L;
};
struct LnParams {
  LnSeg seg[2];  // rows of seg[0] come first
  int D;
  float eps;
};

// NI = D / 256 when known at compile time (12 for hidden 3072: the row lives in 48 registers); NI = 0 is the generic
// predicated form.  Called by all 32 lanes of a warp with the same `row` (row < G.rows).
template <int FMT, int NI>
__device__ __forceinline__ void ln_mod_quant_row(const LnSeg& G, int row, int D, float eps, int lane) {
  const __nv_bfloat16* __restrict__ x = G.x;
  const __nv_bfloat16* __restrict__ shift = G.shift;
  const __nv_bfloat16* __restrict__ scale = G.scale;
  uint8_t* __restrict__ yq = G.yq;
  __nv_bfloat16* __restrict__ yb = G.yb;
  const float* __restrict__ in_scale = G.in_scale;
  const int64_t ldx = G.ldx, ldy = G.ldy, ldyb = G.ldyb, mod_stride = G.mod_stride;
  const int L = G.L;
  const int b = row / L;
  const int ni = NI ? NI : D / 256;
  constexpr int kIter = NI ? NI : kLnMaxIter;
  const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 xv[kIter];
  // the bf16 halves enter the fp32 arithmetic directly (FHADD.BF16: no unpack instructions in any of the three passes)
  float sum = 0.f, sum1 = 0.f;  // low / high halves: two independent chains
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    if (NI || i < ni) {
      xv[i] = __ldg(xp + i * 32 + lane);
      uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        sum = add_f32_bf16lo(w[t], sum);
        sum1 = add_f32_bf16hi(w[t], sum1);
      }
    }
  }
  sum += sum1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / D;
  // Opaque no-op on the packed row: stops the compiler from keeping the 96 unpacked fp32 values of the previous
  // pass alive (CSE of the unpack), which costs 2x the registers of the packed row and halves occupancy.
#pragma unroll
  for (int i = 0; i < kIter; ++i)
    asm volatile("" : "+r"(xv[i].x), "+r"(xv[i].y), "+r"(xv[i].z), "+r"(xv[i].w));
  float var = 0.f;
  const float neg_mean = -mean;
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    if (NI || i < ni) {
      uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float d0 = add_f32_bf16lo(w[t], neg_mean), d1 = add_f32_bf16hi(w[t], neg_mean);
        var = fmaf(d0, d0, var);
        var = fmaf(d1, d1, var);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / D + eps);
  // Opaque no-op on the packed row: stops the compiler from keeping the 96 unpacked fp32 values of the previous
  // pass alive (CSE of the unpack), which costs 2x the registers of the packed row and halves occupancy.
#pragma unroll
  for (int i = 0; i < kIter; ++i)
    asm volatile("" : "+r"(xv[i].x), "+r"(xv[i].y), "+r"(xv[i].z), "+r"(xv[i].w));
  const float s = in_scale ? __ldg(in_scale) : 1.f;
  // bf16-in / bf16-out products and sums are done with packed HMUL2/HADD2.BF16: for bf16 operands they round the
  // exact result once, which equals the reference's fp32 op followed by a bf16 rounding (the fp32 intermediate is
  // exact, or differs from the exact value far below half a bf16 ulp).  The quantising multiply can go packed only
  // when the scale itself is a bf16 value (always true under the CUDA scale semantics).
  const bool s_is_bf16 = bf16r(s) == s;
  const __nv_bfloat162 one2 = __floats2bfloat162_rn(1.f, 1.f);
  const __nv_bfloat162 s2 = __floats2bfloat162_rn(s, s);
  const uint4* shp = reinterpret_cast<const uint4*>(shift + static_cast<int64_t>(b) * mod_stride);
  const uint4* scp = reinterpret_cast<const uint4*>(scale + static_cast<int64_t>(b) * mod_stride);
  // one top-level branch on the scale's representability -> two straight-line copies of the loop
  auto apply = [&](auto packed_tag) {
  constexpr bool kPacked = decltype(packed_tag)::value;
#pragma unroll
  for (int i = 0; i < kIter; ++i) {
    if (NI || i < ni) {
      uint4 sh = __ldg(shp + i * 32 + lane);
      uint4 sc = __ldg(scp + i * 32 + lane);
      const uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
      const uint32_t shw[4] = {sh.x, sh.y, sh.z, sh.w};
      const uint32_t scw[4] = {sc.x, sc.y, sc.z, sc.w};
      uint32_t mb[4];   // modulated values as bf16 pairs
      float q[8];       // quantiser inputs (pre-clamp)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const __nv_bfloat162 n2 =
            __floats2bfloat162_rn(add_f32_bf16lo(w[t], neg_mean) * rstd, add_f32_bf16hi(w[t], neg_mean) * rstd);
        const __nv_bfloat162 sc2 = *reinterpret_cast<const __nv_bfloat162*>(&scw[t]);
        const __nv_bfloat162 sh2 = *reinterpret_cast<const __nv_bfloat162*>(&shw[t]);
        // (1 + scale) * ln + shift with eager bf16 rounding after each op
        // (_rn forms: no contraction of the product and the sum into one HFMA2 -- each op rounds, as eager torch does)
        const __nv_bfloat162 m2 = __hadd2_rn(__hmul2_rn(__hadd2_rn(one2, sc2), n2), sh2);
        mb[t] = *reinterpret_cast<const uint32_t*>(&m2);
        if (kPacked) {
          const __nv_bfloat162 p2 = __hmul2_rn(m2, s2);
          const float2 pf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&p2));
          q[t * 2] = pf.x, q[t * 2 + 1] = pf.y;
        } else {
          const float2 mf = unpack_bf16x2(mb[t]);
          q[t * 2] = bf16r(mf.x * s), q[t * 2 + 1] = bf16r(mf.y * s);
        }
      }
      const int64_t col = static_cast<int64_t>(i * 32 + lane) * 8;
      if (NI == 0 && yb) *reinterpret_cast<uint4*>(yb + static_cast<int64_t>(row) * ldyb + col) = make_uint4(mb[0], mb[1], mb[2], mb[3]);
      if (NI || yq) {
        // clamp(+-max) then cast == saturating cast
        uint2 o;
        o.x = to_fp8x2<FMT>(q[0], q[1]) | (static_cast<uint32_t>(to_fp8x2<FMT>(q[2], q[3])) << 16);
        o.y = to_fp8x2<FMT>(q[4], q[5]) | (static_cast<uint32_t>(to_fp8x2<FMT>(q[6], q[7])) << 16);
        *reinterpret_cast<uint2*>(yq + static_cast<int64_t>(row) * ldy + col) = o;
      }
    }
  }
  };
  if (s_is_bf16) apply(std::true_type{});
  else apply(std::false_type{});
}

}  // namespace fb
