// Text-encoder kernels (second half): the arithmetic of the Hugging Face `T5EncoderModel` and
// `CLIPTextModel` that the reference's `modules/conditioner.py:HFEmbedder` wraps (:80-92 construction, :101-117 forward),
// as far as it is not a dense linear layer (those run on the tcgen05 implicit-GEMM kernel of vae.cu in its dense form).
//
//   rows_norm_kernel       T5LayerNorm (RMS, no bias; transformers/models/t5/modeling_t5.py `T5LayerNorm.forward`) and
//                          nn.LayerNorm with affine (CLIP), one warp per row, fp32 statistics.
//   gated_act_kernel       T5DenseGatedActDense's `gelu_new(wi_0 x) * (wi_1 x)` and CLIP's quick_gelu.
//   attention_d64_kernel   multi-head attention with head dim 64: T5 (no scaling, additive relative-position bias shared
//                          by all layers) and CLIP (scale 1/8, causal).  FlashAttention-2 style on mma.sync.m16n8k16:
//                          these are ~0.1 TFLOP per encoder, latency-bound shapes (S <= 512, 64 heads x 8 query tiles =
//                          512 CTAs); the tcgen05 kernels of this library are built for the 30-TFLOP attention of the
//                          denoise step, not for this.
#include <cuda_bf16.h>

#include "host_util.h"
#include "ptx.cuh"

namespace fb {

// y[r, :] = bf16( w * bf16( x * rsqrt(mean(x^2) + eps) ) )                      (bias == nullptr: T5LayerNorm)
// y[r, :] = bf16( (x - mean) * rsqrt(var + eps) * w + b )                        (nn.LayerNorm, biased variance)
constexpr int kNormRowsPerBlock = 2;  // few hundred rows in all: small blocks so that every SM gets some
__global__ void __launch_bounds__(kNormRowsPerBlock * 32) rows_norm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                        const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ y, int64_t ldy, int rows, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * kNormRowsPerBlock + warp;
  if (r >= rows) return;
  const __nv_bfloat16* xr = x + static_cast<int64_t>(r) * ldx;
  __nv_bfloat16* yr = y + static_cast<int64_t>(r) * ldy;
  float s = 0.f, q = 0.f;
  for (int i = lane * 8; i < D; i += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack_bf16x2(wd[j]);
      s += f.x + f.y;
      q += f.x * f.x + f.y * f.y;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const bool rms = bias == nullptr;
  const float mean = rms ? 0.f : s / D;
  const float var = rms ? q / D : fmaxf(q / D - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  for (int i = lane * 8; i < D; i += 256) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + i));
    const uint32_t xd[4] = {v.x, v.y, v.z, v.w}, wd[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t bd[4] = {0, 0, 0, 0};
    if (!rms) {
      const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + i));
      bd[0] = bv.x, bd[1] = bv.y, bd[2] = bv.z, bd[3] = bv.w;
    }
    uint32_t od[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 xf = unpack_bf16x2(xd[j]), wf = unpack_bf16x2(wd[j]);
      if (rms) {
        od[j] = pack_bf16x2(wf.x * bf16r(xf.x * rstd), wf.y * bf16r(xf.y * rstd));
      } else {
        const float2 bf = unpack_bf16x2(bd[j]);
        od[j] = pack_bf16x2((xf.x - mean) * rstd * wf.x + bf.x, (xf.y - mean) * rstd * wf.y + bf.y);
      }
    }
    *reinterpret_cast<uint4*>(yr + i) = make_uint4(od[0], od[1], od[2], od[3]);
  }
}

__device__ __forceinline__ float gelu_new_f32(float x) {  // transformers NewGELUActivation
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}

// mode 0: out[r, c] = bf16( bf16(gelu_new(in[r, c])) * in[r, F + c] )     T5DenseGatedActDense (wi_0 | wi_1 fused along N)
// mode 1: out[r, c] = bf16( in[r, c] * sigmoid(1.702 in[r, c]) )           CLIP quick_gelu
__global__ void __launch_bounds__(256) gated_act_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in,
                                                        __nv_bfloat16* __restrict__ out, int64_t ld_out, int rows, int F, int mode) {
  const int vec_per_row = F >> 3;
  const int64_t total = static_cast<int64_t>(rows) * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int r = static_cast<int>(i / vec_per_row), c = static_cast<int>(i % vec_per_row) * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(in + static_cast<int64_t>(r) * ld_in + c);
    const uint32_t ad[4] = {a.x, a.y, a.z, a.w};
    uint32_t od[4];
    if (mode == 0) {
      const uint4 b = *reinterpret_cast<const uint4*>(in + static_cast<int64_t>(r) * ld_in + F + c);
      const uint32_t bd[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 af = unpack_bf16x2(ad[j]), bf = unpack_bf16x2(bd[j]);
        od[j] = pack_bf16x2(bf16r(gelu_new_f32(af.x)) * bf.x, bf16r(gelu_new_f32(af.y)) * bf.y);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 af = unpack_bf16x2(ad[j]);
        od[j] = pack_bf16x2(af.x / (1.f + __expf(-1.702f * af.x)), af.y / (1.f + __expf(-1.702f * af.y)));
      }
    }
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * ld_out + c) = make_uint4(od[0], od[1], od[2], od[3]);
  }
}

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                               uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// One CTA = one (batch, head, 64-query tile); 4 warps x 16 query rows; keys in chunks of 64 with an online softmax.
// scores = bf16(q k^T) [* scale -> bf16] [+ bias -> bf16] as the eager Hugging Face modules round them, softmax in fp32,
// probabilities to bf16, out = bf16( P v / l ).
constexpr int kA64Pitch = 72;  // 64 + 8 bf16: conflict-free fragment reads
__global__ void __launch_bounds__(128) attention_d64_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                            const __nv_bfloat16* __restrict__ v, int64_t ld,
                                                            const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                            int64_t ldo, int H, int S, float scale, int causal) {
  __shared__ __align__(16) __nv_bfloat16 Qs[64][kA64Pitch];
  __shared__ __align__(16) __nv_bfloat16 Ks[64][kA64Pitch];   // [key][d]
  __shared__ __align__(16) __nv_bfloat16 Vs[64][kA64Pitch];   // [key][d]: the P V product reads it through ldmatrix.trans
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const int64_t row_base = static_cast<int64_t>(b) * S;
  // stage Q (rows past S read as zero)
  for (int i = threadIdx.x; i < 64 * 8; i += 128) {
    const int r = i >> 3, c = (i & 7) * 8;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (q0 + r < S) val = *reinterpret_cast<const uint4*>(q + (row_base + q0 + r) * ld + h * 64 + c);
    *reinterpret_cast<uint4*>(&Qs[r][c]) = val;
  }
  __syncthreads();
  uint32_t qa[4][4];
  const int wr = warp * 16;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    qa[kk][0] = *reinterpret_cast<const uint32_t*>(&Qs[wr + g][kk * 16 + 2 * t]);
    qa[kk][1] = *reinterpret_cast<const uint32_t*>(&Qs[wr + g + 8][kk * 16 + 2 * t]);
    qa[kk][2] = *reinterpret_cast<const uint32_t*>(&Qs[wr + g][kk * 16 + 2 * t + 8]);
    qa[kk][3] = *reinterpret_cast<const uint32_t*>(&Qs[wr + g + 8][kk * 16 + 2 * t + 8]);
  }
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows g and g + 8 of this warp's 16
  const int qr0 = q0 + wr + g, qr1 = qr0 + 8;
  constexpr float kLog2e = 1.4426950408889634f;
  const int k_end = causal ? (q0 + 64 < S ? q0 + 64 : S) : S;
  for (int k0 = 0; k0 < k_end; k0 += 64) {
    __syncthreads();  // previous chunk fully consumed
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {
      const int r = i >> 3, c = (i & 7) * 8;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (k0 + r < S) {
        kv = *reinterpret_cast<const uint4*>(k + (row_base + k0 + r) * ld + h * 64 + c);
        vv = *reinterpret_cast<const uint4*>(v + (row_base + k0 + r) * ld + h * 64 + c);
      }
      *reinterpret_cast<uint4*>(&Ks[r][c]) = kv;
      *reinterpret_cast<uint4*>(&Vs[r][c]) = vv;
    }
    __syncthreads();
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ks[j * 8 + g][kk * 16 + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ks[j * 8 + g][kk * 16 + 2 * t + 8]);
        mma_bf16_16816(s[j], qa[kk][0], qa[kk][1], qa[kk][2], qa[kk][3], b0, b1);
      }
    }
    // roundings of the eager modules, bias, masks
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = k0 + j * 8 + 2 * t + (e & 1);
        const int qr = (e < 2) ? qr0 : qr1;
        float x = bf16r(s[j][e]);
        if (scale != 1.f) x = bf16r(x * scale);
        if (bias != nullptr && qr < S && key < S)
          x = bf16r(x + __bfloat162float(bias[(static_cast<int64_t>(h) * S + qr) * S + key]));
        if (key >= S || (causal && key > qr)) x = -INFINITY;
        s[j][e] = x;
        if (e < 2) mx0 = fmaxf(mx0, x); else mx1 = fmaxf(mx1, x);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = (mn0 == -INFINITY) ? 1.f : exp2f((m0 - mn0) * kLog2e);
    const float a1 = (mn1 == -INFINITY) ? 1.f : exp2f((m1 - mn1) * kLog2e);
    m0 = mn0, m1 = mn1;
    float r0 = 0.f, r1 = 0.f;
    uint32_t pa[8][2];  // probabilities as bf16 pairs: [n-tile][row g | row g + 8]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p00 = (s[j][0] == -INFINITY) ? 0.f : exp2f((s[j][0] - m0) * kLog2e);
      const float p01 = (s[j][1] == -INFINITY) ? 0.f : exp2f((s[j][1] - m0) * kLog2e);
      const float p10 = (s[j][2] == -INFINITY) ? 0.f : exp2f((s[j][2] - m1) * kLog2e);
      const float p11 = (s[j][3] == -INFINITY) ? 0.f : exp2f((s[j][3] - m1) * kLog2e);
      r0 += p00 + p01;
      r1 += p10 + p11;
      pa[j][0] = pack_bf16x2(p00, p01);
      pa[j][1] = pack_bf16x2(p10, p11);
    }
    l0 = l0 * a0 + r0;
    l1 = l1 * a1 + r1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= a0, o[j][1] *= a0, o[j][2] *= a1, o[j][3] *= a1;
    }
    // O += P V: k-step i covers keys [16 i, 16 i + 16) = score n-tiles 2 i and 2 i + 1; n-tile j of the output = d [8 j, 8 j + 8).
    // The B fragment of mma.m16n8k16 wants {V[k][n], V[k+1][n]} pairs (k = key, n = d): ldmatrix.trans hands them out
    // straight from the row-major [key][d] tile, four 8x8 matrices (two k halves x two n-tiles) per instruction.
    const int lm_row = (lane & 7) + ((lane >> 3) & 1) * 8, lm_col = (lane >> 4) * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        uint32_t b00, b01, b10, b11;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(b00), "=r"(b01), "=r"(b10), "=r"(b11)
                     : "r"(smem_u32(&Vs[i * 16 + lm_row][j * 8 + lm_col])));
        mma_bf16_16816(o[j], pa[2 * i][0], pa[2 * i][1], pa[2 * i + 1][0], pa[2 * i + 1][1], b00, b01);
        mma_bf16_16816(o[j + 1], pa[2 * i][0], pa[2 * i][1], pa[2 * i + 1][0], pa[2 * i + 1][1], b10, b11);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (qr0 < S)
      *reinterpret_cast<uint32_t*>(out + (row_base + qr0) * ldo + h * 64 + j * 8 + 2 * t) = pack_bf16x2(o[j][0] * i0, o[j][1] * i0);
    if (qr1 < S)
      *reinterpret_cast<uint32_t*>(out + (row_base + qr1) * ldo + h * 64 + j * 8 + 2 * t) = pack_bf16x2(o[j][2] * i1, o[j][3] * i1);
  }
}

}  // namespace fb

extern "C" {

int fluxb200_rows_norm(const void* x_bf16, int64_t ldx, const void* weight_bf16, const void* bias_bf16, void* y_bf16, int64_t ldy,
                       int rows, int D, float eps, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x_bf16 && weight_bf16 && y_bf16, "fluxb200_rows_norm: null operand");
  FB_REQUIRE(rows > 0 && D > 0 && D % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "fluxb200_rows_norm: D, ldx, ldy must be multiples of 8");
  FB_REQUIRE(((reinterpret_cast<uintptr_t>(x_bf16) | reinterpret_cast<uintptr_t>(y_bf16) | reinterpret_cast<uintptr_t>(weight_bf16) |
               reinterpret_cast<uintptr_t>(bias_bf16)) & 15) == 0, "fluxb200_rows_norm: operands must be 16-byte aligned");
  rows_norm_kernel<<<(rows + kNormRowsPerBlock - 1) / kNormRowsPerBlock, kNormRowsPerBlock * 32, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x_bf16), ldx,
                                                       reinterpret_cast<const __nv_bfloat16*>(weight_bf16),
                                                       reinterpret_cast<const __nv_bfloat16*>(bias_bf16),
                                                       reinterpret_cast<__nv_bfloat16*>(y_bf16), ldy, rows, D, eps);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

int fluxb200_gated_act(const void* in_bf16, int64_t ld_in, void* out_bf16, int64_t ld_out, int rows, int F, int mode,
                       fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(in_bf16 && out_bf16 && rows > 0 && F > 0 && F % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && (mode == 0 || mode == 1),
             "fluxb200_gated_act: bad arguments (F, ld_in, ld_out multiples of 8; mode 0 | 1)");
  FB_REQUIRE(((reinterpret_cast<uintptr_t>(in_bf16) | reinterpret_cast<uintptr_t>(out_bf16)) & 15) == 0,
             "fluxb200_gated_act: operands must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(rows) * (F / 8);
  const int64_t want = (total + 255) / 256;
  const int blocks = static_cast<int>(want < sm_count() * 8 ? want : sm_count() * 8);
  gated_act_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in_bf16), ld_in,
                                               reinterpret_cast<__nv_bfloat16*>(out_bf16), ld_out, rows, F, mode);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

int fluxb200_attention_d64(const void* q_bf16, const void* k_bf16, const void* v_bf16, int64_t ld, const void* bias_bf16,
                           void* out_bf16, int64_t ldo, int B, int H, int S, float scale, int causal, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(q_bf16 && k_bf16 && v_bf16 && out_bf16, "fluxb200_attention_d64: null operand");
  FB_REQUIRE(B > 0 && H > 0 && S > 0 && ld % 8 == 0 && ldo % 2 == 0 && ld >= 64 * H && ldo >= 64 * H,
             "fluxb200_attention_d64: bad geometry (row strides must cover H * 64 and be multiples of 8)");
  FB_REQUIRE(((reinterpret_cast<uintptr_t>(q_bf16) | reinterpret_cast<uintptr_t>(k_bf16) | reinterpret_cast<uintptr_t>(v_bf16)) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out_bf16) & 3) == 0,
             "fluxb200_attention_d64: q / k / v must be 16-byte aligned");
  dim3 grid((S + 63) / 64, H, B);
  attention_d64_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(q_bf16), reinterpret_cast<const __nv_bfloat16*>(k_bf16),
                                                 reinterpret_cast<const __nv_bfloat16*>(v_bf16), ld,
                                                 reinterpret_cast<const __nv_bfloat16*>(bias_bf16),
                                                 reinterpret_cast<__nv_bfloat16*>(out_bf16), ldo, H, S, scale, causal);
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
