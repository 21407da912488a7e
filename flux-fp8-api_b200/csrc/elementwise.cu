// HBM-bound kernels of the hot path: input quantisation, calibration amax, SiLU+quantise,
// LayerNorm+modulate+quantise, stand-alone QKNorm+RoPE and the skinny-M fp8 GEMV used by Modulation.
// All reproduce the reference's eager bf16 rounding points.
#include "flux_b200.h"
#include "host_util.h"
#include <cstdlib>
#include <type_traits>

#include "ln_row.cuh"
#include "ptx.cuh"

namespace fb {

// ---------------------------------------------------------------------------------------------
// quantize: y = fp8(clamp(bf16(x*s)))            float8_quantize.py:217-218, 274-276
// ---------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) quantize_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ y,
                                                       int64_t n, const float* __restrict__ scale) {
  pdl_wait();
  const float s = __ldg(scale);
  const int64_t nvec = n / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    uint4 in = __ldg(reinterpret_cast<const uint4*>(x) + i);
    uint32_t w[4] = {in.x, in.y, in.z, in.w};
    uint16_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = unpack_bf16x2(w[t]);
      o[t] = to_fp8x2<FMT>(quant_pre<FMT>(f.x, s), quant_pre<FMT>(f.y, s));
    }
    uint2 out;
    out.x = o[0] | (static_cast<uint32_t>(o[1]) << 16);
    out.y = o[2] | (static_cast<uint32_t>(o[3]) << 16);
    reinterpret_cast<uint2*>(y)[i] = out;
  }
  // tail
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    int64_t i = nvec * 8 + threadIdx.x;
    y[i] = to_fp8<FMT>(quant_pre<FMT>(__bfloat162float(x[i]), s));
  }
}

// ---------------------------------------------------------------------------------------------
// amax: *amax = max(*amax, max|x|)               float8_quantize.py:197, 227
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) amax_kernel(const __nv_bfloat16* __restrict__ x, int64_t n,
                                                   float* __restrict__ amax) {
  pdl_wait();
  float m = 0.f;
  const int64_t nvec = n / 8;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    uint4 in = __ldg(reinterpret_cast<const uint4*>(x) + i);
    uint32_t w[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float2 f = unpack_bf16x2(w[t]);
      m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) m = fmaxf(m, fabsf(__bfloat162float(x[nvec * 8 + threadIdx.x])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    // non-negative floats order like their bit patterns (NaN propagates as a large pattern)
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(m));
  }
}

// ---------------------------------------------------------------------------------------------
// silu + quantise                                 modules/flux_model.py:249,252
// ---------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) silu_quant_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ y,
                                                         __nv_bfloat16* __restrict__ yb, int64_t n,
                                                         const float* __restrict__ scale) {
  pdl_wait();
  const float s = scale ? __ldg(scale) : 1.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = __bfloat162float(x[i]);
    float a = bf16r(v / (1.f + expf(-v)));
    if (yb) yb[i] = __float2bfloat16_rn(a);
    if (y) y[i] = to_fp8<FMT>(quant_pre<FMT>(a, s));
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (no affine) -> (1+scale)*x + shift -> quantise.  One warp per row.
//   modules/flux_model.py:367-368 (and 374-375, 389, 395, 469-470)
// ---------------------------------------------------------------------------------------------
// (LnSeg / LnParams / ln_mod_quant_row: ln_row.cuh)
template <int FMT, int NI>
__global__ void __launch_bounds__(256, NI == 0 ? 2 : 4) ln_mod_quant_kernel(const __grid_constant__ LnParams P) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // the second row set starts on a block boundary (the grid is padded, not the rows)
  const int blocks0 = (P.seg[0].rows + 7) >> 3;
  const bool second = static_cast<int>(blockIdx.x) >= blocks0;
  const LnSeg& G = second ? P.seg[1] : P.seg[0];
  const int row = (static_cast<int>(blockIdx.x) - (second ? blocks0 : 0)) * 8 + warp;
  if (row >= G.rows) return;
  ln_mod_quant_row<FMT, NI>(G, row, P.D, P.eps, lane);
}

// ---------------------------------------------------------------------------------------------
// stand-alone QKNorm + RoPE on [B,H,S,128]; one warp per row, lane owns 4 consecutive elements
//   modules/flux_model.py:164 (RMSNorm), 60-65 (apply_rope)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qknorm_rope_kernel(const __nv_bfloat16* __restrict__ x,
                                                          __nv_bfloat16* __restrict__ y,
                                                          const float* __restrict__ norm_w,
                                                          const __nv_bfloat16* __restrict__ rcos,
                                                          const __nv_bfloat16* __restrict__ rsin, int64_t rope_bstride,
                                                          int64_t rows, int H, int S, float eps) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + warp;
  if (row >= rows) return;
  const int64_t s = row % S;
  const int64_t b = row / (static_cast<int64_t>(S) * H);
  uint2 in = *reinterpret_cast<const uint2*>(x + row * 128 + lane * 4);
  float2 a = unpack_bf16x2(in.x), c = unpack_bf16x2(in.y);
  float v[4] = {a.x, a.y, c.x, c.y};
  if (norm_w) {
    float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rinv = rsqrtf(ss * (1.f / 128.f) + eps);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = bf16r(v[j] * rinv * __ldg(norm_w + lane * 4 + j));
  }
  if (rcos && rsin) {
    const int64_t off = b * rope_bstride + s * 64 + lane * 2;
    float2 cs = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(rcos + off));
    float2 sn = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(rsin + off));
    float o0 = bf16r(cs.x * v[0]) + bf16r(-sn.x * v[1]);
    float o1 = bf16r(sn.x * v[0]) + bf16r(cs.x * v[1]);
    float o2 = bf16r(cs.y * v[2]) + bf16r(-sn.y * v[3]);
    float o3 = bf16r(sn.y * v[2]) + bf16r(cs.y * v[3]);
    v[0] = o0, v[1] = o1, v[2] = o2, v[3] = o3;
  }
  uint2 out;
  out.x = pack_bf16x2(v[0], v[1]);
  out.y = pack_bf16x2(v[2], v[3]);
  *reinterpret_cast<uint2*>(y + row * 128 + lane * 4) = out;
}

// ---------------------------------------------------------------------------------------------
// skinny-M fp8 GEMV: out[m,n] = bf16(acc*s + bias).  Weight-streaming; A cached in smem as fp16
// (exact for both fp8 formats); products and sums in fp32 (exact products, like the tensor core).
// ---------------------------------------------------------------------------------------------
constexpr int kGemvColsPerWarp = 4;
constexpr int kGemvWarps = 8;
constexpr int kGemvMT = 4;  // rows of A handled per pass

template <int FMT>
__device__ __forceinline__ void fp8x16_to_float(const uint4& in, float (&f)[16]) {
  const uint32_t w[4] = {in.x, in.y, in.z, in.w};
#pragma unroll
  for (int t = 0; t < 4; ++t) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>((w[t] >> (16 * h)) & 0xffff),
                                                  FMT == 0 ? __NV_E4M3 : __NV_E5M2);
      float2 ff = __half22float2(*reinterpret_cast<__half2*>(&hr));
      f[t * 4 + h * 2 + 0] = ff.x;
      f[t * 4 + h * 2 + 1] = ff.y;
    }
  }
}

template <int AFMT, int WFMT>
__global__ void __launch_bounds__(kGemvWarps * 32) f8_gemv_kernel(const uint8_t* __restrict__ a,
                                                                  const uint8_t* __restrict__ w,
                                                                  const __nv_bfloat16* __restrict__ bias,
                                                                  const float* __restrict__ sa,
                                                                  const float* __restrict__ sw,
                                                                  __nv_bfloat16* __restrict__ out, int M, int N, int K) {
  pdl_wait();
  extern __shared__ __half a_sm[];  // [kGemvMT][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float s = __ldg(sa) * __ldg(sw);
  const int n0 = (blockIdx.x * kGemvWarps + warp) * kGemvColsPerWarp;
  for (int m0 = 0; m0 < M; m0 += kGemvMT) {
    const int mt = min(kGemvMT, M - m0);
    __syncthreads();
    for (int i = threadIdx.x; i < mt * K / 2; i += blockDim.x) {
      const int m = (i * 2) / K, k = (i * 2) % K;
      uint16_t two = *reinterpret_cast<const uint16_t*>(a + static_cast<int64_t>(m0 + m) * K + k);
      __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(two, AFMT == 0 ? __NV_E4M3 : __NV_E5M2);
      *reinterpret_cast<__half2*>(a_sm + m * K + k) = *reinterpret_cast<__half2*>(&hr);
    }
    __syncthreads();
    if (n0 < N) {
      float acc[kGemvColsPerWarp][kGemvMT];
#pragma unroll
      for (int c = 0; c < kGemvColsPerWarp; ++c)
#pragma unroll
        for (int m = 0; m < kGemvMT; ++m) acc[c][m] = 0.f;
      for (int k = lane * 16; k < K; k += 32 * 16) {
        float af[kGemvMT][16];
#pragma unroll
        for (int m = 0; m < kGemvMT; ++m) {
          if (m < mt) {
            const uint4* ap = reinterpret_cast<const uint4*>(a_sm + m * K + k);
            uint4 lo = ap[0], hi = ap[1];
            const __half2* h = reinterpret_cast<const __half2*>(&lo);
            const __half2* h2 = reinterpret_cast<const __half2*>(&hi);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 f0 = __half22float2(h[t]), f1 = __half22float2(h2[t]);
              af[m][t * 2] = f0.x, af[m][t * 2 + 1] = f0.y;
              af[m][8 + t * 2] = f1.x, af[m][8 + t * 2 + 1] = f1.y;
            }
          }
        }
#pragma unroll
        for (int c = 0; c < kGemvColsPerWarp; ++c) {
          if (n0 + c < N) {
            uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + static_cast<int64_t>(n0 + c) * K + k));
            float wf[16];
            fp8x16_to_float<WFMT>(wv, wf);
#pragma unroll
            for (int m = 0; m < kGemvMT; ++m)
              if (m < mt) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c][m] = fmaf(wf[j], af[m][j], acc[c][m]);
              }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kGemvColsPerWarp; ++c)
#pragma unroll
        for (int m = 0; m < kGemvMT; ++m) {
          float v = acc[c][m];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && m < mt && n0 + c < N) {
            float bb = bias ? __bfloat162float(bias[n0 + c]) : 0.f;
            out[static_cast<int64_t>(m0 + m) * N + n0 + c] = __float2bfloat16_rn(fmaf(v, s, bb));
          }
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// batched modulation: (1) per-layer silu + quantise of the shared `vec`, (2) one weight-streaming GEMV over
// every layer's rows.  Each warp owns 4 weight rows at a time and issues all of a row group's 16-byte loads
// before consuming them (K = 3072: 24 loads in flight per lane).
// ---------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) silu_quant_layers_kernel(const __nv_bfloat16* __restrict__ vec,
                                                                const fluxb200_gemv_layer* __restrict__ layers,
                                                                uint8_t* __restrict__ aq, int BK) {
  pdl_wait();
  const float s = __ldg(layers[blockIdx.x].in_qscale);
  uint8_t* dst = aq + static_cast<int64_t>(blockIdx.x) * BK;
  for (int i = threadIdx.x; i < BK; i += blockDim.x) {
    float v = __bfloat162float(vec[i]);
    float a = bf16r(v / (1.f + expf(-v)));
    dst[i] = to_fp8<FMT>(quant_pre<FMT>(a, s));
  }
}

constexpr int kModColsPerBlock = 64;
constexpr int kModMT = 2;    // rows of the (tiny) batch handled per pass
constexpr int kModCols = 2;  // weight rows a warp streams at a time

template <int AFMT, int WFMT, int NK>  // NK = ceil(K / 512): 16-byte chunks per lane per weight row
__global__ void __launch_bounds__(256, 2) gemv_layers_kernel(const fluxb200_gemv_layer* __restrict__ layers,
                                                             int num_layers, const uint8_t* __restrict__ aq,
                                                             __nv_bfloat16* __restrict__ out, int64_t ld_out, int B,
                                                             int K) {
  pdl_wait();
  __shared__ __align__(16) __half a_sm[kModMT * NK * 512];
  __shared__ int s_layer;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    int l = 0;
    while (l + 1 < num_layers && layers[l + 1].block_start <= static_cast<int>(blockIdx.x)) ++l;
    s_layer = l;
  }
  __syncthreads();
  const int l = s_layer;
  const fluxb200_gemv_layer L = layers[l];
  const float s = __ldg(L.a_scale_recip) * __ldg(L.w_scale_recip);
  const uint8_t* w = static_cast<const uint8_t*>(L.w);
  const __nv_bfloat16* bias = static_cast<const __nv_bfloat16*>(L.bias);
  const int nblk = (static_cast<int>(blockIdx.x) - L.block_start) * kModColsPerBlock;

  for (int m0 = 0; m0 < B; m0 += kModMT) {
    const int mt = min(kModMT, B - m0);
    __syncthreads();
    const uint8_t* asrc = aq + (static_cast<int64_t>(l) * B + m0) * K;
    for (int i = threadIdx.x; i < mt * K / 2; i += blockDim.x) {
      uint16_t two = *reinterpret_cast<const uint16_t*>(asrc + i * 2);
      __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(two, AFMT == 0 ? __NV_E4M3 : __NV_E5M2);
      *reinterpret_cast<__half2*>(a_sm + i * 2) = *reinterpret_cast<__half2*>(&hr);
    }
    __syncthreads();
#pragma unroll 1
    for (int grp = 0; grp < kModColsPerBlock / (8 * kModCols); ++grp) {
      const int n0 = nblk + (grp * 8 + warp) * kModCols;
      if (n0 >= L.N) continue;
      uint4 wv[kModCols][NK];
#pragma unroll
      for (int c = 0; c < kModCols; ++c)
#pragma unroll
        for (int i = 0; i < NK; ++i)
          wv[c][i] = (n0 + c < L.N && (i * 32 + lane) * 16 < K)
                         ? __ldg(reinterpret_cast<const uint4*>(w + static_cast<int64_t>(n0 + c) * K + (i * 32 + lane) * 16))
                         : make_uint4(0, 0, 0, 0);
      float acc[kModCols][kModMT];
#pragma unroll
      for (int c = 0; c < kModCols; ++c)
#pragma unroll
        for (int m = 0; m < kModMT; ++m) acc[c][m] = 0.f;
#pragma unroll
      for (int i = 0; i < NK; ++i) {
        const int k = (i * 32 + lane) * 16;
        if (k >= K) continue;  // zero weights were loaded; skip the smem read beyond K
        float wf[kModCols][16];
#pragma unroll
        for (int c = 0; c < kModCols; ++c) fp8x16_to_float<WFMT>(wv[c][i], wf[c]);
#pragma unroll
        for (int m = 0; m < kModMT; ++m) {
          if (m < mt) {
            const uint4* ap = reinterpret_cast<const uint4*>(a_sm + m * K + k);
            uint4 lo = ap[0], hi = ap[1];
            const __half2* h = reinterpret_cast<const __half2*>(&lo);
            const __half2* h2 = reinterpret_cast<const __half2*>(&hi);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float2 f0 = __half22float2(h[t]), f1 = __half22float2(h2[t]);
#pragma unroll
              for (int c = 0; c < kModCols; ++c) {
                acc[c][m] = fmaf(wf[c][t * 2], f0.x, acc[c][m]);
                acc[c][m] = fmaf(wf[c][t * 2 + 1], f0.y, acc[c][m]);
                acc[c][m] = fmaf(wf[c][8 + t * 2], f1.x, acc[c][m]);
                acc[c][m] = fmaf(wf[c][8 + t * 2 + 1], f1.y, acc[c][m]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < kModCols; ++c)
#pragma unroll
        for (int m = 0; m < kModMT; ++m) {
          float v = acc[c][m];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0 && m < mt && n0 + c < L.N) {
            float bb = bias ? __bfloat162float(bias[n0 + c]) : 0.f;
            out[static_cast<int64_t>(m0 + m) * ld_out + L.out_offset + n0 + c] = __float2bfloat16_rn(fmaf(v, s, bb));
          }
        }
    }
  }
}

// Tensor-core form of the batched modulation GEMV (the default when K % 64 == 0).  The CUDA-core form above spends
// ~45 instructions per 16 weight bytes on fp8->fp32 conversion and FMAs and tops out near 3 TB/s.  Here the
// WEIGHTS are the 16-row A operand of mma.sync.m16n8k16 (f16 operands, fp32 accumulate) and the batch (<= 8 rows
// per pass, zero padded) is the 8-column B operand: a lane loads 16 bytes from weight row g and 16 from row g+8,
// F2FP-unpacks them (exact: every e4m3 value is an f16 value, and products of two fp8 values are exact in fp32)
// straight into the a0..a3 registers of four MMAs, and reads the matching activations (converted to f16 once per
// block in shared memory) as two 16-byte pieces that are the b0/b1 pairs of those MMAs: ~12 instructions per 16
// weight bytes, no register shuffling.  (mma.sync with fp8 operand TYPES is not native on sm_100a: ptxas expands
// it into conversions + HMMA + FADDs.)
// K order inside a dot product is free, so K is consumed in 64-byte chunks permuted so that every lane's load is
// one contiguous 16-byte piece: quad lane t owns bytes [16t, 16t+16) of the chunk, word j of it feeds MMA j (low
// half = k-slots 2t,2t+1, high half = k-slots 8+2t,8+2t+1), and the activation fragment is read from the same k's.
// One warp = 16 output columns over the whole K; a block of 4 warps = 64 columns (same block->layer table).
__device__ __forceinline__ void mma_m16n8k16_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                 uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int FMT>
__device__ __forceinline__ uint32_t fp8x2_to_f16x2(uint16_t two) {
  __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2(two, FMT == 0 ? __NV_E4M3 : __NV_E5M2);
  return *reinterpret_cast<uint32_t*>(&hr);
}

constexpr int kModAPad = 32;     // smem row pitch 2K + 32 bytes: the 8 batch rows of a quad-column land in distinct banks
constexpr int kModMmaWarps = 4;  // 4 warps x 16 weight rows = kModColsPerBlock
constexpr int kModU = 4;         // chunks per pipeline group

template <int AFMT>
__global__ void __launch_bounds__(kModMmaWarps * 32, 4) gemv_layers_mma_kernel(
    const fluxb200_gemv_layer* __restrict__ layers, int num_layers, const uint8_t* __restrict__ aq,
    __nv_bfloat16* __restrict__ out, int64_t ld_out, int B, int K) {
  pdl_wait();
  extern __shared__ __align__(16) uint8_t a_sm8[];
  __shared__ int s_layer;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // block -> layer: every thread tests a few table entries, one barrier-wide count (a serial scan by one thread
  // costs one dependent L2 round trip per layer: ~20 us for the last of 76 layers)
  if (threadIdx.x == 0) s_layer = 0;
  __syncthreads();
  {
    int mine = 0;
    for (int i = threadIdx.x; i < num_layers; i += blockDim.x)
      mine += layers[i].block_start <= static_cast<int>(blockIdx.x) ? 1 : 0;
    if (mine) atomicAdd(&s_layer, mine);
  }
  __syncthreads();
  const int l = s_layer - 1;
  const fluxb200_gemv_layer L = layers[l];
  const int pitch = 2 * K + kModAPad;
  const int n0 = (static_cast<int>(blockIdx.x) - L.block_start) * kModColsPerBlock + warp * 16;
  const int g = lane >> 2, t = lane & 3;
  // weight rows beyond N re-read row N-1 (valid memory); their results are never stored
  const uint8_t* wbase = static_cast<const uint8_t*>(L.w) + t * 16;
  const uint8_t* wr0 = wbase + static_cast<int64_t>(min(n0 + g, L.N - 1)) * K;
  const uint8_t* wr1 = wbase + static_cast<int64_t>(min(n0 + g + 8, L.N - 1)) * K;
  const float s = __ldg(L.a_scale_recip) * __ldg(L.w_scale_recip);
  const __nv_bfloat16* bias = static_cast<const __nv_bfloat16*>(L.bias);
  const int chunks = K / 64;

  for (int m0 = 0; m0 < B; m0 += 8) {
    const int mt = min(8, B - m0);
    __syncthreads();
    {  // rows m0..m0+mt-1 of this layer's quantised activations -> f16; row `mt` of the buffer is all zeros
      const uint8_t* src = aq + (static_cast<int64_t>(l) * B + m0) * K;
      const int per_row = K / 16;  // 16 fp8 values in, 32 bytes of f16 out per item; rows are contiguous in aq
      for (int i = threadIdx.x; i < mt * per_row; i += blockDim.x) {
        const int r = i / per_row, c = i - r * per_row;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
        uint4 lo, hi;
        lo.x = fp8x2_to_f16x2<AFMT>(v.x & 0xffffu), lo.y = fp8x2_to_f16x2<AFMT>(v.x >> 16);
        lo.z = fp8x2_to_f16x2<AFMT>(v.y & 0xffffu), lo.w = fp8x2_to_f16x2<AFMT>(v.y >> 16);
        hi.x = fp8x2_to_f16x2<AFMT>(v.z & 0xffffu), hi.y = fp8x2_to_f16x2<AFMT>(v.z >> 16);
        hi.z = fp8x2_to_f16x2<AFMT>(v.w & 0xffffu), hi.w = fp8x2_to_f16x2<AFMT>(v.w >> 16);
        uint4* dst = reinterpret_cast<uint4*>(a_sm8 + r * pitch + c * 32);
        dst[0] = lo, dst[1] = hi;
      }
      for (int i = threadIdx.x; i < pitch / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_sm8 + mt * pitch)[i] = 0u;
    }
    __syncthreads();
    if (n0 < L.N) {
      const uint8_t* ab = a_sm8 + (g < mt ? g : mt) * pitch + t * 32;
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      auto consume = [&](const uint4& w0, const uint4& w1, int kc) {
        const uint4 b01 = *reinterpret_cast<const uint4*>(ab + kc * 128);
        const uint4 b23 = *reinterpret_cast<const uint4*>(ab + kc * 128 + 16);
        mma_m16n8k16_f16(c0, fp8x2_to_f16x2<0>(w0.x & 0xffffu), fp8x2_to_f16x2<0>(w1.x & 0xffffu),
                         fp8x2_to_f16x2<0>(w0.x >> 16), fp8x2_to_f16x2<0>(w1.x >> 16), b01.x, b01.y);
        mma_m16n8k16_f16(c1, fp8x2_to_f16x2<0>(w0.y & 0xffffu), fp8x2_to_f16x2<0>(w1.y & 0xffffu),
                         fp8x2_to_f16x2<0>(w0.y >> 16), fp8x2_to_f16x2<0>(w1.y >> 16), b01.z, b01.w);
        mma_m16n8k16_f16(c0, fp8x2_to_f16x2<0>(w0.z & 0xffffu), fp8x2_to_f16x2<0>(w1.z & 0xffffu),
                         fp8x2_to_f16x2<0>(w0.z >> 16), fp8x2_to_f16x2<0>(w1.z >> 16), b23.x, b23.y);
        mma_m16n8k16_f16(c1, fp8x2_to_f16x2<0>(w0.w & 0xffffu), fp8x2_to_f16x2<0>(w1.w & 0xffffu),
                         fp8x2_to_f16x2<0>(w0.w >> 16), fp8x2_to_f16x2<0>(w1.w >> 16), b23.z, b23.w);
      };
      // software pipeline over groups of kModU chunks, two register buffers: the loads of group i+1 are in flight
      // while group i is consumed (8..16 outstanding 16-byte loads per lane; the kernel is latency/MLP bound)
      uint4 bufA[kModU][2], bufB[kModU][2];
      auto fetch = [&](uint4 (&buf)[kModU][2], int kc0) {
#pragma unroll
        for (int u = 0; u < kModU; ++u) {
          buf[u][0] = __ldg(reinterpret_cast<const uint4*>(wr0 + (kc0 + u) * 64));
          buf[u][1] = __ldg(reinterpret_cast<const uint4*>(wr1 + (kc0 + u) * 64));
        }
      };
      const int full = chunks / (2 * kModU) * (2 * kModU);
      if (full > 0) fetch(bufA, 0);
      for (int kc0 = 0; kc0 < full; kc0 += 2 * kModU) {
        fetch(bufB, kc0 + kModU);
#pragma unroll
        for (int u = 0; u < kModU; ++u) consume(bufA[u][0], bufA[u][1], kc0 + u);
        if (kc0 + 2 * kModU < full) fetch(bufA, kc0 + 2 * kModU);
#pragma unroll
        for (int u = 0; u < kModU; ++u) consume(bufB[u][0], bufB[u][1], kc0 + kModU + u);
      }
      for (int kc = full; kc < chunks; ++kc) {
        const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(wr0 + kc * 64));
        const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(wr1 + kc * 64));
        consume(w0, w1, kc);
      }
      // D fragment: c[0],c[1] -> weight row n0+g, batch rows 2t, 2t+1;  c[2],c[3] -> weight row n0+g+8
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + g + (j >> 1) * 8, row = t * 2 + (j & 1);
        if (row < mt && col < L.N) {
          const float bb = bias ? __bfloat162float(bias[col]) : 0.f;
          out[static_cast<int64_t>(m0 + row) * ld_out + L.out_offset + col] = __float2bfloat16_rn(fmaf(c0[j] + c1[j], s, bb));
        }
      }
    }
  }
}

// bf16 form of the batched modulation GEMV: Modulation.lin left as nn.Linear (quantize_modulation = False, BASELINE
// config c5; reference modules/flux_model.py:233-257 with float8_quantize.py:346 skipping the swap) =
//   out[b, n] = bf16( sum_k bf16(silu(vec[b,k])) * W[n,k] + bias[n] )       (fp32 accumulate, as cuBLAS does)
// Same structure as gemv_layers_mma_kernel: the bf16 WEIGHTS are the 16-row A operand of mma.sync.m16n8k16 (bf16
// operands, fp32 accumulate), the batch (<= 8 rows per pass) the B operand.  K is consumed in 64-byte (32-element)
// chunks permuted so that a lane's load is one contiguous 16 bytes: quad lane t owns elements [8t, 8t+8) of the chunk;
// words (0,1) of it are the (k 2t.., k 2t+8..) slots of the chunk's first MMA, words (2,3) of the second, and the
// activation fragment is read from the same elements.  silu is applied while the block stages vec into shared
// memory (every layer consumes the same bf16(silu(vec)): there is no per-layer quantisation), so the whole step's
// modulation is ONE launch streaming 6.46 GB of bf16 weights.
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                  uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// row pitch 2K + 64 bytes: a quarter-warp (two batch rows x four 16-byte pieces) covers all 32 banks exactly once
constexpr int kModBf16Pad = 64;

// SILU: apply bf16(silu(.)) to the input while staging it (Modulation / MLPEmbedder.out_layer / LastLayer.adaLN).
// ONE: a single layer passed BY VALUE (`one`, no device table) with up to two bf16 [B, N] addends applied after the
// bias, each followed by its own bf16 rounding (Flux.forward: vec = time_in(..) + guidance_in(..) + vector_in(y)).
struct GemvOne {
  fluxb200_gemv_layer layer;
  const __nv_bfloat16* add0;
  const __nv_bfloat16* add1;
  int64_t ld_add;
};

template <bool SILU, bool ONE>
__global__ void __launch_bounds__(kModMmaWarps * 32, 4) gemv_layers_bf16_kernel(
    const fluxb200_gemv_layer* __restrict__ layers, int num_layers, const __nv_bfloat16* __restrict__ vec,
    __nv_bfloat16* __restrict__ out, int64_t ld_out, int B, int K, const GemvOne one) {
  pdl_wait();
  extern __shared__ __align__(16) uint8_t a_sm8[];
  __shared__ int s_layer;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int l = 0;
  if constexpr (!ONE) {
    if (threadIdx.x == 0) s_layer = 0;
    __syncthreads();
    {
      int mine = 0;
      for (int i = threadIdx.x; i < num_layers; i += blockDim.x)
        mine += layers[i].block_start <= static_cast<int>(blockIdx.x) ? 1 : 0;
      if (mine) atomicAdd(&s_layer, mine);
    }
    __syncthreads();
    l = s_layer - 1;
  }
  const fluxb200_gemv_layer L = ONE ? one.layer : layers[l];
  const int pitch = 2 * K + kModBf16Pad;
  const int n0 = (static_cast<int>(blockIdx.x) - L.block_start) * kModColsPerBlock + warp * 16;
  const int g = lane >> 2, t = lane & 3;
  const uint8_t* wbase = static_cast<const uint8_t*>(L.w) + t * 16;
  const uint8_t* wr0 = wbase + static_cast<int64_t>(min(n0 + g, L.N - 1)) * K * 2;
  const uint8_t* wr1 = wbase + static_cast<int64_t>(min(n0 + g + 8, L.N - 1)) * K * 2;
  const __nv_bfloat16* bias = static_cast<const __nv_bfloat16*>(L.bias);
  const int chunks = K / 32;

  for (int m0 = 0; m0 < B; m0 += 8) {
    const int mt = min(8, B - m0);
    __syncthreads();
    {  // bf16(silu(vec)) rows m0..m0+mt-1; row `mt` of the buffer is all zeros (batch padding of the B operand)
      for (int i = threadIdx.x; i < mt * K; i += blockDim.x) {
        const int r = i / K, c = i - r * K;
        const float v = __bfloat162float(vec[static_cast<int64_t>(m0 + r) * K + c]);
        *reinterpret_cast<__nv_bfloat16*>(a_sm8 + r * pitch + c * 2) = __float2bfloat16_rn(SILU ? v / (1.f + expf(-v)) : v);
      }
      for (int i = threadIdx.x; i < pitch / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_sm8 + mt * pitch)[i] = 0u;
    }
    __syncthreads();
    if (n0 < L.N) {
      const uint8_t* ab = a_sm8 + (g < mt ? g : mt) * pitch + t * 16;
      float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
      auto consume = [&](const uint4& w0, const uint4& w1, int kc) {
        const uint4 b = *reinterpret_cast<const uint4*>(ab + kc * 64);
        mma_m16n8k16_bf16(c0, w0.x, w1.x, w0.y, w1.y, b.x, b.y);
        mma_m16n8k16_bf16(c1, w0.z, w1.z, w0.w, w1.w, b.z, b.w);
      };
      uint4 bufA[kModU][2], bufB[kModU][2];
      auto fetch = [&](uint4 (&buf)[kModU][2], int kc0) {
#pragma unroll
        for (int u = 0; u < kModU; ++u) {
          buf[u][0] = __ldg(reinterpret_cast<const uint4*>(wr0 + (kc0 + u) * 64));
          buf[u][1] = __ldg(reinterpret_cast<const uint4*>(wr1 + (kc0 + u) * 64));
        }
      };
      const int full = chunks / (2 * kModU) * (2 * kModU);
      if (full > 0) fetch(bufA, 0);
      for (int kc0 = 0; kc0 < full; kc0 += 2 * kModU) {
        fetch(bufB, kc0 + kModU);
#pragma unroll
        for (int u = 0; u < kModU; ++u) consume(bufA[u][0], bufA[u][1], kc0 + u);
        if (kc0 + 2 * kModU < full) fetch(bufA, kc0 + 2 * kModU);
#pragma unroll
        for (int u = 0; u < kModU; ++u) consume(bufB[u][0], bufB[u][1], kc0 + kModU + u);
      }
      for (int kc = full; kc < chunks; ++kc) {
        const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(wr0 + kc * 64));
        const uint4 w1 = __ldg(reinterpret_cast<const uint4*>(wr1 + kc * 64));
        consume(w0, w1, kc);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + g + (j >> 1) * 8, row = t * 2 + (j & 1);
        if (row < mt && col < L.N) {
          const float bb = bias ? __bfloat162float(bias[col]) : 0.f;
          float y = bf16r(c0[j] + c1[j] + bb);
          if constexpr (ONE) {
            if (one.add0) y = bf16r(y + __bfloat162float(one.add0[static_cast<int64_t>(m0 + row) * one.ld_add + col]));
            if (one.add1) y = bf16r(y + __bfloat162float(one.add1[static_cast<int64_t>(m0 + row) * one.ld_add + col]));
          }
          out[static_cast<int64_t>(m0 + row) * ld_out + L.out_offset + col] = __float2bfloat16_rn(y);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// timestep_embedding(t, dim) .type(bf16)          modules/flux_model.py:95-116 (called at :687, :694)
//   t' = bf16(time_factor * t) (the product is formed in t's dtype, bf16);  args = float(t') * freqs[j]  (fp32);
//   emb = [cos(args) | sin(args)] -> bf16.   freqs = exp(-ln(max_period) * arange(half) / half) is passed in (fp32).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) timestep_embedding_kernel(const __nv_bfloat16* __restrict__ t,
                                                                 const float* __restrict__ freqs,
                                                                 __nv_bfloat16* __restrict__ out, int B, int half,
                                                                 float time_factor) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, j = i - b * half;
  const float ts = bf16r(time_factor * __bfloat162float(t[b]));
  const float a = ts * freqs[j];
  out[static_cast<int64_t>(b) * 2 * half + j] = __float2bfloat16_rn(cosf(a));
  out[static_cast<int64_t>(b) * 2 * half + half + j] = __float2bfloat16_rn(sinf(a));
}

// ---------------------------------------------------------------------------------------------
// Euler step of the flow: img + (t_prev - t_curr) * pred          flux_pipeline.py:651
//   eager torch: the python scalar multiplies in fp32, the product is rounded to bf16, the sum is rounded to bf16
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) euler_update_kernel(const __nv_bfloat16* __restrict__ img,
                                                           const __nv_bfloat16* __restrict__ pred,
                                                           const float* __restrict__ dt, __nv_bfloat16* __restrict__ out,
                                                           int64_t n) {
  pdl_wait();
  const float d = __ldg(dt);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __float2bfloat16_rn(__bfloat162float(img[i]) + bf16r(d * __bfloat162float(pred[i])));
}

// ---------------------------------------------------------------------------------------------
// Small bf16 GEMM for the two un-quantised linears that bracket the block stack:
//   img_in      [B*L, 64]   x [3072, 64]^T   (modules/flux_model.py:686)          K = 64, output-write bound
//   final_layer [B*L, 3072] x [64, 3072]^T   (:502, 716) + the Euler update (flux_pipeline.py:651) fused
//   out[m, n] = bf16( sum_k x[m,k] W[n,k] + bias[n] )            (fp32 accumulate, as cuBLAS)
//   euler:  out[m, n] = bf16( img[m, n] + bf16( dt * out[m, n] ) )
// 0.03 % of the step's flops: mma.sync.m16n8k16 (bf16) from shared memory, 32 x 64 tiles, 4-stage cp.async ring; the
// point is that the step launches no library kernel, not tensor-core peak.
// ---------------------------------------------------------------------------------------------
constexpr int kSgBM = 32, kSgBN = 64, kSgBK = 32, kSgStages = 4, kSgPitch = kSgBK + 8;  // pitch 40 bf16 = 80 B

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(128) bf16_gemm_small_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                              const __nv_bfloat16* __restrict__ w,
                                                              const __nv_bfloat16* __restrict__ bias,
                                                              __nv_bfloat16* __restrict__ out, int64_t ldo,
                                                              const __nv_bfloat16* __restrict__ euler_img,
                                                              const float* __restrict__ euler_dt, int M, int N, int K) {
  pdl_wait();
  __shared__ __align__(16) __nv_bfloat16 As[kSgStages][kSgBM][kSgPitch];
  __shared__ __align__(16) __nv_bfloat16 Ws[kSgStages][kSgBN][kSgPitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.x * kSgBM, n0 = blockIdx.y * kSgBN;
  const int wm = (warp & 1) * 16, wn = (warp >> 1) * 32;  // warp tile: 16 rows x 32 columns
  const int kt = K / kSgBK;

  auto load_stage = [&](int stage, int kb) {
    // A: 32 rows x 64 B = 128 16-byte pieces; W: 64 rows x 64 B = 256 pieces; 128 threads -> 1 + 2 pieces each
    {
      const int r = threadIdx.x >> 2, c = threadIdx.x & 3;
      const bool ok = m0 + r < M;
      cp_async16(&As[stage][r][c * 8], x + static_cast<int64_t>(ok ? m0 + r : 0) * ldx + kb * kSgBK + c * 8, ok);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int p = threadIdx.x + i * 128, r = p >> 2, c = p & 3;
      const bool ok = n0 + r < N;
      cp_async16(&Ws[stage][r][c * 8], w + static_cast<int64_t>(ok ? n0 + r : 0) * K + kb * kSgBK + c * 8, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

  for (int s = 0; s < kSgStages - 1; ++s) {
    if (s < kt) load_stage(s, s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  for (int kb = 0; kb < kt; ++kb) {
    asm volatile("cp.async.wait_group %0;" ::"n"(kSgStages - 2) : "memory");
    __syncthreads();
    {  // prefetch k-block kb + stages - 1 into the slot freed last iteration
      const int nk = kb + kSgStages - 1;
      if (nk < kt) load_stage(nk % kSgStages, nk);
      else asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const int st = kb % kSgStages;
#pragma unroll
    for (int kk = 0; kk < kSgBK; kk += 16) {
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(&As[st][wm + g][kk + 2 * t]);
      const uint32_t a1 = *reinterpret_cast<const uint32_t*>(&As[st][wm + g + 8][kk + 2 * t]);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(&As[st][wm + g][kk + 2 * t + 8]);
      const uint32_t a3 = *reinterpret_cast<const uint32_t*>(&As[st][wm + g + 8][kk + 2 * t + 8]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(&Ws[st][wn + j * 8 + g][kk + 2 * t]);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(&Ws[st][wn + j * 8 + g][kk + 2 * t + 8]);
        mma_m16n8k16_bf16(acc[j], a0, a1, a2, a3, b0, b1);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const float dt = euler_dt ? __ldg(euler_dt) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = m0 + wm + g + h * 8, col = n0 + wn + j * 8 + 2 * t;
      if (row >= M || col >= N) continue;  // N is even (validated): col + 1 < N as well
      float y0 = bf16r(acc[j][h * 2] + (bias ? __bfloat162float(bias[col]) : 0.f));
      float y1 = bf16r(acc[j][h * 2 + 1] + (bias ? __bfloat162float(bias[col + 1]) : 0.f));
      if (euler_img) {
        const __nv_bfloat162 im = *reinterpret_cast<const __nv_bfloat162*>(euler_img + static_cast<int64_t>(row) * ldo + col);
        y0 = __bfloat162float(im.x) + bf16r(dt * y0);
        y1 = __bfloat162float(im.y) + bf16r(dt * y1);
      }
      *reinterpret_cast<uint32_t*>(out + static_cast<int64_t>(row) * ldo + col) = pack_bf16x2(y0, y1);
    }
  }
}

static int grid_for(int64_t work_items, int threads, int max_blocks_per_sm = 8) {
  int64_t blocks = (work_items + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(sm_count()) * max_blocks_per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace fb

using namespace fb;

extern "C" int fluxb200_quantize(const void* x, void* y, int64_t n, const float* scale, int fmt,
                                 fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && y && scale && n >= 0, "fluxb200_quantize: null operand");
  FB_REQUIRE(fmt == 0 || fmt == 1, "fluxb200_quantize: bad fp8 format %d", fmt);
  FB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0,
             "fluxb200_quantize: x must be 16B and y 8B aligned");
  if (n == 0) return 0;
  const int grid = grid_for(n / 8 + 1, 256);
  FB_CUDA_OK(launch_kernel(fmt == 0 ? quantize_kernel<0> : quantize_kernel<1>, dim3(grid), dim3(256), 0, stream, 1,
                           static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(y), n, scale));
  return 0;
}

extern "C" int fluxb200_amax(const void* x, int64_t n, float* amax, fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && amax && n >= 0, "fluxb200_amax: null operand");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "fluxb200_amax: x must be 16B aligned");
  if (n == 0) return 0;
  FB_CUDA_OK(launch_kernel(amax_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, stream, 1,
                           static_cast<const __nv_bfloat16*>(x), n, amax));
  return 0;
}

extern "C" int fluxb200_silu_quant(const void* x, void* y_fp8, void* y_bf16, int64_t n, const float* scale, int fmt,
                                   fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && (y_fp8 || y_bf16), "fluxb200_silu_quant: null operand");
  FB_REQUIRE(!y_fp8 || scale, "fluxb200_silu_quant: scale required for fp8 output");
  FB_REQUIRE(fmt == 0 || fmt == 1, "fluxb200_silu_quant: bad fp8 format %d", fmt);
  if (n == 0) return 0;
  const int grid = grid_for(n, 256);
  FB_CUDA_OK(launch_kernel(fmt == 0 ? silu_quant_kernel<0> : silu_quant_kernel<1>, dim3(grid), dim3(256), 0, stream, 1,
                           static_cast<const __nv_bfloat16*>(x), static_cast<uint8_t*>(y_fp8),
                           static_cast<__nv_bfloat16*>(y_bf16), n, scale));
  return 0;
}

extern "C" int fluxb200_ln_mod_quant(const void* x, int64_t ldx, const void* shift, const void* scale,
                                     int64_t mod_batch_stride, void* y_fp8, int64_t ldy, void* y_bf16,
                                     int64_t ldy_bf16, const float* in_scale, int fmt, int B, int L, int D, float eps,
                                     fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && shift && scale && (y_fp8 || y_bf16), "fluxb200_ln_mod_quant: null operand");
  FB_REQUIRE(!y_fp8 || in_scale, "fluxb200_ln_mod_quant: in_scale required for fp8 output");
  FB_REQUIRE(B > 0 && L > 0 && D > 0 && D % 256 == 0 && D <= 256 * kLnMaxIter,
             "fluxb200_ln_mod_quant: D=%d must be a multiple of 256 and <= %d", D, 256 * kLnMaxIter);
  FB_REQUIRE(ldx % 8 == 0 && mod_batch_stride % 8 == 0 && (!y_fp8 || ldy % 8 == 0) && (!y_bf16 || ldy_bf16 % 8 == 0),
             "fluxb200_ln_mod_quant: strides must keep 16-byte (bf16) / 8-byte (fp8) alignment");
  FB_REQUIRE(fmt == 0 || fmt == 1, "fluxb200_ln_mod_quant: bad fp8 format %d", fmt);
  fb::LnParams P{};
  P.seg[0] = fb::LnSeg{static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(shift),
                       static_cast<const __nv_bfloat16*>(scale), static_cast<uint8_t*>(y_fp8),
                       static_cast<__nv_bfloat16*>(y_bf16), in_scale, ldx, ldy, ldy_bf16, mod_batch_stride, B * L, L};
  P.seg[1] = P.seg[0];
  P.seg[1].rows = 0;
  P.D = D;
  P.eps = eps;
  const int grid = (B * L + 7) / 8;
  // the specialised form is the steady-state one: fp8 output only
  auto kern = (D == 3072 && y_fp8 && !y_bf16) ? (fmt == 0 ? ln_mod_quant_kernel<0, 12> : ln_mod_quant_kernel<1, 12>)
                                              : (fmt == 0 ? ln_mod_quant_kernel<0, 0> : ln_mod_quant_kernel<1, 0>);
  FB_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(256), 0, stream, 1, P));
  return 0;
}

extern "C" int fluxb200_ln_mod_quant_grouped(const fluxb200_ln_args* args, int count, int fmt, int D, float eps,
                                             fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(args && (count == 1 || count == 2), "fluxb200_ln_mod_quant_grouped: 1 or 2 row sets");
  FB_REQUIRE(fmt == 0 || fmt == 1, "fluxb200_ln_mod_quant_grouped: bad fp8 format %d", fmt);
  FB_REQUIRE(D > 0 && D % 256 == 0 && D <= kLnMaxIter * 256, "fluxb200_ln_mod_quant_grouped: D=%d must be a multiple of 256, <= %d",
             D, kLnMaxIter * 256);
  fb::LnParams P{};
  int total = 0;
  for (int i = 0; i < count; ++i) {
    const fluxb200_ln_args& a = args[i];
    FB_REQUIRE(a.x && a.shift && a.scale && a.y_fp8 && a.B > 0 && a.L > 0, "fluxb200_ln_mod_quant_grouped: bad row set %d", i);
    FB_REQUIRE(a.ldx % 8 == 0 && a.mod_batch_stride % 8 == 0 && a.ldy % 8 == 0,
               "fluxb200_ln_mod_quant_grouped: strides must keep 16-byte (bf16) / 8-byte (fp8) alignment");
    P.seg[i] = fb::LnSeg{static_cast<const __nv_bfloat16*>(a.x), static_cast<const __nv_bfloat16*>(a.shift),
                         static_cast<const __nv_bfloat16*>(a.scale), static_cast<uint8_t*>(a.y_fp8), nullptr, a.in_scale,
                         a.ldx, a.ldy, 0, a.mod_batch_stride, a.B * a.L, a.L};
    total += a.B * a.L;
  }
  if (count == 1) {
    P.seg[1] = P.seg[0];
    P.seg[1].rows = 0;
  }
  P.D = D;
  P.eps = eps;
  // rows of the second set start on a block boundary only if the first count is a multiple of 8: pad the grid instead
  const int grid = (P.seg[0].rows + 7) / 8 + (P.seg[1].rows + 7) / 8;
  auto kern = D == 3072 ? (fmt == 0 ? ln_mod_quant_kernel<0, 12> : ln_mod_quant_kernel<1, 12>)
                        : (fmt == 0 ? ln_mod_quant_kernel<0, 0> : ln_mod_quant_kernel<1, 0>);
  (void)total;
  FB_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(256), 0, stream, 1, P));
  return 0;
}

extern "C" int fluxb200_qknorm_rope(const void* x, void* y, const float* norm_w, const void* rope_cos,
                                    const void* rope_sin, int64_t rope_batch_stride, int B, int H, int S, float eps,
                                    fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && y && B > 0 && H > 0 && S > 0, "fluxb200_qknorm_rope: bad operand");
  FB_REQUIRE((rope_cos == nullptr) == (rope_sin == nullptr), "fluxb200_qknorm_rope: cos and sin go together");
  const int64_t rows = static_cast<int64_t>(B) * H * S;
  const int64_t grid = (rows + 7) / 8;
  FB_CUDA_OK(launch_kernel(qknorm_rope_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), 0, stream, 1,
                           static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), norm_w,
                           static_cast<const __nv_bfloat16*>(rope_cos), static_cast<const __nv_bfloat16*>(rope_sin),
                           rope_batch_stride, rows, H, S, eps));
  return 0;
}

extern "C" int fluxb200_f8_gemv(const void* a, int a_fmt, const void* w, int w_fmt, const void* bias,
                                const float* sa, const float* sw, void* out, int M, int N, int K,
                                fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(a && w && sa && sw && out, "fluxb200_f8_gemv: null operand");
  FB_REQUIRE(M > 0 && M <= 16 && N > 0 && K > 0 && K % 16 == 0, "fluxb200_f8_gemv: need 0<M<=16, K %% 16 == 0");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(a) & 1) == 0,
             "fluxb200_f8_gemv: w must be 16B aligned");
  FB_REQUIRE((a_fmt == 0 || a_fmt == 1) && (w_fmt == 0 || w_fmt == 1), "fluxb200_f8_gemv: bad fp8 format");
  const size_t smem = static_cast<size_t>(kGemvMT) * K * sizeof(__half);
  FB_REQUIRE(smem <= 200 * 1024, "fluxb200_f8_gemv: K=%d too large", K);
  const int cols_per_block = kGemvWarps * kGemvColsPerWarp;
  const int grid = (N + cols_per_block - 1) / cols_per_block;
#define FB_GEMV(AF, WF)                                                                                         \
  do {                                                                                                          \
    auto kern = f8_gemv_kernel<AF, WF>;                                                                         \
    if (smem > 48 * 1024) FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    FB_CUDA_OK(launch_kernel(kern, dim3(grid), dim3(kGemvWarps * 32), smem, stream, 1, static_cast<const uint8_t*>(a), \
                             static_cast<const uint8_t*>(w), static_cast<const __nv_bfloat16*>(bias), sa, sw,        \
                             static_cast<__nv_bfloat16*>(out), M, N, K));                                            \
  } while (0)
  if (a_fmt == 0 && w_fmt == 0) FB_GEMV(0, 0);
  else if (a_fmt == 1 && w_fmt == 0) FB_GEMV(1, 0);
  else if (a_fmt == 0 && w_fmt == 1) FB_GEMV(0, 1);
  else FB_GEMV(1, 1);
#undef FB_GEMV
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int fluxb200_modulation_batched(const void* vec, const fluxb200_gemv_layer* layers, int num_layers,
                                           int total_blocks, void* aq, void* out, int64_t ld_out, int B, int K,
                                           int a_fmt, int w_fmt, fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(vec && layers && aq && out, "fluxb200_modulation_batched: null operand");
  FB_REQUIRE(num_layers > 0 && total_blocks > 0 && B > 0 && B <= 16, "fluxb200_modulation_batched: bad sizes");
  FB_REQUIRE(K % 16 == 0 && K <= 4096, "fluxb200_modulation_batched: K=%d must be a multiple of 16 and <= 4096", K);
  FB_REQUIRE((a_fmt == 0 || a_fmt == 1) && w_fmt == 0, "fluxb200_modulation_batched: e4m3 weights, e4m3/e5m2 inputs");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(aq) & 15) == 0, "fluxb200_modulation_batched: workspace must be 16B aligned");
  const __nv_bfloat16* v = static_cast<const __nv_bfloat16*>(vec);
  uint8_t* a8 = static_cast<uint8_t*>(aq);
  __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
#define FB_MOD(AF, NK_)                                                                                   \
  do {                                                                                                    \
    FB_CUDA_OK(launch_kernel(silu_quant_layers_kernel<AF>, dim3(num_layers), dim3(256), 0, stream, 1, v, layers, a8, B * K)); \
    FB_CUDA_OK(launch_kernel(gemv_layers_kernel<AF, 0, NK_>, dim3(total_blocks), dim3(256), 0, stream, 1, layers, num_layers, \
                             static_cast<const uint8_t*>(a8), o, ld_out, B, K));                                          \
  } while (0)
  static const bool use_mma = [] {
    const char* e = getenv("FLUXB200_GEMV_MMA");
    return e == nullptr || atoi(e) != 0;
  }();
  if (use_mma && K % 64 == 0) {
    const size_t smem = static_cast<size_t>((B < 8 ? B : 8) + 1) * (2 * K + kModAPad);
    auto kern = a_fmt == 0 ? gemv_layers_mma_kernel<0> : gemv_layers_mma_kernel<1>;
    if (smem > 48 * 1024) FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_CUDA_OK(launch_kernel(a_fmt == 0 ? silu_quant_layers_kernel<0> : silu_quant_layers_kernel<1>, dim3(num_layers),
                             dim3(256), 0, stream, 1, v, layers, a8, B * K));
    FB_CUDA_OK(launch_kernel(kern, dim3(total_blocks), dim3(kModMmaWarps * 32), smem, stream, 1, layers, num_layers,
                             static_cast<const uint8_t*>(a8), o, ld_out, B, K));
    FB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const int nk = (K + 511) / 512;
  if (a_fmt == 0) {
    if (nk <= 1) FB_MOD(0, 1); else if (nk <= 2) FB_MOD(0, 2); else if (nk <= 6) FB_MOD(0, 6); else FB_MOD(0, 8);
  } else {
    if (nk <= 1) FB_MOD(1, 1); else if (nk <= 2) FB_MOD(1, 2); else if (nk <= 6) FB_MOD(1, 6); else FB_MOD(1, 8);
  }
#undef FB_MOD
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int fluxb200_modulation_batched_bf16(const void* vec, const fluxb200_gemv_layer* layers, int num_layers,
                                                int total_blocks, void* out, int64_t ld_out, int B, int K,
                                                fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(vec && layers && out, "fluxb200_modulation_batched_bf16: null operand");
  FB_REQUIRE(num_layers > 0 && total_blocks > 0 && B > 0 && B <= 16, "fluxb200_modulation_batched_bf16: bad sizes");
  FB_REQUIRE(K % 32 == 0 && K <= 4096, "fluxb200_modulation_batched_bf16: K=%d must be a multiple of 32 and <= 4096", K);
  const size_t smem = static_cast<size_t>((B < 8 ? B : 8) + 1) * (2 * K + kModBf16Pad);
  auto kern = gemv_layers_bf16_kernel<true, false>;
  if (smem > 48 * 1024) FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FB_CUDA_OK(launch_kernel(kern, dim3(total_blocks), dim3(kModMmaWarps * 32), smem, stream, 1, layers, num_layers,
                           static_cast<const __nv_bfloat16*>(vec), static_cast<__nv_bfloat16*>(out), ld_out, B, K,
                           GemvOne{}));
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int fluxb200_bf16_gemv(const void* x, const void* w, const void* bias, const void* add0, const void* add1,
                                  int64_t ld_add, void* out, int64_t ld_out, int B, int N, int K, int silu_input,
                                  fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && w && out, "fluxb200_bf16_gemv: null operand");
  FB_REQUIRE(B > 0 && B <= 16 && N > 0, "fluxb200_bf16_gemv: need 0 < B <= 16, N > 0");
  FB_REQUIRE(K % 32 == 0 && K > 0 && K <= 4096, "fluxb200_bf16_gemv: K=%d must be a multiple of 32 and <= 4096", K);
  FB_REQUIRE((reinterpret_cast<uintptr_t>(w) & 15) == 0, "fluxb200_bf16_gemv: w must be 16-byte aligned");
  FB_REQUIRE(ld_out >= N && (add0 == nullptr || ld_add >= N) && (add1 == nullptr || ld_add >= N),
             "fluxb200_bf16_gemv: row strides smaller than N");
  GemvOne one{};
  one.layer.w = w;
  one.layer.bias = bias;
  one.layer.N = N;
  one.layer.out_offset = 0;
  one.layer.block_start = 0;
  one.add0 = static_cast<const __nv_bfloat16*>(add0);
  one.add1 = static_cast<const __nv_bfloat16*>(add1);
  one.ld_add = ld_add;
  const size_t smem = static_cast<size_t>((B < 8 ? B : 8) + 1) * (2 * K + kModBf16Pad);
  const int blocks = (N + kModColsPerBlock - 1) / kModColsPerBlock;
  auto kern = silu_input ? gemv_layers_bf16_kernel<true, true> : gemv_layers_bf16_kernel<false, true>;
  if (smem > 48 * 1024) FB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  FB_CUDA_OK(launch_kernel(kern, dim3(blocks), dim3(kModMmaWarps * 32), smem, stream, 1,
                           static_cast<const fluxb200_gemv_layer*>(nullptr), 1, static_cast<const __nv_bfloat16*>(x),
                           static_cast<__nv_bfloat16*>(out), ld_out, B, K, one));
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int fluxb200_timestep_embedding(const void* t, const float* freqs, void* out, int B, int dim, float time_factor,
                                           fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(t && freqs && out && B > 0 && dim > 0 && dim % 2 == 0, "fluxb200_timestep_embedding: bad operand (dim even)");
  const int half = dim / 2;
  FB_CUDA_OK(launch_kernel(timestep_embedding_kernel, dim3((B * half + 127) / 128), dim3(128), 0, stream, 1,
                           static_cast<const __nv_bfloat16*>(t), freqs, static_cast<__nv_bfloat16*>(out), B, half,
                           time_factor));
  return 0;
}

extern "C" int fluxb200_euler_update(const void* img, const void* pred, const float* dt, void* out, int64_t n,
                                     fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(img && pred && dt && out && n >= 0, "fluxb200_euler_update: null operand");
  if (n == 0) return 0;
  FB_CUDA_OK(launch_kernel(euler_update_kernel, dim3(grid_for(n, 256)), dim3(256), 0, stream, 1,
                           static_cast<const __nv_bfloat16*>(img), static_cast<const __nv_bfloat16*>(pred), dt,
                           static_cast<__nv_bfloat16*>(out), n));
  return 0;
}

extern "C" int fluxb200_bf16_gemm_small(const void* x, int64_t ldx, const void* w, const void* bias, void* out, int64_t ldo,
                                        const void* euler_img, const float* euler_dt, int M, int N, int K,
                                        fluxb200_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(x && w && out && M > 0 && N > 0 && K > 0, "fluxb200_bf16_gemm_small: null operand / empty shape");
  FB_REQUIRE(K % 32 == 0 && N % 2 == 0, "fluxb200_bf16_gemm_small: K=%d must be a multiple of 32 and N=%d even", K, N);
  FB_REQUIRE(ldx % 8 == 0 && ldo % 2 == 0 && ldx >= K && ldo >= N, "fluxb200_bf16_gemm_small: ldx %% 8, ldo %% 2, ld >= width");
  FB_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(out) & 3) == 0,
             "fluxb200_bf16_gemm_small: x / w must be 16-byte, out 4-byte aligned");
  FB_REQUIRE((euler_img == nullptr) == (euler_dt == nullptr), "fluxb200_bf16_gemm_small: euler_img and euler_dt go together");
  const dim3 grid((M + kSgBM - 1) / kSgBM, (N + kSgBN - 1) / kSgBN);
  FB_REQUIRE(grid.y <= 65535, "fluxb200_bf16_gemm_small: N too large");
  FB_CUDA_OK(launch_kernel(bf16_gemm_small_kernel, grid, dim3(128), 0, stream, 1, static_cast<const __nv_bfloat16*>(x), ldx,
                           static_cast<const __nv_bfloat16*>(w), static_cast<const __nv_bfloat16*>(bias),
                           static_cast<__nv_bfloat16*>(out), ldo, static_cast<const __nv_bfloat16*>(euler_img), euler_dt, M,
                           N, K));
  return 0;
}
