// On-device LoRA fuse / unfuse for one F8Linear.
//
// Replaces, for a quantised layer, the reference's chain of full-size fp32 temporaries
//   extract_weight_from_linear  (lora_loading.py:615-626)  W   = float8_data.float() * scale_reciprocal
//   calculate_lora_weight       (lora_loading.py:509-547)  D   = lora_scale * (lora_B @ lora_A)      [fp32 mm]
//   apply_lora_weight_to_module (lora_loading.py:566-577)  W'  = (W + D).to(weight dtype)            [unfuse: W - D, :549-563]
//   F8Linear.set_weight_tensor -> quantize_weight (float8_quantize.py:209-212, 195-207)  amax(W') -> scale -> fp8
// by one kernel that dequantises, applies the rank-R update from shared-memory tiles of the two factors, rounds to
// bf16 once, writes W' and reduces max|W'| -- 1 byte read + 2 bytes written per weight -- followed by the existing
// quantise kernel (fluxb200_quantize) writing the new bytes over the old ones IN PLACE, so CUDA graphs and the
// modulation table captured over the layer's buffers stay valid across a LoRA hot-swap.
// HBM-bound for small ranks (R <= 32), fp32-FMA bound above; not on the per-step path.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "flux_b200.h"
#include "host_util.h"
#include "ptx.cuh"

namespace fb {

constexpr int kLoraTile = 64;  // output tile (rows of W x columns of W) per block
constexpr int kLoraRT = 16;    // rank slice staged in shared memory

template <int FMT>
__device__ __forceinline__ float fp8_to_float(uint8_t v) {
  __half_raw hr = __nv_cvt_fp8_to_halfraw(v, FMT == 0 ? __NV_E4M3 : __NV_E5M2);
  return __half2float(*reinterpret_cast<__half*>(&hr));
}

// down: [N, R] fp32 (lora_B), up: [chunks*R, K] fp32 (lora_A, already scaled by alpha/rank on the host exactly as the
// reference does).  Thread (ty, tx) of the 16x16 block owns rows ty*4..+3 and columns tx*4..+3 of the tile.
template <int FMT>
__global__ void __launch_bounds__(256) lora_fuse_kernel(const uint8_t* __restrict__ w8,
                                                        const float* __restrict__ scale_recip,
                                                        const float* __restrict__ down, const float* __restrict__ up,
                                                        int N, int K, int R, int chunks, float coeff, int unfuse,
                                                        __nv_bfloat16* __restrict__ out, unsigned int* __restrict__ amax_bits) {
  __shared__ float d_sm[kLoraTile][kLoraRT + 1];
  __shared__ float u_sm[kLoraRT][kLoraTile];
  __shared__ float red[8];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int n0 = blockIdx.y * kLoraTile, k0 = blockIdx.x * kLoraTile;
  float fused[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) fused[i][j] = 0.f;

  for (int c = 0; c < chunks; ++c) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int r0 = 0; r0 < R; r0 += kLoraRT) {
      __syncthreads();
      for (int i = threadIdx.x; i < kLoraTile * kLoraRT; i += 256) {
        const int n = i / kLoraRT, r = i % kLoraRT;
        d_sm[n][r] = (n0 + n < N && r0 + r < R) ? down[static_cast<int64_t>(n0 + n) * R + r0 + r] : 0.f;
        const int rr = i / kLoraTile, k = i % kLoraTile;
        u_sm[rr][k] = (r0 + rr < R && k0 + k < K) ? up[static_cast<int64_t>(c * R + r0 + rr) * K + k0 + k] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < kLoraRT; ++r) {
        float dv[4], uv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dv[i] = d_sm[ty * 4 + i][r];
#pragma unroll
        for (int j = 0; j < 4; ++j) uv[j] = u_sm[r][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(dv[i], uv[j], acc[i][j]);
      }
    }
    // fused_lora = fused_lora + (lora_scale * mm(...)): a separate product and sum, as in the reference
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float term = __fmul_rn(coeff, acc[i][j]);
        fused[i][j] = c == 0 ? term : __fadd_rn(fused[i][j], term);
      }
  }

  const float sr = __ldg(scale_recip);
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= K) continue;
      const float w = __fmul_rn(fp8_to_float<FMT>(w8[static_cast<int64_t>(n) * K + k]), sr);
      const float f = unfuse ? __fsub_rn(w, fused[i][j]) : __fadd_rn(w, fused[i][j]);
      const __nv_bfloat16 b = __float2bfloat16_rn(f);
      out[static_cast<int64_t>(n) * K + k] = b;
      amax = fmaxf(amax, fabsf(__bfloat162float(b)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(amax_bits, __float_as_uint(m));  // non-negative floats order like their bit patterns
  }
}

}  // namespace fb

extern "C" int fluxb200_lora_fuse(const void* w_fp8, int w_fmt, const float* w_scale_recip, const float* lora_down,
                                  const float* lora_up, int N, int K, int R, int chunks, float coeff, int unfuse,
                                  void* w_out_bf16, float* amax_out, fluxb200_stream_t stream_) {
  using namespace fb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  FB_REQUIRE(w_fp8 && w_scale_recip && lora_down && lora_up && w_out_bf16 && amax_out, "fluxb200_lora_fuse: null operand");
  FB_REQUIRE(N > 0 && K > 0 && R > 0 && chunks > 0, "fluxb200_lora_fuse: bad sizes N=%d K=%d R=%d chunks=%d", N, K, R, chunks);
  FB_REQUIRE(w_fmt == 0 || w_fmt == 1, "fluxb200_lora_fuse: bad fp8 format %d", w_fmt);
  FB_CUDA_OK(cudaMemsetAsync(amax_out, 0, sizeof(float), stream));
  dim3 grid((K + kLoraTile - 1) / kLoraTile, (N + kLoraTile - 1) / kLoraTile);
  FB_REQUIRE(grid.y <= 65535, "fluxb200_lora_fuse: N=%d too large", N);
  auto kern = w_fmt == 0 ? lora_fuse_kernel<0> : lora_fuse_kernel<1>;
  kern<<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(w_fp8), w_scale_recip, lora_down, lora_up, N, K, R, chunks,
                                 coeff, unfuse, static_cast<__nv_bfloat16*>(w_out_bf16),
                                 reinterpret_cast<unsigned int*>(amax_out));
  FB_CUDA_OK(cudaGetLastError());
  return 0;
}
