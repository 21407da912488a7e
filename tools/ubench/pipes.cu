// Micro-benchmark: per-SM-sub-partition issue rate of the instruction classes the attention softmax is made of
// (B200, sm_100a).  One block per SM, W warps per block (1 per sub-partition when W = 4), 8 independent chains per
// thread, N iterations; prints cycles per warp instruction per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/pipes tools/ubench/pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float v[CHAINS];
  unsigned long long p[CHAINS / 2];
  for (int i = 0; i < CHAINS; ++i) v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f;
  for (int i = 0; i < CHAINS / 2; ++i) p[i] = (static_cast<unsigned long long>(__float_as_uint(v[2 * i + 1])) << 32) | __float_as_uint(v[2 * i]);
  const unsigned long long c2 = (static_cast<unsigned long long>(__float_as_uint(0.999f)) << 32) | __float_as_uint(0.999f);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(v[i]) : "f"(0.999f));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(0.5f), "f"(seed));
      if (OP == 4) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(v[i])); v[i] = __uint_as_float(u); }
      if (OP == 5) { int x = __float_as_int(v[i]); asm volatile("shl.b32 %0, %0, 1;" : "+r"(x)); asm volatile("add.s32 %0, %0, 3;" : "+r"(x)); v[i] = __int_as_float(x); }
      if (OP == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(0.5f));
      if (OP == 8) { unsigned u = __float_as_uint(v[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 9) { unsigned u = __float_as_uint(v[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
      if (OP == 10) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
    }
    if (OP == 2) {
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
    }
    if (OP == 11 || OP == 12 || OP == 13) {
      // mixed stream: 8 MUFU.EX2 (OP 11, 12) and/or 16 FFMA2 (OP 11, 13) per iteration, independent chains
      if (OP != 13) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      }
      if (OP != 12) {
#pragma unroll
        for (int r2 = 0; r2 < 4; ++r2)
#pragma unroll
          for (int i = 0; i < CHAINS / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
      }
    }
    if (OP == 7) {
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < CHAINS; ++i) s += v[i];
  for (int i = 0; i < CHAINS / 2; ++i) s += __uint_as_float(static_cast<unsigned>(p[i])) + __uint_as_float(static_cast<unsigned>(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int warps, int instr_per_iter) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  k<OP><<<148, warps * 32>>>(out, cyc, 0.3f);
  k<OP><<<148, warps * 32>>>(out, cyc, 0.3f);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // warps are spread over the 4 sub-partitions: warps/4 per sub-partition
  const double per_sp_instr = static_cast<double>(ITERS) * instr_per_iter * (warps / 4.0);
  printf("%-34s warps/SM=%2d  %.2f cycles per warp-instruction per sub-partition\n", name, warps, h / per_sp_instr);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<8>("ex2.approx.f16x2 (per PTX instr)", 4, CHAINS); run<8>("ex2.approx.f16x2 (per PTX instr)", 16, CHAINS);
  run<9>("ex2.approx.ftz.bf16x2 (per PTX instr)", 4, CHAINS); run<9>("ex2.approx.ftz.bf16x2 (per PTX instr)", 16, CHAINS);
  run<10>("tanh.approx.bf16x2 (per PTX instr)", 4, CHAINS); run<10>("tanh.approx.bf16x2 (per PTX instr)", 16, CHAINS);
  // does FMA-pipe work issue underneath MUFU work of the same warp(s)?  time per ITERATION, in cycles
  for (int w : {4, 8}) {
    run<12>("8 MUFU.EX2 per iteration", w, 1); run<13>("16 FFMA2 per iteration", w, 1); run<11>("8 MUFU.EX2 + 16 FFMA2 per iteration", w, 1);
  }
  for (int w : {4, 8, 16}) {
    if (w == 4) {
      run<0>("MUFU.EX2", 4, CHAINS); run<1>("FFMA", 4, CHAINS); run<2>("FFMA2 (fma.rn.f32x2)", 4, CHAINS / 2);
      run<7>("FADD2 (add.rn.f32x2)", 4, CHAINS / 2); run<3>("FMNMX3 (max.f32 a,b,c)", 4, CHAINS); run<6>("FMNMX", 4, CHAINS);
      run<4>("F2FP.BF16 (cvt.rn.bf16x2.f32)", 4, CHAINS); run<5>("SHL + IADD", 4, 2 * CHAINS);
    } else if (w == 8) {
      run<0>("MUFU.EX2", 8, CHAINS); run<1>("FFMA", 8, CHAINS); run<2>("FFMA2 (fma.rn.f32x2)", 8, CHAINS / 2);
      run<7>("FADD2 (add.rn.f32x2)", 8, CHAINS / 2); run<3>("FMNMX3 (max.f32 a,b,c)", 8, CHAINS); run<6>("FMNMX", 8, CHAINS);
      run<4>("F2FP.BF16 (cvt.rn.bf16x2.f32)", 8, CHAINS); run<5>("SHL + IADD", 8, 2 * CHAINS);
    } else {
      run<0>("MUFU.EX2", 16, CHAINS); run<1>("FFMA", 16, CHAINS); run<2>("FFMA2 (fma.rn.f32x2)", 16, CHAINS / 2);
      run<7>("FADD2 (add.rn.f32x2)", 16, CHAINS / 2); run<3>("FMNMX3 (max.f32 a,b,c)", 16, CHAINS); run<6>("FMNMX", 16, CHAINS);
      run<4>("F2FP.BF16 (cvt.rn.bf16x2.f32)", 16, CHAINS); run<5>("SHL + IADD", 16, 2 * CHAINS);
    }
  }
  return 0;
}
