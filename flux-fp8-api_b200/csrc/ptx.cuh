// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM),
// shared-memory matrix descriptors and instruction descriptors.
//
// Everything in this file is hand-written PTX; no CUTLASS/CuTe is included.  Bit layouts of the
// descriptors follow the PTX ISA "tcgen05 matrix descriptor" / "instruction descriptor" tables.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

namespace fb {

#ifndef FLUXB200_HANG_TRAP_NS
// A waiter that has spun for this long traps instead of hanging the GPU box (0 disables).
#define FLUXB200_HANG_TRAP_NS 4000000000ull
#endif

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// Explicit shared-space accesses by 32-bit shared address.  The dynamic shared buffer is re-aligned with integer
// arithmetic, after which the compiler only knows a GENERIC pointer and emits LD.E / ST.E through the global-memory
// pipe (and its LG throttle) for what are shared-memory reads.
__device__ __forceinline__ float4 lds_f32x4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}

// 256-bit global accesses (LDG/STG.E.ENL2.256, sm_100+, PTX 8.8): one instruction moves a thread's whole 32-byte L2
// sector.  The GEMM epilogues are thread == output row, so a warp's access touches 32 different rows: with 16-byte
// accesses every instruction half-fills 32 sectors and the L2 sees twice the transactions.  32-byte aligned addresses only.
struct u32x8 {
  uint32_t v[8];
};
__device__ __forceinline__ u32x8 ldg_v8(const void* p) {
  u32x8 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ u32x8 ldg_nc_v8(const void* p) {  // read-only data (never written by this grid)
  u32x8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                       uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4),
               "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {  // one FMNMX3
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// Packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 issue two IEEE fp32 operations per lane per instruction; each half
// rounds exactly like the scalar instruction).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// Register re-partitioning between warpgroups (all four warps of a warpgroup execute it together).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// programmatic dependent launch (no-ops when the kernel was launched without the PDL attribute)
// ---------------------------------------------------------------------------------------------
// Blocks until every kernel this one depends on has completed and its memory is visible.  Everything before it
// (barrier init, TMEM allocation, tensor-map prefetch) overlaps the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Allows the next kernel in the stream to be scheduled once all CTAs of this grid have called it (or exited).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy writes to shared memory -> visible to the async proxy (TMA / tcgen05.mma reads)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_global() {
  // generic-proxy accesses to global memory <-> async-proxy (TMA) accesses to the same locations
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Grid-wide barrier for a grid whose CTAs are all co-resident (persistent kernels: <= 1 CTA per SM, grid <= #SMs).
// `ws` points to two zero-initialised 32-bit words (arrive counter, generation) that only kernels of ONE stream use:
// launches never overlap inside the barrier (it sits behind griddepcontrol.wait), the last arriver re-arms the
// counter, so the words need no host-side reset between launches or CUDA-graph replays.  One thread per CTA calls it.
__device__ __forceinline__ void grid_barrier(unsigned int* ws, unsigned int num_ctas) {
  unsigned int g0;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g0) : "l"(ws + 1) : "memory");
  __threadfence();
  const unsigned int old = atomicAdd(ws, 1u);
  if (old == num_ctas - 1) {
    atomicExch(ws, 0u);
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ws + 1) : "memory");
  } else {
    unsigned int g = g0;
#if FLUXB200_HANG_TRAP_NS
    const uint64_t t0 = globaltimer_ns();
    uint32_t spins = 0;
#endif
    while (g == g0) {
      __nanosleep(64);
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(ws + 1) : "memory");
#if FLUXB200_HANG_TRAP_NS
      if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > FLUXB200_HANG_TRAP_NS) {
        printf("fluxb200: grid barrier timed out (block %d of %u)\n", blockIdx.x, num_ctas);
        __trap();
      }
#endif
    }
  }
  __threadfence();
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Blocks until the phase with the given parity has completed.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if FLUXB200_HANG_TRAP_NS
  uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
#endif
  while (!mbar_try_wait(bar, parity)) {
#if FLUXB200_HANG_TRAP_NS
    if ((++spins & 0x3ff) == 0 && globaltimer_ns() - t0 > FLUXB200_HANG_TRAP_NS) {
      printf("fluxb200: mbarrier wait timed out (block %d,%d thread %d bar@%u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// L2 cache-hint policies (createpolicy encodings used by CUTLASS' TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1, int32_t c2, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}

// 2-SM (cta_group::2) form: data lands in the issuing CTA's shared memory, the transaction bytes are
// signalled on the barrier at the same offset in the pair's leader CTA (peer bit 24 of the
// shared::cluster address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0,
                                                int32_t c1, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                                int32_t c2, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                                int32_t c2, int32_t c3, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "l"(hint)
      : "memory");
}

// 2-SM form with cluster multicast: the box lands at the same shared-memory offset in every CTA of `cta_mask`, and each
// destination signals the transaction bytes on the barrier (same offset) of ITS pair's leader CTA.
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                                   uint16_t cta_mask, uint64_t hint = kEvictNormal) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      ".L2::cache_hint [%0], [%1, {%4, %5}], [%2], %3, %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar) & kPeerBitMask), "h"(cta_mask), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA in the cluster
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}

// Same, without release semantics.  `.release.cluster` compiles to MEMBAR.ALL.GPU in front of the arrive: the warp
// then sits until every global store it has issued is acknowledged by L2.  Handing a TMEM accumulator back needs no
// memory ordering at all (tcgen05.wait::ld has already put the data in registers; tcgen05.fence::before_thread_sync
// orders the tensor-memory reads), so the GEMM epilogue uses this form and lets its stores drain in the background.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta));
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, fences, commit
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {  // whole warp
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// cta_group::2 forms: executed by the same warp index in both CTAs of the pair
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// All previously issued cta_group::2 MMAs arrive on the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// All previously issued tcgen05.mma of this thread arrive on `bar` when complete (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, operand stored as rows of 128 bytes
// (8-row x 128-byte swizzle atoms, atoms 1024 B apart).
//  * K-major operand: a row is one M/N index, the 128 bytes run along K.  SBO = 1024 (next 8 rows).
//  * MN-major operand: a row is one K index, the 128 bytes run along M/N (64 x 16-bit).
//    SBO = 1024 (next 8 K rows), LBO = byte distance to the next 64-element M/N chunk.
// bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SW128)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

// Advance a descriptor's start address by `byte_off` (a multiple of 16; the 14-bit address field never carries
// because shared memory is < 256 KB).  Lets the MMA issuer precompute one descriptor per tile and pay a single
// 64-bit add per MMA instead of rebuilding the bit-fields.
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t byte_off) {
  return desc + static_cast<uint64_t>(byte_off >> 4);
}

// Instruction descriptor (upper 32 bits of the "idesc" operand).
// bits: [4,6) D fmt (1=f32) | [7,10) A fmt | [10,13) B fmt | 15 A major (1=MN) | 16 B major (1=MN)
//       [17,23) N>>3 | [24,29) M>>4
enum : uint32_t { kFmtE4M3 = 0, kFmtE5M2 = 1 };  // kind::f8f6f4
enum : uint32_t { kFmtF16 = 0, kFmtBF16 = 1 };   // kind::f16
__host__ __device__ constexpr uint32_t make_idesc(uint32_t a_fmt, uint32_t b_fmt, uint32_t M, uint32_t N,
                                                   uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
         ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void mma_f8f6f4_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// cta_group::2: M = 256 over the SM pair (A rows 0-127 from the leader's smem, 128-255 from the peer's, B's N
// halves likewise); issued by the leader CTA only.
__device__ __forceinline__ void mma_f8f6f4_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Compile-time accumulate flag (no setp in the issue stream).
template <bool ACC>
__device__ __forceinline__ void mma_f16_ss_c(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  if constexpr (ACC)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
                 "l"(a_desc), "l"(b_desc), "r"(idesc)
                 : "memory");
}
// cta_group::2 forms (M = 256 over the SM pair; issued by the leader CTA)
__device__ __forceinline__ void mma_f16_ss_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts_2sm(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand from tensor memory (TS form).
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMEM <-> registers.  32x32b shape: lane i of warp w reads TMEM lane 32*(w%4)+i, N consecutive
// 32-bit columns starting at the column in taddr.  taddr = (lane << 16) | column.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
        "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
        "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]),
      "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
      "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {  // 16 consecutive columns
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// numeric helpers shared by the epilogues.  bf16 round-trips reproduce the reference's eager
// "every op rounds to bf16" semantics.
// ---------------------------------------------------------------------------------------------
// Round-to-nearest-even to bf16, result kept as fp32.  cvt.rn.bf16x2.f32 (SASS F2FP, full-rate ALU pipe) with a
// zero low half yields the rounded value directly as fp32 bits; the scalar cvt.rn.bf16.f32 compiles to F2F,
// which issues at quarter rate through the XU pipe and dominated the LayerNorm / epilogue instruction mix.
__device__ __forceinline__ float bf16r(float x) {
  uint32_t u;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x), "f"(0.f));
  return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
// sm_100a mixed-precision scalar arithmetic (PTX add / fma .f32.bf16, SASS FHADD.BF16 / FHFMA.BF16): a bf16 operand --
// either half of a packed pair, selected in the instruction, no unpack -- enters an fp32 add or multiply-add.  The result
// is the fp32 operation on the exactly converted operand(s): bit-identical to unpack + FADD / FFMA, one instruction fewer
// per element.
__device__ __forceinline__ float add_f32_bf16lo(uint32_t pair, float c) {
  float d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, lo, %2;\n\t}" : "=f"(d) : "r"(pair), "f"(c));
  return d;
}
__device__ __forceinline__ float add_f32_bf16hi(uint32_t pair, float c) {
  float d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tadd.rn.f32.bf16 %0, hi, %2;\n\t}" : "=f"(d) : "r"(pair), "f"(c));
  return d;
}
// c + a * a for the low / high bf16 of a pair, fp32 (FHFMA.BF16: exact product, one rounding -- the same value as
// fmaf(float(a), float(a), c))
__device__ __forceinline__ float fma_sq_f32_bf16lo(uint32_t pair, float c) {
  float d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.bf16 %0, lo, lo, %2;\n\t}" : "=f"(d) : "r"(pair), "f"(c));
  return d;
}
__device__ __forceinline__ float fma_sq_f32_bf16hi(uint32_t pair, float c) {
  float d;
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %1;\n\tfma.rn.f32.bf16 %0, hi, hi, %2;\n\t}" : "=f"(d) : "r"(pair), "f"(c));
  return d;
}
// (shift + mask: two ALU instructions per pair; `__bfloat1622float2` compiles to PRMT + 2 SHF, three)
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// fp8 cast of a value that is already clamped to the format's finite range: RNE, saturating.
template <int FMT>  // 0 = e4m3, 1 = e5m2
__device__ __forceinline__ uint8_t to_fp8(float x) {
  return static_cast<uint8_t>(
      __nv_cvt_float_to_fp8(x, __NV_SATFINITE, FMT == 0 ? __NV_E4M3 : __NV_E5M2));
}
template <int FMT>
__device__ __forceinline__ uint16_t to_fp8x2(float lo, float hi) {
  return static_cast<uint16_t>(
      __nv_cvt_float2_to_fp8x2(make_float2(lo, hi), __NV_SATFINITE, FMT == 0 ? __NV_E4M3 : __NV_E5M2));
}
// reference: to_fp8_saturated(x, s, max).to(fp8) with bf16 x  (float8_quantize.py:217-218,274-276):
// fp8( clamp( bf16(x*s), -max, max ) )
template <int FMT>
__device__ __forceinline__ float quant_pre(float x_bf16_exact, float scale) {
  constexpr float kMax = FMT == 0 ? 448.f : 57344.f;
  float p = bf16r(x_bf16_exact * scale);
  return fminf(fmaxf(p, -kMax), kMax);
}
__device__ __forceinline__ float exp2f_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// The same quantiser on a PAIR, for a scale that is itself a bf16 value (always true under torch's CUDA scalar
// semantics): bf16(x*s) is one packed HMUL2 (the fp32 product of two bf16 values is exact, so
// rounding it once equals the reference's fp32 multiply + bf16 rounding), and the saturating fp8 conversion
// subsumes the clamp.  x pair given as fp32 values still to be rounded to bf16.
template <int FMT>
__device__ __forceinline__ uint16_t quant_pair_bf16scale(float x0, float x1, __nv_bfloat162 s2) {
  const __nv_bfloat162 p = __hmul2_rn(__floats2bfloat162_rn(x0, x1), s2);
  const float2 pf = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(&p));
  return to_fp8x2<FMT>(pf.x, pf.y);
}
// 0.5 x (1 + tanh(u)) == x * sigmoid(2u) == x / (1 + 2^(-2 u log2 e)):  8 instructions instead of 12.
// u = sqrt(2/pi) (x + 0.044715 x^3).  Saturates cleanly: 2^(+big) = inf -> x * 0, 2^(-big) = 0 -> x * 1.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  constexpr float kC0 = -2.f * 1.4426950408889634f * 0.7978845608028654f;            // -2 log2(e) sqrt(2/pi)
  constexpr float kC1 = kC0 * 0.044715f;
  const float w = fmaf(x * x, kC1, kC0);
  const float e = exp2f_approx(x * w);
  return x * rcp_approx(1.f + e);
}
// The same on a pair with packed fp32 arithmetic (FMUL2 / FFMA2 / FADD2: each half rounds exactly like the scalar
// instruction, so the result is bit-identical to two gelu_tanh_fast calls; 5 packed + 4 MUFU instructions per pair
// instead of 10 + 4).
__device__ __forceinline__ float2 gelu_tanh_fast2(float2 x) {
  constexpr float kC0 = -2.f * 1.4426950408889634f * 0.7978845608028654f;
  constexpr float kC1 = kC0 * 0.044715f;
  const float2 w = ffma2(fmul2(x, x), make_float2(kC1, kC1), make_float2(kC0, kC0));
  const float2 u = fmul2(x, w);
  const float2 d = fadd2(make_float2(1.f, 1.f), make_float2(exp2f_approx(u.x), exp2f_approx(u.y)));
  return fmul2(x, make_float2(rcp_approx(d.x), rcp_approx(d.y)));
}
// nn.GELU(approximate="tanh") evaluated in fp32 (ATen opmath) on a bf16-exact input.
__device__ __forceinline__ float gelu_tanh(float x) {
  const float kBeta = 0.7978845608028654f;  // sqrt(2/pi)
  const float kKappa = 0.044715f;
  float inner = kBeta * (x + kKappa * x * x * x);
  // tanh(u) = 1 - 2/(exp(2u)+1); saturates cleanly for |u| large.
  float e = __expf(2.f * inner);
  float t = 1.f - __fdividef(2.f, e + 1.f);
  return 0.5f * x * (1.f + t);
}

}  // namespace fb
