// Host-side helpers shared by the C-ABI translation units: per-thread error string, CUDA error
// checks that never throw, TMA tensor-map encoding through the driver entry point (the library
// does not link libcuda, so it loads on machines without a driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "flux_b200.h"

namespace fb {

std::string& last_error_ref();
int set_error(int code, const char* fmt, ...);
int sm_count();

#define FB_CUDA_OK(expr)                                                                        \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fb::set_error(FLUXB200_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                           __FILE__, __LINE__);                                                 \
  } while (0)

#define FB_REQUIRE(cond, ...)                                          \
  do {                                                                 \
    if (!(cond)) return fb::set_error(FLUXB200_ERR_INVALID, __VA_ARGS__); \
  } while (0)

// Programmatic dependent launch for our kernels (env FLUXB200_PDL=0 disables).
bool pdl_enabled();

// cudaLaunchKernelEx with an optional thread-block cluster (x dimension) and the PDL attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// 2-D row-major tensor [rows, cols] of `elem_bytes`-wide elements, row pitch `pitch_bytes`,
// box = [box_rows, box_cols], SWIZZLE_128B (box_cols*elem_bytes must be 128).
int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols);
// 3-D tensor [d2, d1, d0] (d0 innermost, contiguous), strides in bytes for d1 and d2.
int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);
// 4-D tensor [d3, d2, d1, d0] (d0 innermost, contiguous; NHWC activations: d0 = C, d1 = W, d2 = H, d3 = B), strides in
// bytes for d1..d3, SWIZZLE_128B (box0 * elem_bytes must be 128), out-of-bounds elements (negative coordinates
// included) read as zero -- the zero padding of a convolution.
// elem_strides (optional): traversal stride per dimension (a box of box[i] elements loads ceil(box[i] / stride) of them).
int make_tmap_4d(CUtensorMap* out, const void* base, int elem_bytes, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4], const uint32_t* elem_strides = nullptr);

}  // namespace fb
