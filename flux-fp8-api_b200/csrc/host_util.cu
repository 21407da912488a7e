#include "host_util.h"

#include <cstdlib>
#include <mutex>

namespace fb {

std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FLUXB200_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static CUtensorMapDataType dtype_of(int elem_bytes) {
  return elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(FLUXB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (pitch_bytes & 15))
    return set_error(FLUXB200_ERR_INVALID, "TMA operand must be 16-byte aligned (ptr %p pitch %llu)", base,
                     (unsigned long long)pitch_bytes);
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype_of(elem_bytes), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FLUXB200_ERR_CUDA, "cuTensorMapEncodeTiled(2d %llux%llu pitch %llu box %ux%u) -> %d",
                     (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)pitch_bytes,
                     box_rows, box_cols, (int)r);
  return 0;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(FLUXB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (stride1_bytes & 15) || (stride2_bytes & 15))
    return set_error(FLUXB200_ERR_INVALID, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, dtype_of(elem_bytes), 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FLUXB200_ERR_CUDA, "cuTensorMapEncodeTiled(3d) -> %d", (int)r);
  return 0;
}

int make_tmap_4d(CUtensorMap* out, const void* base, int elem_bytes, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4], const uint32_t* elem_strides) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(FLUXB200_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides_bytes[0] & 15) || (strides_bytes[1] & 15) || (strides_bytes[2] & 15))
    return set_error(FLUXB200_ERR_INVALID, "TMA operand must be 16-byte aligned");
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (elem_strides != nullptr)
    for (int i = 0; i < 4; ++i) estr[i] = elem_strides[i];
  CUresult r = fn(out, dtype_of(elem_bytes), 4, const_cast<void*>(base), d, st, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(FLUXB200_ERR_CUDA, "cuTensorMapEncodeTiled(4d %llux%llux%llux%llu box %ux%ux%ux%u) -> %d",
                     (unsigned long long)dims[3], (unsigned long long)dims[2], (unsigned long long)dims[1],
                     (unsigned long long)dims[0], box[3], box[2], box[1], box[0], (int)r);
  return 0;
}

}  // namespace fb

extern "C" {

int fluxb200_version(void) { return FLUXB200_VERSION; }

const char* fluxb200_last_error(void) { return fb::last_error_ref().c_str(); }

int fluxb200_device_check(int* sm_count_out) {
  int dev = 0;
  FB_CUDA_OK(cudaGetDevice(&dev));
  int major = 0, minor = 0, sms = 0;
  FB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  FB_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  FB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count_out) *sm_count_out = sms;
  if (major != 10)
    return fb::set_error(FLUXB200_ERR_UNSUPPORTED, "device %d is sm_%d%d; libflux_b200 needs sm_100 (B200)", dev,
                         major, minor);
  return 0;
}

}  // extern "C"
