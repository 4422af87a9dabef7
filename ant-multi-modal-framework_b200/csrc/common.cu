// b200mm — host-side helpers shared by all C-ABI entry points.
#include "common.cuh"

#include <stdarg.h>
#include <string.h>

#include <mutex>

namespace b200mm {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  return B200MM_OK;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled fn = get_encode_fn();
  B200MM_REQUIRE(fn != nullptr, B200MM_ERR_LAUNCH, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  B200MM_REQUIRE((reinterpret_cast<uintptr_t>(gptr) & 15) == 0, B200MM_ERR_ALIGN, "TMA operand base %p not 16B aligned", gptr);
  B200MM_REQUIRE((pitch_elems * 2) % 16 == 0, B200MM_ERR_ALIGN, "TMA operand pitch %llu elems not a multiple of 8",
                 (unsigned long long)pitch_elems);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200MM_REQUIRE(r == CUDA_SUCCESS, B200MM_ERR_LAUNCH,
                 "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu box=%ux%u", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_elems, box_inner, box_outer);
  return B200MM_OK;
}

}  // namespace b200mm

extern "C" {

const char* b200mm_last_error(void) { return b200mm::g_last_error; }

int b200mm_version(void) { return 100; }

int b200mm_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    b200mm::set_last_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return B200MM_ERR_LAUNCH;
  }
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    b200mm::set_last_error("device %d has compute capability major %d; b200mm kernels are sm_100a only", dev, major);
    return B200MM_ERR_ARCH;
  }
  return B200MM_OK;
}

}  // extern "C"
