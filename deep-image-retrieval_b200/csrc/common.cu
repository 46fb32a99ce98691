#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace dirb {

int g_use_pdl = 1;
static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
int64_t launches_total() { return g_launches.load(std::memory_order_relaxed); }

// cuTensorMapEncodeTiled is a driver-API symbol; resolve it through the runtime so the library has no
// link-time dependency on libcuda.so (it must load on a box without a driver for the build/ABI checks).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int encode_tmap_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                   uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  DIRB_REQUIRE(enc != nullptr, DIRB200_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIRB_REQUIRE(r == CUDA_SUCCESS, DIRB200_EDRIVER,
               "cuTensorMapEncodeTiled(2d) failed: %d (inner=%llu outer=%llu stride=%llu box=%u,%u base=%p)", (int)r,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
               box_outer, base);
  return 0;
}

int encode_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int tw, int th, int nb, int es) {
  EncodeTiledFn enc = get_encode();
  DIRB_REQUIRE(enc != nullptr, DIRB200_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(tw * es), (cuuint32_t)(th * es), (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIRB_REQUIRE(r == CUDA_SUCCESS, DIRB200_EDRIVER,
               "cuTensorMapEncodeTiled(nhwc) failed: %d (B=%d H=%d W=%d C=%d box=%d,%d,%d es=%d base=%p)", (int)r, B, H,
               W, C, tw, th, nb, es, base);
  return 0;
}

int encode_tmap_nhwc16(CUtensorMap* m, const void* base, int B, int H, int W, int tw, int th, int nb) {
  EncodeTiledFn enc = get_encode();
  DIRB_REQUIRE(enc != nullptr, DIRB200_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {32, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32};
  cuuint32_t box[4] = {16, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIRB_REQUIRE(r == CUDA_SUCCESS, DIRB200_EDRIVER, "cuTensorMapEncodeTiled(nhwc16) failed: %d (B=%d H=%d W=%d box=%d,%d,%d)",
               (int)r, B, H, W, tw, th, nb);
  return 0;
}

int encode_tmap_2d_sw32(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                        uint32_t box_outer) {
  EncodeTiledFn enc = get_encode();
  DIRB_REQUIRE(enc != nullptr, DIRB200_EDRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {16, box_outer};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DIRB_REQUIRE(r == CUDA_SUCCESS, DIRB200_EDRIVER, "cuTensorMapEncodeTiled(2d sw32) failed: %d", (int)r);
  return 0;
}

}  // namespace dirb

extern "C" {

int dirb200_version(void) { return 100; }
const char* dirb200_last_error(void) { return dirb::last_error(); }

int dirb200_device_check(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    dirb::set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return DIRB200_ENODEVICE;
  }
  DIRB_REQUIRE(device >= 0 && device < n, DIRB200_EINVAL, "device %d out of range (%d devices)", device, n);
  int major = 0;
  DIRB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  DIRB_REQUIRE(major == 10, DIRB200_ENODEVICE, "device %d has compute capability %d.x; this library is sm_100a only",
               device, major);
  return 0;
}

}  // extern "C"
