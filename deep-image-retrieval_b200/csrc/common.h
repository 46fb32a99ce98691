// Host-side plumbing shared by all translation units: error reporting, status codes, TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/dirb200.h"

namespace dirb {

void set_error(const char* fmt, ...);
const char* last_error();

#define DIRB_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::dirb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return static_cast<int>(_e);                                                              \
    }                                                                                           \
  } while (0)

#define DIRB_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      ::dirb::set_error(__VA_ARGS__);   \
      return (code);                    \
    }                                   \
  } while (0)

#define DIRB_TRY(expr)         \
  do {                         \
    int _s = (expr);           \
    if (_s != 0) return _s;    \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// fp16 row-major [outer][inner] matrix, box = box_inner x box_outer elements, 128-byte swizzle.
int encode_tmap_2d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                   uint32_t box_inner, uint32_t box_outer);
// fp16 NHWC activation viewed as (C, W, H, B); box (64, tw*es, th*es, nb) traversed with element stride es
// on W and H (es = conv stride), 128-byte swizzle, out-of-bounds elements read as zero.
int encode_tmap_nhwc(CUtensorMap* m, const void* base, int B, int H, int W, int C, int tw, int th, int nb, int es);
// Same for 16-channel (32-byte) pixels: box (16, tw, th, nb), 32-byte swizzle (space-to-depth stem input).
int encode_tmap_nhwc16(CUtensorMap* m, const void* base, int B, int H, int W, int tw, int th, int nb);
// fp16 row-major [outer][inner] matrix, box = 16 x box_outer, 32-byte swizzle.
int encode_tmap_2d_sw32(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                        uint32_t box_outer);

// Launch `kern` with programmatic stream serialization (PDL): its CTAs may start while the previous kernel of the
// stream drains; the kernel must call pdl_wait() before touching data produced by its predecessor.
extern int g_use_pdl;
template <typename Kern, typename... Args>
inline cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// cudaFuncSetAttribute is per (function, device) and costs ~1-2 us of host time: do it once per device, not per launch.
// `mask` is a function-local static of the launcher template instantiation.
inline bool first_launch_on_device(std::atomic<uint64_t>& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const uint64_t bit = 1ull << (dev & 63);
  if (mask.load(std::memory_order_relaxed) & bit) return false;
  mask.fetch_or(bit, std::memory_order_relaxed);
  return true;
}

// Launch counter (the "gpu_launches" the benchmark reports): every kernel launch of this library bumps it.
void count_launch(int n = 1);
int64_t launches_total();

}  // namespace dirb
