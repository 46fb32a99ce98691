// Bilinear resize of uint8 HWC images, bit-exact with PIL's Image.resize(size, Image.BILINEAR).
//
// The reference rescales images on the CPU with PIL before the network (dirtorch/utils/transforms.py:133-185,
// `Scale`, img.resize(size2, Image.BILINEAR) at :183); multi-scale extraction (scales 0.7 / 1.4, BASELINE configs[4])
// makes that resize the next bottleneck after the GPU path.  PIL's resampler is a separable
// triangle filter whose support grows with the down-scaling factor (anti-aliasing), evaluated in fixed point: per output
// coordinate a [xmin, xmin+n) window of coefficients normalised to sum 1 and rounded to 22 fractional bits, a
// horizontal pass to an 8-bit intermediate (accumulator starts at 1 << 21, result clipped to [0,255]) and a vertical
// pass of the same form.  The coefficient tables are computed on the host in double precision exactly as PIL does;
// the two kernels apply them in 32-bit integer arithmetic, so the output is identical byte for byte.
#include <math.h>

#include <vector>

#include "conv.h"

namespace dirb {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

void resample_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int* ksize_out) {
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;                 // bilinear: support 1
  const int ksize = static_cast<int>(ceil(support)) * 2 + 1;
  bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  kk.assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> pre(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      pre[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? pre[x] / ww : pre[x];
      kk[static_cast<size_t>(xx) * ksize + x] =
          v < 0 ? static_cast<int>(-0.5 + v * (1 << PRECISION_BITS)) : static_cast<int>(0.5 + v * (1 << PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  *ksize_out = ksize;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [B][H][W][3] -> out [B][H][Wo][3]
__global__ void resize_h_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W, int Wo,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over B*H*Wo
  if (i >= total) return;
  const int xx = static_cast<int>(i % Wo);
  const int64_t row = i / Wo;                                                      // b*H + y
  const int x0 = bounds[2 * xx], n = bounds[2 * xx + 1];
  const uint8_t* src = in + (row * W + x0) * 3;
  const int* k = kk + static_cast<int64_t>(xx) * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int c = k[x];
    s0 += src[3 * x] * c;
    s1 += src[3 * x + 1] * c;
    s2 += src[3 * x + 2] * c;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = clip8(s0);
  dst[1] = clip8(s1);
  dst[2] = clip8(s2);
}

// in [B][H][Wo][3] -> out [B][Ho][Wo][3]
__global__ void resize_v_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int Ho, int Wo,
                                const int* __restrict__ bounds, const int* __restrict__ kk, int ksize, int64_t total) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;   // over B*Ho*Wo
  if (i >= total) return;
  const int xx = static_cast<int>(i % Wo);
  const int yy = static_cast<int>((i / Wo) % Ho);
  const int64_t b = i / (static_cast<int64_t>(Wo) * Ho);
  const int y0 = bounds[2 * yy], n = bounds[2 * yy + 1];
  const uint8_t* src = in + ((b * H + y0) * Wo + xx) * 3;
  const int* k = kk + static_cast<int64_t>(yy) * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const int c = k[y];
    const uint8_t* p = src + static_cast<int64_t>(y) * Wo * 3;
    s0 += p[0] * c;
    s1 += p[1] * c;
    s2 += p[2] * c;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = clip8(s0);
  dst[1] = clip8(s1);
  dst[2] = clip8(s2);
}

}  // namespace
}  // namespace dirb

using namespace dirb;

extern "C" {

// Host-only: PIL's fixed-point bilinear coefficient table for one axis.  Call with kk_out == NULL to get ksize.
int dirb200_resize_coeffs(int in_size, int out_size, int* bounds_out, int* kk_out, int* ksize_out) {
  DIRB_REQUIRE(in_size > 0 && out_size > 0 && ksize_out, DIRB200_EINVAL, "bad arguments");
  std::vector<int> b, k;
  resample_coeffs(in_size, out_size, b, k, ksize_out);
  if (bounds_out) std::copy(b.begin(), b.end(), bounds_out);
  if (kk_out) std::copy(k.begin(), k.end(), kk_out);
  return 0;
}

int dirb200_resize_bilinear_u8(const uint8_t* in_dev, int B, int H, int W, int Ho, int Wo, uint8_t* out_dev,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DIRB_REQUIRE(in_dev && out_dev && B > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, DIRB200_EINVAL, "bad arguments");
  std::vector<int> bx, kx, by, ky;
  int ksx = 0, ksy = 0;
  resample_coeffs(W, Wo, bx, kx, &ksx);
  resample_coeffs(H, Ho, by, ky, &ksy);
  const size_t n_tab = bx.size() + kx.size() + by.size() + ky.size();
  int* tab = nullptr;
  uint8_t* tmp = nullptr;
  DIRB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tab), n_tab * sizeof(int), stream));
  DIRB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&tmp), static_cast<size_t>(B) * H * Wo * 3, stream));
  int* d_bx = tab;
  int* d_kx = d_bx + bx.size();
  int* d_by = d_kx + kx.size();
  int* d_ky = d_by + by.size();
  DIRB_CUDA(cudaMemcpyAsync(d_bx, bx.data(), bx.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  DIRB_CUDA(cudaMemcpyAsync(d_kx, kx.data(), kx.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  DIRB_CUDA(cudaMemcpyAsync(d_by, by.data(), by.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  DIRB_CUDA(cudaMemcpyAsync(d_ky, ky.data(), ky.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
  const int64_t t1 = static_cast<int64_t>(B) * H * Wo, t2 = static_cast<int64_t>(B) * Ho * Wo;
  resize_h_kernel<<<static_cast<unsigned>(ceil_div(t1, 256)), 256, 0, stream>>>(in_dev, tmp, H, W, Wo, d_bx, d_kx, ksx, t1);
  resize_v_kernel<<<static_cast<unsigned>(ceil_div(t2, 256)), 256, 0, stream>>>(tmp, out_dev, H, Ho, Wo, d_by, d_ky, ksy, t2);
  count_launch(2);
  DIRB_CUDA(cudaGetLastError());
  DIRB_CUDA(cudaFreeAsync(tmp, stream));
  DIRB_CUDA(cudaFreeAsync(tab, stream));
  return 0;
}

}  // extern "C"
