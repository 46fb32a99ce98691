// tcgen05 tile GEMM for sm_100a:  D[128 x BN] = A[128 x K] * B[BN x K]^T, fp16 operands, fp32 accumulation in TMEM.
//
//   warp 0      : TMA producer  (one elected lane) - fills a STAGES-deep smem ring with 128x64 A and BNx64 B tiles
//   warp 1      : MMA issuer    (one elected lane) - tcgen05.mma.cta_group::1.kind::f16, 4 x (K=16) per ring slot;
//                 also owns the TMEM allocation
//   warps 2..5  : epilogue      - tcgen05.ld the accumulator (one row per thread), apply the fused epilogue, store
//
// The A operand is either a flat row-major matrix [M][K] (2-D tensor map) or an NHWC activation read as an
// implicit-GEMM operand: for filter tap (kh,kw) and channel block kc the 128 rows of the tile are a
// (nb x th x tw) patch of output pixels, fetched by ONE 4-D TMA box at input offset (stride*ho0+kh-pad,
// stride*wo0+kw-pad); padding and ragged edges come from TMA out-of-bounds zero fill, stride-2 convolutions
// from the tensor map's element stride.  smem tiles are K-major rows of 128 B with the 128-byte swizzle, which
// is exactly what the UMMA smem descriptor (ptx.cuh: umma_desc_sw128) describes.
//
// Epilogue: y = acc*scale[c] + shift[c] (+ residual) (ReLU) -> fp16 NHWC, one output row per thread, direct
// global stores (conv + BN + add + ReLU).
//
// This one-tile-per-CTA kernel was the first working tcgen05 path and is kept as the A/B baseline (conv_impl = 2)
// and as the simplest statement of the TMA -> UMMA -> TMEM pipeline; the production kernels are the persistent
// ones in conv_pers.cuh / conv_halo.cuh / stem_pers.cuh.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace dirb {

enum { EPI_CONV = 0 };

struct GemmTcParams {
  // A addressing
  int a_spatial;               // 0 flat rows, 1 NHWC patch
  int taps, kw_taps;           // KH*KW, KW
  int cin_blocks;              // K-blocks (of 64) per tap
  int stride, pad;
  int tw, th, nb;              // patch shape, tw*th*nb == 128
  int tiles_w, tiles_h;        // patches per image row / column
  int B, Ho, Wo;               // output extent (spatial mode)
  int M;                       // valid rows of A / D (flat mode)
  int N;                       // valid rows of B = columns of D
  int n_tiles;                 // ceil(N / BN)
  // EPI_CONV
  const float* scale;
  const float* shift;
  const __half* res;
  __half* out;
  int relu;
};

template <int BN, int STAGES>
struct GemmTcSmem {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int AUX_OFF = BAR_OFF + 8 * (2 * STAGES + 1) + 8;   // barriers + tmem slot
  static constexpr int TOTAL = AUX_OFF + 2 * BN * 4 + 1024;            // + scale/shift + alignment slack
};

template <int BN, int STAGES, int EPI, int MINB>
__global__ void __launch_bounds__(192, MINB)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmTcParams p) {
  using L = GemmTcSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
  float* s_scale = reinterpret_cast<float*>(smem + L::AUX_OFF);
  float* s_shift = s_scale + BN;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  const int m_tile = blockIdx.x / p.n_tiles;
  const int k_iters = p.taps * p.cin_blocks;

  int wo0 = 0, ho0 = 0, n0 = 0;
  if (p.a_spatial) {
    const int tx = m_tile % p.tiles_w;
    const int ty = (m_tile / p.tiles_w) % p.tiles_h;
    const int tb = m_tile / (p.tiles_w * p.tiles_h);
    wo0 = tx * p.tw;
    ho0 = ty * p.th;
    n0 = tb * p.nb;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
  if (EPI == EPI_CONV && warp >= 2) {
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      s_scale[i] = p.scale[n_tile * BN + i];
      s_shift[i] = p.shift[n_tile * BN + i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
        uint8_t* sa = smem + s * L::STAGE_BYTES;
        uint8_t* sb = sa + L::A_BYTES;
        const int tap = it / p.cin_blocks;
        const int kc = it - tap * p.cin_blocks;
        if (p.a_spatial) {
          const int kh = tap / p.kw_taps;
          const int kw = tap - kh * p.kw_taps;
          tma_load_4d(sa, &tmA, &full_bar[s], kc * 64, wo0 * p.stride + kw - p.pad, ho0 * p.stride + kh - p.pad, n0);
        } else {
          tma_load_2d(sa, &tmA, &full_bar[s], it * 64, m_tile * 128);
        }
        tma_load_2d(sb, &tmB, &full_bar[s], it * 64, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = umma_idesc_f16(128, BN);
      for (int it = 0; it < k_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + L::A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // 4 x K=16 inside the 128-byte swizzle atom: +32 B = +2 encoded units
          umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        umma_commit(&empty_bar[s]);   // frees the ring slot once these MMAs have read it
      }
      umma_commit(acc_bar);           // accumulator complete
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = quarter * 32 + lane;
    bool valid;
    int64_t pix;
    if (p.a_spatial) {
      const int iw = row % p.tw;
      const int ih = (row / p.tw) % p.th;
      const int ib = row / (p.tw * p.th);
      const int wo = wo0 + iw, ho = ho0 + ih, n = n0 + ib;
      valid = (wo < p.Wo) && (ho < p.Ho) && (n < p.B);
      pix = (static_cast<int64_t>(n) * p.Ho + ho) * p.Wo + wo;
    } else {
      const int m = m_tile * 128 + row;
      valid = m < p.M;
      pix = m;
    }
    mbar_wait(acc_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(taddr + c0, v);
      tmem_ld_wait();
      if (EPI == EPI_CONV) {
        if (valid) {
          const int64_t off = pix * p.N + n_tile * BN + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], s_scale[c0 + j], s_shift[c0 + j]);
          if (p.res != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res + off);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 r = __ldg(rp + q);
              const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = unpack_h2(rr[e]);
                v[q * 8 + e * 2] += f.x;
                v[q * 8 + e * 2 + 1] += f.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          uint4* op = reinterpret_cast<uint4*>(p.out + off);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_h2(v[q * 8 + 0], v[q * 8 + 1]);
            o.y = pack_h2(v[q * 8 + 2], v[q * 8 + 3]);
            o.z = pack_h2(v[q * 8 + 4], v[q * 8 + 5]);
            o.w = pack_h2(v[q * 8 + 6], v[q * 8 + 7]);
            op[q] = o;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
  }
}

template <int BN, int STAGES, int EPI, int MINB>
int gemm_tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmTcParams& p, int64_t m_tiles,
                   cudaStream_t stream) {
  using L = GemmTcSmem<BN, STAGES>;
  auto kern = gemm_tc_kernel<BN, STAGES, EPI, MINB>;
  static bool configured = false;
  if (!configured) {
    DIRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    configured = true;
  }
  const int64_t grid = m_tiles * p.n_tiles;
  DIRB_REQUIRE(grid > 0 && grid < (int64_t(1) << 31), DIRB200_ENOTSUP, "gemm_tc grid %lld out of range",
               (long long)grid);
  kern<<<static_cast<unsigned>(grid), 192, L::TOTAL, stream>>>(tmA, tmB, p);
  count_launch();
  DIRB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace dirb
