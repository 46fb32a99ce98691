// Persistent tcgen05 kernel for the ResNet stem: conv 7x7 / stride 2 / pad 3 (3 -> 64 channels) + BN + ReLU.
// Reference: dirtorch/nets/backbones/resnet.py:115-118,158-160.
//
// The input is first rewritten (s2d_kernel in ops.cu) as a zero-padded "space-to-depth" tensor
//     S[n][Y][X][(dy*2+dx)*4 + c] = img[n][c][2Y+dy-3][2X+dx-3]        (c < 3; channel 3 and the border are 0)
// with 16 fp16 channels per 2x2 pixel block.  The 7x7/s2 convolution over img is then a 4x4/s1 convolution over S:
//     out[ho][wo][co] = sum_{a,b in 0..3} sum_{j<16} S[ho+a][wo+b][j] * W2[co][a][b][j],
//     W2[co][a][b][(dy*2+dx)*4+c] = w[co][c][2a+dy][2b+dx]              (0 where 2a+dy or 2b+dx = 7, or c = 3)
// i.e. an implicit GEMM with K = 16 taps x 16 = 256.  Each tap is ONE tcgen05.mma (M=128, N=64, K=16) whose B
// operand is the 64 x 16 slice of W2 for that tap (all 16 slices, 32 KB, stay resident in shared memory) and whose
// A operand is a VIEW of the tile's input patch: the output tile is 8 x 16 pixels, ONE 4-D TMA box brings the
// 11 x 19 halo patch of S (32-byte pixels, SWIZZLE_32B, 6.5 KB) into shared memory, and tap (a,b) reads it through
// a UMMA descriptor starting (a*11 + b) * 32 B into the patch with 8-pixel core matrices 11 * 32 B apart
// (same shifted-view mechanism as conv_halo.cuh) - the patch crosses the L2 -> SM path once instead of 16 times.
//
//   warp 0  TMA producer   weights once; per tile one halo patch into a ring slot
//   warp 1  MMA issuer     16 MMAs per tile into one of two TMEM accumulators (64 columns each)
//   warp 3  TMEM allocator
//   warps 4-7 epilogue     TMEM -> scale/shift/ReLU -> fp16 into a 128-byte-swizzled staging tile -> TMA store
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace dirb {

struct StemParams {
  int tw, th, nb, tiles_w, tiles_h, total_tiles;
  const float* scale;
  const float* shift;
};

struct StemSmem {
  static constexpr int STAGES = 4;                       // ring slots (one slot = the halo patch of one tile)
  static constexpr int HALO_W = 11, HALO_H = 19;
  static constexpr int HALO_DATA = HALO_W * HALO_H * 32; // 6 688 B landed by TMA
  static constexpr int SLOT_BYTES = 7 * 1024;
  static constexpr int W_BYTES = 16 * 64 * 32;           // 16 taps x (64 x 16 fp16)
  static constexpr int STG_BYTES = 128 * 128;            // 128 pixels x 64 channels fp16
  static constexpr int W_OFF = STAGES * SLOT_BYTES;
  static constexpr int STG_OFF = W_OFF + W_BYTES;
  static constexpr int BAR_OFF = STG_OFF + 2 * STG_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 1;
  static constexpr int TOTAL = BAR_OFF + 8 * NUM_BARS + 16 + 1024;
};

// K-major operand with rows of 16 halfs (32 B), 32-byte swizzle: 8-row groups `sbo` bytes apart (256 when dense).
__device__ __forceinline__ uint64_t umma_desc_sw32(uint32_t smem_addr, uint32_t sbo = 256) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(6) << 61;   // SWIZZLE_32B
  return d;
}

__global__ void __launch_bounds__(256, 1)
stem_pers_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmO, const StemParams p) {
  using L = StemSmem;
  constexpr int STAGES = L::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem + L::W_OFF;
  uint8_t* stg = smem + L::STG_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_bar = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4);
    }
    mbar_init(w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may start its own prologue while this one runs / drains
  pdl_wait();                // everything below reads tensors written by the previous kernel of the stream

  auto tile_origin = [&](int t, int& wo0, int& ho0, int& n0) {
    const int tx = t % p.tiles_w;
    const int ty = (t / p.tiles_w) % p.tiles_h;
    const int tb = t / (p.tiles_w * p.tiles_h);
    wo0 = tx * p.tw;
    ho0 = ty * p.th;
    n0 = tb * p.nb;
  };

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_bar, L::W_BYTES);
      for (int tap = 0; tap < 16; ++tap) tma_load_2d(wsm + tap * 2048, &tmW, w_bar, tap * 16, 0);
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int wo0, ho0, n0;
        tile_origin(t, wo0, ho0, n0);
        const int s = g % STAGES;
        mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], L::HALO_DATA);
        tma_load_4d(smem + s * L::SLOT_BYTES, &tmS, &full_bar[s], 0, wo0, ho0, n0);
        ++g;
      }
    }
  } else if (warp == 1) {
    // whole warp walks the loop (uniform operands), one elected lane issues the 16 MMAs + commits of a tile
    constexpr uint32_t idesc = umma_idesc_f16(128, 64);
    mbar_wait(w_bar, 0);
    tc_fence_after();
    const uint64_t bdesc0 = umma_desc_sw32(smem_u32(wsm));
    uint32_t g = 0, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i, ++g) {
      const uint32_t acc = i & 1;
      mbar_wait(&acc_empty[acc], ((i >> 1) & 1) ^ 1);
      const int s = g % STAGES;
      mbar_wait(&full_bar[s], (g / STAGES) & 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 64;
      const uint64_t adesc0 = umma_desc_sw32(smem_u32(smem + s * L::SLOT_BYTES), L::HALO_W * 32u);
      if (elect_one()) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            umma_f16(d_tmem, adesc0 + static_cast<uint64_t>((a * L::HALO_W + b) * 2),
                     bdesc0 + static_cast<uint64_t>((a * 4 + b) * 128), idesc, (a | b) != 0);
        umma_commit(&empty_bar[s]);
        umma_commit(&acc_full[acc]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int quarter = warp - 4;
    const int row = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 128);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint32_t i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      int wo0, ho0, n0;
      tile_origin(t, wo0, ho0, n0);
      const uint32_t acc = i & 1;
      uint8_t* buf = stg + (i & 1) * L::STG_BYTES;
      mbar_wait(&acc_full[acc], (i >> 1) & 1);
      tc_fence_after();
      if (leader) bulk_wait_read<1>();                  // the store issued two tiles ago has left this buffer
      named_bar_sync(1, 128);
      const uint32_t taddr = tmem_base + acc * 64 + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[32];
        tmem_ld32(taddr + half * 32, v);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
        const float4* sc4 = reinterpret_cast<const float4*>(p.scale + half * 32);
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift + half * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s = __ldg(sc4 + q), h = __ldg(sh4 + q);
          v[4 * q + 0] = fmaxf(fmaf(v[4 * q + 0], s.x, h.x), 0.f);
          v[4 * q + 1] = fmaxf(fmaf(v[4 * q + 1], s.y, h.y), 0.f);
          v[4 * q + 2] = fmaxf(fmaf(v[4 * q + 2], s.z, h.z), 0.f);
          v[4 * q + 3] = fmaxf(fmaf(v[4 * q + 3], s.w, h.w), 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = static_cast<uint32_t>(half * 4 + j);
          uint4 o;
          o.x = pack_h2(v[j * 8 + 0], v[j * 8 + 1]);
          o.y = pack_h2(v[j * 8 + 2], v[j * 8 + 3]);
          o.z = pack_h2(v[j * 8 + 4], v[j * 8 + 5]);
          o.w = pack_h2(v[j * 8 + 6], v[j * 8 + 7]);
          *reinterpret_cast<uint4*>(buf + row_off + ((chunk ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (leader) {
        tma_store_4d(&tmO, buf, 0, wo0, ho0, n0);
        bulk_commit();
      }
    }
    if (leader) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace dirb
