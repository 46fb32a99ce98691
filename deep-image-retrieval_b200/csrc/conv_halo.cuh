// Persistent tcgen05 kernel for 3x3 / stride-1 / pad-1 convolutions with the input patch loaded ONCE per tile.
//
// conv_pers.cuh fetches the A operand of a 3x3 convolution tap by tap: nine 16 KB TMA boxes per 64-channel block,
// i.e. the same activation bytes cross the L2 -> SM path nine times.  Here the output tile is a fixed 8 (wide) x 16
// (tall) pixel patch, and for each 64-channel block ONE TMA box brings the 10 x 18 halo patch (22.5 KB) into shared
// memory.  The A operand of tap (kh,kw) is then just a different VIEW of that buffer: rows = pixels (y+kh, x+kw),
// i.e. start address + (kh*10 + kw) * 128 B, 8 consecutive pixels of a patch row = one 8-row core matrix, consecutive
// patch rows 10 * 128 B = 1280 B apart (the descriptor's stride byte offset).  The 128-byte swizzle is a function of
// the shared-memory address bits, so TMA (writer) and UMMA (reader) agree for any 128-byte-aligned start - the same
// mechanism as the +32 B K-advance inside a swizzled row.  Weights stream through a ring of (tap, channel-block)
// tiles, or stay resident in shared memory when the whole filter fits (layer1: 64 x 576 fp16 = 72 KB).
//
// Epilogue, accumulator double buffering and warp roles are those of conv_pers.cuh (conv_epilogue_tile).
#pragma once
#include "conv_pers.cuh"

namespace dirb {

template <int BN, int BSTAGES, int NB, int NA_ = 2>
struct ConvHaloSmem {
  static constexpr int HALO_W = 10, HALO_H = 18;
  static constexpr int HALO_DATA = HALO_W * HALO_H * 128;          // 23 040 B landed by TMA
  static constexpr int HALO_SLOT = 24 * 1024;                      // 1024-byte aligned slot
  static constexpr int NA = NA_;                                   // halo slots (prefetch depth of the input patches)
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STG_BYTES = 128 * 128;
  static constexpr int B_OFF = NA * HALO_SLOT;
  static constexpr int STG_OFF = B_OFF + BSTAGES * B_BYTES;
  static constexpr int BAR_OFF = STG_OFF + NB * STG_BYTES;
  static constexpr int NUM_BARS = 2 * NA + 2 * BSTAGES + 4 + 2 * NB;
  static constexpr int TOTAL = BAR_OFF + 8 * NUM_BARS + 16 + 1024;
};

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// p: tw = 8, th = 16, nb = 1, taps = 9, a_spatial = 1; tmA = halo map (box 64 x 10 x 18 x 1).
template <int BN, int BSTAGES, int NB, bool BRES, int NA_>
__global__ void __launch_bounds__(PersThreads<0>::THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const ConvPersParams p) {
  using L = ConvHaloSmem<BN, BSTAGES, NB, NA_>;
  constexpr int NA = L::NA;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bsm = smem + L::B_OFF;
  uint8_t* stg = smem + L::STG_OFF;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* a_empty = a_full + NA;
  uint64_t* b_full = a_empty + NA;
  uint64_t* b_empty = b_full + BSTAGES;
  uint64_t* acc_full = b_empty + BSTAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_full = acc_empty + 2;
  uint64_t* res_empty = res_full + NB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + NB);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int cin = p.cin_blocks * 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (p.has_res) tma_prefetch_desc(&tmR);
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], PersThreads<0>::EPI_WARPS); }
    for (int b = 0; b < NB; ++b) { mbar_init(&res_full[b], 1); mbar_init(&res_empty[b], 1); }
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 2 * BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may start its own prologue while this one runs / drains
  pdl_wait();                // everything below reads tensors written by the previous kernel of the stream

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ producer 1: weight tiles
      uint32_t gb = 0;
      bool first = true;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        if (!BRES || first) {
          for (int kc = 0; kc < p.cin_blocks; ++kc) {
            for (int tap = 0; tap < 9; ++tap, ++gb) {
              const int sb = BRES ? (kc * 9 + tap) : static_cast<int>(gb % BSTAGES);
              if (!BRES) mbar_wait(&b_empty[sb], ((gb / BSTAGES) & 1) ^ 1);
              mbar_expect_tx(&b_full[sb], L::B_BYTES);
              tma_load_2d(bsm + sb * L::B_BYTES, &tmB, &b_full[sb], tap * cin + kc * 64, c.n_tile * BN);
            }
          }
        }
        first = false;
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ------------------------------------------------------------ producer 2: input halo patches, running up to NA
      // patches ahead of the MMA warp independently of the weight stream (3x3 convolutions carry no residual, so
      // this warp has no residual tiles to prefetch)
      uint32_t ga = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        for (int kc = 0; kc < p.cin_blocks; ++kc, ++ga) {
          const int sa = ga % NA;
          mbar_wait(&a_empty[sa], ((ga / NA) & 1) ^ 1);
          mbar_expect_tx(&a_full[sa], L::HALO_DATA);
          tma_load_4d(smem + sa * L::HALO_SLOT, &tmA, &a_full[sa], kc * 64, c.wo0 - 1, c.ho0 - 1, c.n0);
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0 && p.has_res) {
      // ------------------------------------------------------------ residual producer (BasicBlock conv2: 3x3 + BN +
      // residual, resnet.py:27-41): the same protocol as warp 2 of conv_pers_kernel, run by the otherwise idle
      // TMEM-allocator warp
      constexpr int CHUNKS = BN / 64;
      uint32_t cc = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        for (int ch = 0; ch < CHUNKS; ++ch, ++cc) {
          const int b = cc % NB;
          mbar_wait(&res_empty[b], ((cc / NB) & 1) ^ 1);
          mbar_expect_tx(&res_full[b], L::STG_BYTES);
          tma_load_4d(stg + b * L::STG_BYTES, &tmR, &res_full[b], c.n_tile * BN + ch * 64, c.wo0, c.ho0, c.n0);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (whole warp walks the loop, one
    // elected lane issues; see conv_pers.cuh)
    constexpr uint32_t idesc = umma_idesc_f16(128, BN);
    uint32_t ga = 0, gb = 0, i = 0;
    bool first = true;
    const uint64_t bdesc0 = umma_desc_sw128(smem_u32(bsm));
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const uint32_t a = i & 1;
      mbar_wait(&acc_empty[a], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * BN;
      for (int kc = 0; kc < p.cin_blocks; ++kc, ++ga) {
        const int sa = ga % NA;
        mbar_wait(&a_full[sa], (ga / NA) & 1);
        tc_fence_after();
        // One thread issues every MMA, so its instruction count per MMA bounds small-N tiles (a 128x64x16 MMA
        // occupies the tensor core for only 32 cycles): build the base descriptors once and reach every (tap, k)
        // operand by adding a compile-time constant to the 14-bit start-address field.
        const uint64_t adesc0 = umma_desc_sw128_sbo(smem_u32(smem + sa * L::HALO_SLOT), L::HALO_W * 128u);
        const bool last_kc = (kc == p.cin_blocks - 1);
        if (BRES) {
          if (first) {
            for (int tap = 0; tap < 9; ++tap) mbar_wait(&b_full[kc * 9 + tap], 0);
            tc_fence_after();
          }
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kh = tap / 3, kw = tap - kh * 3;
              const uint64_t ad = adesc0 + static_cast<uint64_t>((kh * L::HALO_W + kw) * 8);      // * 128 B >> 4
              const uint64_t bd = bdesc0 + static_cast<uint64_t>(kc * 9 + tap) * (L::B_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kc | tap | k) != 0);
            }
            umma_commit(&a_empty[sa]);
            if (last_kc) umma_commit(&acc_full[a]);
          }
          __syncwarp();
        } else {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap, ++gb) {
            const int sb = static_cast<int>(gb % BSTAGES);
            mbar_wait(&b_full[sb], (gb / BSTAGES) & 1);
            tc_fence_after();
            const int kh = tap / 3, kw = tap - kh * 3;
            const uint64_t ad = adesc0 + static_cast<uint64_t>((kh * L::HALO_W + kw) * 8);
            const uint64_t bd = bdesc0 + static_cast<uint64_t>(sb) * (L::B_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kc | tap | k) != 0);
              umma_commit(&b_empty[sb]);
              if (tap == 8) {
                umma_commit(&a_empty[sa]);
                if (last_kc) umma_commit(&acc_full[a]);
              }
            }
            __syncwarp();
          }
        }
      }
      first = false;
    }
  } else if (warp >= 4) {
    constexpr int EPI_THREADS = 32 * PersThreads<0>::EPI_WARPS;
    const int quarter = warp & 3;
    const int hsel = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 128u + 128u * hsel);       // first thread of this epilogue group
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint32_t cc = 0, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const TileCoord c = decode_tile(p, t);
      const uint32_t a = i & 1;
      mbar_wait(&acc_full[a], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      conv_epilogue_tile<BN, NB, EPI_THREADS>(p, c, taddr, stg, res_full, res_empty, &acc_empty[a], cc, row_off, sw, hsel,
                                              lane, leader, tmO);
    }
    if (leader) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

template <int BN, int BSTAGES, int NB, bool BRES, int NA_>
int conv_halo_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                     const ConvPersParams& p, int num_sms, cudaStream_t stream) {
  using L = ConvHaloSmem<BN, BSTAGES, NB, NA_>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  auto kern = conv_halo_kernel<BN, BSTAGES, NB, BRES, NA_>;
  static std::atomic<uint64_t> attr_done{0};
  if (first_launch_on_device(attr_done))
    DIRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  DIRB_CUDA(launch_pdl(kern, dim3(grid), dim3(PersThreads<0>::THREADS), L::TOTAL, stream, tmA, tmB, tmR, tmO, p));
  count_launch();
  return 0;
}

}  // namespace dirb
