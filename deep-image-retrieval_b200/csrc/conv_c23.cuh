// conv2 (3x3 / s1 / p1) -> BN -> ReLU -> conv3 (1x1, x4) -> BN -> + residual -> ReLU of a Bottleneck as ONE persistent
// tcgen05 kernel (dirtorch/nets/backbones/resnet.py:75-85).  The conv2 output tile never leaves the SM:
//
//   conv2   128 pixels (8 x 16 patch) x CM channels, K = 9 * CM, halo patches as in conv_halo.cuh, accumulator A2 in TMEM
//   E1      A2 -> BN2 scale/shift -> ReLU -> fp16 -> shared memory, written directly in the 128-byte-swizzled K-major
//           layout a TMA load would produce: CM / 64 blocks of [128 pixels x 64 channels] = the A operand of conv3
//   conv3   for each 128-channel slice n3 of the 4 * CM outputs: K = CM from that buffer, accumulator A3[n3 & 1]
//   E2      A3 -> BN3 -> + residual (TMA-prefetched) -> ReLU -> fp16 -> TMA store       (conv_epilogue_tile, conv_pers.cuh)
//
// Per layer3 block of ResNet-101 this removes the write and the re-read of the conv2 output (2 x 134 MB of 1.48 GB at
// batch 64 x 1024^2) and, more importantly, turns the HBM-bound 1x1 expansion into the tail of a tensor-bound kernel:
// its 537 MB residual read and 537 MB output write stream while the tensor core is busy with the 3x3 of the next tile.
//
// TMEM (512 columns): A2 = [0, CM) as CM / 128 halves, A3[0] = [256, 384), A3[1] = [384, 512).
// 13 warps: 0 W2 producer, 1 MMA issuer, 2 halo producer, 3 TMEM allocator + residual producer, 4-11 epilogue (E1 and
// E2, two warps per TMEM lane quarter), 12 W3 producer.  The MMA warp software-pipelines the two GEMMs: the conv3 slices
// of tile i are issued between the (channel block, tap) groups of conv2 of tile i+1 as their accumulators free up, so
// E2 of tile i overlaps the tensor work of tile i+1; W2 and W3 stream through separate FIFO rings because the
// interleaving is decided at run time.
#pragma once
#include "conv_halo.cuh"

namespace dirb {

template <int CM>
struct ConvC23Smem {
  static constexpr int HALO_W = 10, HALO_H = 18;
  static constexpr int HALO_DATA = HALO_W * HALO_H * 128;          // 23 040 B landed by TMA
  static constexpr int HALO_SLOT = 24 * 1024;
  static constexpr int NA = 2;                                     // halo slots
  static constexpr int KB = CM / 64;                               // 64-channel blocks of the conv2 output
  static constexpr int T2_BYTES = KB * 128 * 128;                  // conv3's A operand
  static constexpr int B_BYTES = 128 * 128;                        // one weight tile: 128 output rows x 64 K
  static constexpr int NB2 = 3, NB3 = 2, NSTG = 2;
  static constexpr int STG_BYTES = 128 * 128;
  static constexpr int T2_OFF = NA * HALO_SLOT;
  static constexpr int B2_OFF = T2_OFF + T2_BYTES;
  static constexpr int B3_OFF = B2_OFF + NB2 * B_BYTES;
  static constexpr int STG_OFF = B3_OFF + NB3 * B_BYTES;
  static constexpr int BAR_OFF = STG_OFF + NSTG * STG_BYTES;
  static constexpr int NUM_BARS = 2 * NA + 2 * NB2 + 2 * NB3 + 4 + 4 + 2 * NSTG;
  static constexpr int TOTAL = BAR_OFF + 8 * NUM_BARS + 16 + 1024;
  static constexpr int N2 = CM < 128 ? CM : 128;                   // N of the conv2 MMAs (one W2 tile = N2 rows x 64 K)
  static constexpr int B2_TILE = N2 * 128;                         // bytes landed per W2 tile (slot stride stays B_BYTES)
  static constexpr int HALVES = CM / N2;                           // conv2 MMAs per K step
  static constexpr int NT3 = 4 * CM / 128;                         // 128-channel slices of the conv3 output
  static constexpr int THREADS = 13 * 32;
};

// p: a_spatial = 1, tw = 8, th = 16, nb = 1, n_tiles = 1, cin_blocks = CM / 64; scale/shift = BN3, scale2/shift2 = BN2.
// tmA: halo map over t1 (box 64 x 10 x 18 x 1); tmB2: W2 [CM][9*CM] (box 64 x min(CM, 128)); tmB3: W3 [4*CM][CM] (box 64 x 128);
// tmR / tmO: residual / output (B,H,W,4*CM), box 64 x 8 x 16 x 1.
template <int CM>
__global__ void __launch_bounds__(ConvC23Smem<CM>::THREADS, 1)
conv_c23_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2,
                const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmO, const ConvPersParams p) {
  using L = ConvC23Smem<CM>;
  constexpr int NA = L::NA, KB = L::KB, NB2 = L::NB2, NB3 = L::NB3, NSTG = L::NSTG, HALVES = L::HALVES, NT3 = L::NT3;
  constexpr int EPI_THREADS = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* t2 = smem + L::T2_OFF;
  uint8_t* b2sm = smem + L::B2_OFF;
  uint8_t* b3sm = smem + L::B3_OFF;
  uint8_t* stg = smem + L::STG_OFF;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* a_empty = a_full + NA;
  uint64_t* b2_full = a_empty + NA;
  uint64_t* b2_empty = b2_full + NB2;
  uint64_t* b3_full = b2_empty + NB2;
  uint64_t* b3_empty = b3_full + NB3;
  uint64_t* acc2_full = b3_empty + NB3;      // conv2 of a tile complete            (MMA commit -> E1)
  uint64_t* acc2_empty = acc2_full + 1;      // E1 has read A2 out of TMEM          (8 warps -> MMA)
  uint64_t* t2_full = acc2_empty + 1;        // E1 has written the conv3 operand    (1 thread -> MMA)
  uint64_t* t2_empty = t2_full + 1;          // conv3 of a tile has read it         (MMA commit -> E1)
  uint64_t* acc3_full = t2_empty + 1;        // [2]
  uint64_t* acc3_empty = acc3_full + 2;      // [2]
  uint64_t* res_full = acc3_empty + 2;       // [NSTG]
  uint64_t* res_empty = res_full + NSTG;     // [NSTG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + NSTG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB2);
    tma_prefetch_desc(&tmB3);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NB2; ++s) { mbar_init(&b2_full[s], 1); mbar_init(&b2_empty[s], 1); }
    for (int s = 0; s < NB3; ++s) { mbar_init(&b3_full[s], 1); mbar_init(&b3_empty[s], 1); }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, 8);
    mbar_init(t2_full, 1);
    mbar_init(t2_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&acc3_full[a], 1); mbar_init(&acc3_empty[a], 8); }
    for (int b = 0; b < NSTG; ++b) { mbar_init(&res_full[b], 1); mbar_init(&res_empty[b], 1); }
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int cin = CM;
  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ W2 tiles, in the order conv2 consumes them
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        for (int kc = 0; kc < KB; ++kc)
          for (int tap = 0; tap < 9; ++tap)
            for (int h = 0; h < HALVES; ++h, ++g) {
              const int s = g % NB2;
              mbar_wait(&b2_empty[s], ((g / NB2) & 1) ^ 1);
              mbar_expect_tx(&b2_full[s], L::B2_TILE);
              tma_load_2d(b2sm + s * L::B_BYTES, &tmB2, &b2_full[s], tap * cin + kc * 64, h * L::N2);
            }
      }
    }
  } else if (warp == 12) {
    if (lane == 0) {
      // ------------------------------------------------------------ W3 tiles, in the order conv3 consumes them
      uint32_t g = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        for (int n3 = 0; n3 < NT3; ++n3)
          for (int kb = 0; kb < KB; ++kb, ++g) {
            const int s = g % NB3;
            mbar_wait(&b3_empty[s], ((g / NB3) & 1) ^ 1);
            mbar_expect_tx(&b3_full[s], L::B_BYTES);
            tma_load_2d(b3sm + s * L::B_BYTES, &tmB3, &b3_full[s], kb * 64, n3 * 128);
          }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ------------------------------------------------------------ input halo patches of conv2
      uint32_t ga = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        for (int kc = 0; kc < KB; ++kc, ++ga) {
          const int sa = ga % NA;
          mbar_wait(&a_empty[sa], ((ga / NA) & 1) ^ 1);
          mbar_expect_tx(&a_full[sa], L::HALO_DATA);
          tma_load_4d(smem + sa * L::HALO_SLOT, &tmA, &a_full[sa], kc * 64, c.wo0 - 1, c.ho0 - 1, c.n0);
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------------------------------------ residual tiles, in the order E2 consumes them
      // Only NSTG = 2 staging buffers exist (the conv3 operand occupies the shared memory a deeper residual ring would
      // need), so a residual chunk cannot be requested more than ~one chunk ahead of its use and its HBM latency would
      // sit on the epilogue's critical path (measured: 2.5 us per 16 KB chunk, 40 us per tile).  The residual of the
      // NEXT tile is therefore prefetched into L2 while this tile is processed: 256 KB per CTA, 38 MB for the grid.
      uint32_t cc = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        if (t == static_cast<int>(blockIdx.x)) {
          for (int ch = 0; ch < 4 * CM / 64; ++ch) tma_prefetch_4d(&tmR, ch * 64, c.wo0, c.ho0, c.n0);
        }
        if (t + static_cast<int>(gridDim.x) < p.total_tiles) {
          const TileCoord cn = decode_tile(p, t + gridDim.x);
          for (int ch = 0; ch < 4 * CM / 64; ++ch) tma_prefetch_4d(&tmR, ch * 64, cn.wo0, cn.ho0, cn.n0);
        }
        for (int ch = 0; ch < 4 * CM / 64; ++ch, ++cc) {
          const int b = cc % NSTG;
          mbar_wait(&res_empty[b], ((cc / NSTG) & 1) ^ 1);
          mbar_expect_tx(&res_full[b], L::STG_BYTES);
          tma_load_4d(stg + b * L::STG_BYTES, &tmR, &res_full[b], ch * 64, c.wo0, c.ho0, c.n0);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer (whole warp walks the loops, one elected
    // lane issues; see conv_pers.cuh)
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    constexpr uint32_t idesc2 = umma_idesc_f16(128, L::N2);
    const uint64_t b2desc0 = umma_desc_sw128(smem_u32(b2sm));
    const uint64_t b3desc0 = umma_desc_sw128(smem_u32(b3sm));
    const uint64_t t2desc0 = umma_desc_sw128(smem_u32(t2));
    const uint32_t acc2 = tmem_base;
    uint32_t ga = 0, g2 = 0, g3 = 0, m3 = 0, i = 0;
    int pend = 0;             // conv3 slices of the PREVIOUS tile still to issue (they read t2 of tile i-1)
    bool t2_ready = false;    // t2_full of the previous tile has been observed
    uint32_t pend_i = 0;      // local index of the tile the pending slices belong to

    // Issue one conv3 slice (n3 = NT3 - pend of tile pend_i) if `block` or its accumulator + operand are ready.
    // (try_wait results may differ between the lanes of the warp for an instant: the decision is made warp-uniform,
    // elect_one() below needs the full warp)
    auto ready = [&](uint64_t* bar, uint32_t par, bool block) -> bool {
      if (block) {
        mbar_wait(bar, par);
        return true;
      }
      return __all_sync(0xffffffffu, mbar_try_wait(bar, par) != 0) != 0;
    };
    auto issue_conv3 = [&](bool block) -> bool {
      if (!t2_ready) {
        if (!ready(t2_full, pend_i & 1, block)) return false;
        t2_ready = true;
        tc_fence_after();
      }
      const uint32_t a = m3 & 1;
      if (!ready(&acc3_empty[a], ((m3 >> 1) & 1) ^ 1, block)) return false;
      tc_fence_after();
      const uint32_t d3 = tmem_base + 256 + a * 128;
      for (int kb = 0; kb < KB; ++kb, ++g3) {
        const int s = g3 % NB3;
        mbar_wait(&b3_full[s], (g3 / NB3) & 1);
        tc_fence_after();
        const uint64_t ad = t2desc0 + static_cast<uint64_t>(kb) * (16384 >> 4);
        const uint64_t bd = b3desc0 + static_cast<uint64_t>(s) * (L::B_BYTES >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d3, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&b3_empty[s]);
          if (kb == KB - 1) {
            umma_commit(&acc3_full[a]);
            if (pend == 1) umma_commit(t2_empty);          // last slice of the tile: its operand buffer may be rewritten
          }
        }
        __syncwarp();
      }
      ++m3;
      --pend;
      return true;
    };

    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      // ---- conv2 of tile i (its accumulator must have been drained by E1 of tile i-1)
      mbar_wait(acc2_empty, (i & 1) ^ 1);
      tc_fence_after();
      for (int kc = 0; kc < KB; ++kc, ++ga) {
        const int sa = ga % NA;
        mbar_wait(&a_full[sa], (ga / NA) & 1);
        tc_fence_after();
        const uint64_t adesc0 = umma_desc_sw128_sbo(smem_u32(smem + sa * L::HALO_SLOT), L::HALO_W * 128u);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int kh = tap / 3, kw = tap - kh * 3;
          const uint64_t ad = adesc0 + static_cast<uint64_t>((kh * L::HALO_W + kw) * 8);
#pragma unroll
          for (int h = 0; h < HALVES; ++h, ++g2) {
            const int s = static_cast<int>(g2 % NB2);
            mbar_wait(&b2_full[s], (g2 / NB2) & 1);
            tc_fence_after();
            const uint64_t bd = b2desc0 + static_cast<uint64_t>(s) * (L::B_BYTES >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(acc2 + h * L::N2, ad + 2 * k, bd + 2 * k, idesc2, (kc | tap | k) != 0);
              umma_commit(&b2_empty[s]);
              if (tap == 8 && h == HALVES - 1) {
                umma_commit(&a_empty[sa]);
                if (kc == KB - 1) umma_commit(acc2_full);
              }
            }
            __syncwarp();
          }
          if (pend > 0) issue_conv3(false);     // slot a conv3 slice of the previous tile in between
        }
      }
      // ---- what is left of conv3 of tile i-1, then tile i's slices become the pending ones
      while (pend > 0) issue_conv3(true);
      pend = NT3;
      pend_i = i;
      t2_ready = false;
    }
    while (pend > 0) issue_conv3(true);
  } else if (warp >= 4 && warp < 12) {
    // -------------------------------------------------------------- epilogue warps: E1 then the NT3 slices of E2, per tile
    const int quarter = warp & 3;
    const int hsel = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 128u + 128u * hsel);       // first thread of this epilogue group (E2)
    const bool leader0 = (threadIdx.x == 128);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t cc = 0, m3 = 0, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      TileCoord c = decode_tile(p, t);
      // ---- E1: conv2 accumulator -> BN2 + ReLU -> fp16 -> conv3's A operand in shared memory
      mbar_wait(acc2_full, i & 1);
      mbar_wait(t2_empty, (i & 1) ^ 1);                      // conv3 of the previous tile has finished reading the buffer
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        float v[32];
        tmem_ld32(tmem_base + lane_off + kb * 64 + hsel * 32, v);
        tmem_ld_wait();
        if (kb == KB - 1) {                                  // last TMEM read of A2: conv2 of the next tile may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc2_empty);
        }
        const float4* sc4 = reinterpret_cast<const float4*>(p.scale2 + kb * 64 + hsel * 32);
        const float4* sh4 = reinterpret_cast<const float4*>(p.shift2 + kb * 64 + hsel * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 s = __ldg(sc4 + q), h = __ldg(sh4 + q);
          v[4 * q + 0] = fmaxf(fmaf(v[4 * q + 0], s.x, h.x), 0.0f);
          v[4 * q + 1] = fmaxf(fmaf(v[4 * q + 1], s.y, h.y), 0.0f);
          v[4 * q + 2] = fmaxf(fmaf(v[4 * q + 2], s.z, h.z), 0.0f);
          v[4 * q + 3] = fmaxf(fmaf(v[4 * q + 3], s.w, h.w), 0.0f);
        }
        uint8_t* buf = t2 + kb * 16384;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t chunk = static_cast<uint32_t>(hsel * 4 + j);
          uint4 o;
          o.x = pack_h2(v[j * 8 + 0], v[j * 8 + 1]);
          o.y = pack_h2(v[j * 8 + 2], v[j * 8 + 3]);
          o.z = pack_h2(v[j * 8 + 4], v[j * 8 + 5]);
          o.w = pack_h2(v[j * 8 + 6], v[j * 8 + 7]);
          *reinterpret_cast<uint4*>(buf + row_off + ((chunk ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async_smem();                              // generic-proxy writes -> visible to the tensor core
      named_bar_sync(5, EPI_THREADS);                        // (ids 1-4 belong to the two groups of conv_epilogue_tile)
      if (leader0) mbar_arrive(t2_full);
      // ---- E2: the NT3 output slices of conv3
#pragma unroll 1
      for (int n3 = 0; n3 < NT3; ++n3, ++m3) {
        const uint32_t a = m3 & 1;
        mbar_wait(&acc3_full[a], (m3 >> 1) & 1);
        tc_fence_after();
        c.n_tile = n3;
        conv_epilogue_tile<128, NSTG, EPI_THREADS>(p, c, tmem_base + 256 + a * 128 + lane_off, stg, res_full, res_empty,
                                                   &acc3_empty[a], cc, row_off, sw, hsel, lane, leader, tmO);
      }
    }
    if (leader) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int CM>
int conv_c23_launch(const CUtensorMap& tmA, const CUtensorMap& tmB2, const CUtensorMap& tmB3, const CUtensorMap& tmR,
                    const CUtensorMap& tmO, const ConvPersParams& p, int num_sms, cudaStream_t stream) {
  using L = ConvC23Smem<CM>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static_assert(CM == 64 || CM == 128 || CM == 256, "TMEM budget: conv2 accumulator of CM <= 256 columns");
  auto kern = conv_c23_kernel<CM>;
  static std::atomic<uint64_t> attr_done{0};
  if (first_launch_on_device(attr_done))
    DIRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  DIRB_CUDA(launch_pdl(kern, dim3(grid), dim3(L::THREADS), L::TOTAL, stream, tmA, tmB2, tmB3, tmR, tmO, p));
  count_launch();
  return 0;
}

}  // namespace dirb
