// conv_c23.cuh on CTA PAIRS: the same fused conv2 (3x3) -> BN -> ReLU -> conv3 (1x1, x4) -> BN -> + residual -> ReLU,
// with every MMA spanning two SMs (tcgen05 cta_group::2, thread-block cluster of 2).
//
// Why: the single-CTA kernel is starved by its weight rings.  Every 128-pixel tile streams ALL of W2 and W3 (1.66 MB for
// 256 mid channels) and the conv2 output tile (64 KB) has to stay resident, which leaves 48 KB of shared memory for the
// W2 ring - at the ~1.4 us the L2 needs under that load, a third of what keeps the tensor core fed (measured: 0.69 ms per
// layer3 block instead of 0.42 ms for the two separate kernels).  With cta_group::2 one MMA covers the pixel tiles of BOTH
// CTAs (M = 256) and reads the weight tile half from each CTA's shared memory: each CTA loads and holds only HALF of
// every weight tile, so the same ring bytes hold twice the K depth and the L2 -> SM weight traffic halves.
//
// Division of labour inside a pair (cluster ranks 0 = leader, 1 = peer):
//   both CTAs   halo producer (own pixel tile), W2 / W3 producers (own half of each tile), residual producer, epilogue
//               warps (E1: conv2 accumulator -> BN2/ReLU -> fp16 conv3 operand in OWN shared memory; E2: conv3
//               accumulator -> BN3 + residual + ReLU -> TMA store), TMEM allocation (cta_group::2, same columns)
//   leader only the MMA warp.  It waits on the LEADER's barriers: operand "full" barriers count the TMA bytes of both
//               CTAs (the peer's loads signal the leader's barrier), "accumulator drained" / "operand written" barriers
//               collect remote arrivals from the peer's epilogue warps.  Its tcgen05.commit multicasts to the barrier
//               at the same offset in both CTAs, which is how the peer's producers and epilogue warps see progress.
// Tiles: pair j of P processes tiles 2j + rank, 2(j + P) + rank, ...; an odd tile count makes the last CTA repeat the
// final tile (identical values written twice).
#pragma once
#include "conv_c23.cuh"

namespace dirb {

template <int CM>
struct ConvC23PSmem {
  static constexpr int HALO_W = 10, HALO_H = 18;
  static constexpr int HALO_DATA = HALO_W * HALO_H * 128;
  static constexpr int HALO_SLOT = 24 * 1024;
  static constexpr int NA = 2;
  static constexpr int KB = CM / 64;
  static constexpr int T2_BYTES = KB * 128 * 128;
  static constexpr int B2_ROWS = CM / 2;                           // rows of a W2 tile this CTA holds (N = CM per MMA)
  static constexpr int B2_TILE = B2_ROWS * 128;
  static constexpr int B2_SLOT = B2_TILE < 1024 ? 1024 : B2_TILE;  // slots stay 1024-byte aligned
  static constexpr int B3_ROWS = 64;                               // N = 128 per conv3 slice
  static constexpr int B3_TILE = B3_ROWS * 128;
  static constexpr int NB2 = (CM == 256) ? 4 : 6, NB3 = (CM == 256) ? 2 : 4, NSTG = 2;
  static constexpr int STG_BYTES = 128 * 128;
  static constexpr int T2_OFF = NA * HALO_SLOT;
  static constexpr int B2_OFF = T2_OFF + T2_BYTES;
  static constexpr int B3_OFF = B2_OFF + NB2 * B2_SLOT;
  static constexpr int STG_OFF = B3_OFF + NB3 * B3_TILE;
  static constexpr int BAR_OFF = STG_OFF + NSTG * STG_BYTES;
  static constexpr int NUM_BARS = 2 * NA + 2 * NB2 + 2 * NB3 + 4 + 4 + 2 * NSTG;
  static constexpr int TOTAL = BAR_OFF + 8 * NUM_BARS + 16 + 1024;
  static constexpr int NT3 = 4 * CM / 128;
  static constexpr int THREADS = 13 * 32;
};

// conv_epilogue_tile of conv_pers.cuh with the "accumulator drained" arrival sent to the leader CTA's barrier.
template <int BN, int NBUF, int EPI_THREADS>
__device__ __forceinline__ void conv_epilogue_tile_pair(const ConvPersParams& p, const TileCoord& c, uint32_t taddr, uint8_t* stg,
                                                        uint64_t* res_full, uint64_t* res_empty, uint32_t acc_empty_leader,
                                                        uint32_t& cc, uint32_t row_off, uint32_t sw, int hsel, int lane,
                                                        bool leader, const CUtensorMap& tmO) {
  constexpr int CHUNKS = BN / 64;
  constexpr int STG_BYTES = 128 * 128;
#pragma unroll 1
  for (int ch = 0; ch < CHUNKS; ++ch, ++cc) {
    const int b = cc % NBUF;
    uint8_t* buf = stg + b * STG_BYTES;
    mbar_wait(&res_full[b], (cc / NBUF) & 1);          // residual chunk has landed in `buf`
    const int col0 = c.n_tile * BN + ch * 64;
    float v[32];
    tmem_ld32(taddr + ch * 64 + hsel * 32, v);
    tmem_ld_wait();
    if (ch == CHUNKS - 1) {                            // last TMEM read of this slice: release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty_leader);
    }
    const float4* sc4 = reinterpret_cast<const float4*>(p.scale + col0 + hsel * 32);
    const float4* sh4 = reinterpret_cast<const float4*>(p.shift + col0 + hsel * 32);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 s = __ldg(sc4 + q), h = __ldg(sh4 + q);
      v[4 * q + 0] = fmaf(v[4 * q + 0], s.x, h.x);
      v[4 * q + 1] = fmaf(v[4 * q + 1], s.y, h.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], s.z, h.z);
      v[4 * q + 3] = fmaf(v[4 * q + 3], s.w, h.w);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t chunk = static_cast<uint32_t>(hsel * 4 + j);
      uint4* sp = reinterpret_cast<uint4*>(buf + row_off + ((chunk ^ sw) << 4));
      const uint4 r = *sp;
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_h2(rr[e]);
        v[j * 8 + e * 2] = fmaxf(v[j * 8 + e * 2] + f.x, 0.0f);
        v[j * 8 + e * 2 + 1] = fmaxf(v[j * 8 + e * 2 + 1] + f.y, 0.0f);
      }
      uint4 o;
      o.x = pack_h2(v[j * 8 + 0], v[j * 8 + 1]);
      o.y = pack_h2(v[j * 8 + 2], v[j * 8 + 3]);
      o.z = pack_h2(v[j * 8 + 4], v[j * 8 + 5]);
      o.w = pack_h2(v[j * 8 + 6], v[j * 8 + 7]);
      *sp = o;
    }
    fence_proxy_async_smem();
    named_bar_sync(2, EPI_THREADS);
    if (leader) {
      tma_store_4d(&tmO, buf, col0, c.wo0, c.ho0, c.n0);
      bulk_commit();
      if (cc >= 1) {
        bulk_wait_read<1>();
        mbar_arrive(&res_empty[(cc - 1) % NBUF]);
      }
    }
  }
}

template <int CM>
__global__ void __launch_bounds__(ConvC23PSmem<CM>::THREADS, 1)
conv_c23p_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ CUtensorMap tmR,
                 const __grid_constant__ CUtensorMap tmO, const ConvPersParams p) {
  using L = ConvC23PSmem<CM>;
  constexpr int NA = L::NA, KB = L::KB, NB2 = L::NB2, NB3 = L::NB3, NSTG = L::NSTG, NT3 = L::NT3;
  constexpr int EPI_THREADS = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* t2 = smem + L::T2_OFF;
  uint8_t* b2sm = smem + L::B2_OFF;
  uint8_t* b3sm = smem + L::B3_OFF;
  uint8_t* stg = smem + L::STG_OFF;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);     // leader: bytes of both CTAs
  uint64_t* a_empty = a_full + NA;                                       // each CTA: multicast commit
  uint64_t* b2_full = a_empty + NA;
  uint64_t* b2_empty = b2_full + NB2;
  uint64_t* b3_full = b2_empty + NB2;
  uint64_t* b3_empty = b3_full + NB3;
  uint64_t* acc2_full = b3_empty + NB3;      // each CTA: multicast commit
  uint64_t* acc2_empty = acc2_full + 1;      // leader: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint64_t* t2_full = acc2_empty + 1;        // leader: 2 arrivals
  uint64_t* t2_empty = t2_full + 1;          // each CTA: multicast commit
  uint64_t* acc3_full = t2_empty + 1;        // [2] each CTA: multicast commit
  uint64_t* acc3_empty = acc3_full + 2;      // [2] leader: 16 arrivals
  uint64_t* res_full = acc3_empty + 2;       // [NSTG] local
  uint64_t* res_empty = res_full + NSTG;     // [NSTG] local
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + NSTG);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool lead_cta = (rank == 0);

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
    mbar_init(acc2_empty, 16);
    mbar_init(t2_full, 2);
    mbar_init(t2_empty, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&acc3_full[a], 1); mbar_init(&acc3_empty[a], 16); }
    for (int b = 0; b < NSTG; ++b) { mbar_init(&res_full[b], 1); mbar_init(&res_empty[b], 1); }
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barrier inits of both CTAs are visible before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int pairs = gridDim.x >> 1;
  const int pair_id = blockIdx.x >> 1;
  const int pair_tiles = (p.total_tiles + 1) >> 1;
  auto my_tile = [&](int j) {
    const int t = 2 * j + static_cast<int>(rank);
    return t < p.total_tiles ? t : p.total_tiles - 1;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ this CTA's half of the W2 tiles
      const uint32_t full0 = mapa_shared(b2_full, 0);
      uint32_t g = 0;
      for (int j = pair_id; j < pair_tiles; j += pairs) {
        for (int kc = 0; kc < KB; ++kc)
          for (int tap = 0; tap < 9; ++tap, ++g) {
            const int s = g % NB2;
            mbar_wait(&b2_empty[s], ((g / NB2) & 1) ^ 1);
            if (lead_cta) mbar_expect_tx(&b2_full[s], 2 * L::B2_TILE);
            tma_load_2d_pair(b2sm + s * L::B2_SLOT, &tmB2, full0 + s * 8, tap * CM + kc * 64, static_cast<int>(rank) * L::B2_ROWS);
          }
      }
    }
  } else if (warp == 12) {
    if (lane == 0) {
      // ------------------------------------------------------------ this CTA's half of the W3 tiles
      const uint32_t full0 = mapa_shared(b3_full, 0);
      uint32_t g = 0;
      for (int j = pair_id; j < pair_tiles; j += pairs) {
        for (int n3 = 0; n3 < NT3; ++n3)
          for (int kb = 0; kb < KB; ++kb, ++g) {
            const int s = g % NB3;
            mbar_wait(&b3_empty[s], ((g / NB3) & 1) ^ 1);
            if (lead_cta) mbar_expect_tx(&b3_full[s], 2 * L::B3_TILE);
            tma_load_2d_pair(b3sm + s * L::B3_TILE, &tmB3, full0 + s * 8, kb * 64, n3 * 128 + static_cast<int>(rank) * L::B3_ROWS);
          }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ------------------------------------------------------------ halo patches of this CTA's pixel tile
      const uint32_t full0 = mapa_shared(a_full, 0);
      uint32_t ga = 0;
      for (int j = pair_id; j < pair_tiles; j += pairs) {
        const TileCoord c = decode_tile(p, my_tile(j));
        for (int kc = 0; kc < KB; ++kc, ++ga) {
          const int sa = ga % NA;
          mbar_wait(&a_empty[sa], ((ga / NA) & 1) ^ 1);
          if (lead_cta) mbar_expect_tx(&a_full[sa], 2 * L::HALO_DATA);
          tma_load_4d_pair(smem + sa * L::HALO_SLOT, &tmA, full0 + sa * 8, kc * 64, c.wo0 - 1, c.ho0 - 1, c.n0);
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ------------------------------------------------------------ residual tiles (local barriers)
      // (the residual of the NEXT tile is prefetched into L2 while this one is processed, see conv_c23.cuh)
      uint32_t cc = 0;
      for (int j = pair_id; j < pair_tiles; j += pairs) {
        const TileCoord c = decode_tile(p, my_tile(j));
        if (j == pair_id) {
          for (int ch = 0; ch < 4 * CM / 64; ++ch) tma_prefetch_4d(&tmR, ch * 64, c.wo0, c.ho0, c.n0);
        }
        if (j + pairs < pair_tiles) {
          const TileCoord cn = decode_tile(p, my_tile(j + pairs));
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
    if (lead_cta) {
      // ------------------------------------------------------------ MMA issuer of the pair
      constexpr uint32_t idesc2 = umma_idesc_f16(256, CM);
      constexpr uint32_t idesc3 = umma_idesc_f16(256, 128);
      const uint64_t b2desc0 = umma_desc_sw128(smem_u32(b2sm));
      const uint64_t b3desc0 = umma_desc_sw128(smem_u32(b3sm));
      const uint64_t t2desc0 = umma_desc_sw128(smem_u32(t2));
      const uint32_t acc2 = tmem_base;
      uint32_t ga = 0, g2 = 0, g3 = 0, m3 = 0, i = 0;
      int pend = 0;
      bool t2_ready = false;
      uint32_t pend_i = 0;
      auto ready = [&](uint64_t* bar, uint32_t par, bool block) -> bool {
        if (block) {
          mbar_wait_cluster(bar, par);
          return true;
        }
        return __all_sync(0xffffffffu, mbar_try_wait_cluster(bar, par) != 0) != 0;
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
          mbar_wait_cluster(&b3_full[s], (g3 / NB3) & 1);
          tc_fence_after();
          const uint64_t ad = t2desc0 + static_cast<uint64_t>(kb) * (16384 >> 4);
          const uint64_t bd = b3desc0 + static_cast<uint64_t>(s) * (L::B3_TILE >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_pair(d3, ad + 2 * k, bd + 2 * k, idesc3, (kb | k) != 0);
            umma_commit_pair(&b3_empty[s]);
            if (kb == KB - 1) {
              umma_commit_pair(&acc3_full[a]);
              if (pend == 1) umma_commit_pair(t2_empty);
            }
          }
          __syncwarp();
        }
        ++m3;
        --pend;
        return true;
      };

      for (int j = pair_id; j < pair_tiles; j += pairs, ++i) {
        mbar_wait_cluster(acc2_empty, (i & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < KB; ++kc, ++ga) {
          const int sa = ga % NA;
          mbar_wait_cluster(&a_full[sa], (ga / NA) & 1);
          tc_fence_after();
          const uint64_t adesc0 = umma_desc_sw128_sbo(smem_u32(smem + sa * L::HALO_SLOT), L::HALO_W * 128u);
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap, ++g2) {
            const int kh = tap / 3, kw = tap - kh * 3;
            const uint64_t ad = adesc0 + static_cast<uint64_t>((kh * L::HALO_W + kw) * 8);
            const int s = static_cast<int>(g2 % NB2);
            mbar_wait_cluster(&b2_full[s], (g2 / NB2) & 1);
            tc_fence_after();
            const uint64_t bd = b2desc0 + static_cast<uint64_t>(s) * (L::B2_SLOT >> 4);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_pair(acc2, ad + 2 * k, bd + 2 * k, idesc2, (kc | tap | k) != 0);
              umma_commit_pair(&b2_empty[s]);
              if (tap == 8) {
                umma_commit_pair(&a_empty[sa]);
                if (kc == KB - 1) umma_commit_pair(acc2_full);
              }
            }
            __syncwarp();
            if (pend > 0) issue_conv3(false);
          }
        }
        while (pend > 0) issue_conv3(true);
        pend = NT3;
        pend_i = i;
        t2_ready = false;
      }
      while (pend > 0) issue_conv3(true);
    }
  } else if (warp >= 4 && warp < 12) {
    // -------------------------------------------------------------- epilogue warps of this CTA's pixel tile
    const int quarter = warp & 3;
    const int hsel = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;
    const bool leader = (threadIdx.x == 128);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t acc2_empty0 = mapa_shared(acc2_empty, 0);
    const uint32_t t2_full0 = mapa_shared(t2_full, 0);
    const uint32_t acc3_empty0 = mapa_shared(acc3_empty, 0);
    uint32_t cc = 0, m3 = 0, i = 0;
    for (int j = pair_id; j < pair_tiles; j += pairs, ++i) {
      TileCoord c = decode_tile(p, my_tile(j));
      mbar_wait(acc2_full, i & 1);
      mbar_wait(t2_empty, (i & 1) ^ 1);
      tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        float v[32];
        tmem_ld32(tmem_base + lane_off + kb * 64 + hsel * 32, v);
        tmem_ld_wait();
        if (kb == KB - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(acc2_empty0);
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
        for (int jj = 0; jj < 4; ++jj) {
          const uint32_t chunk = static_cast<uint32_t>(hsel * 4 + jj);
          uint4 o;
          o.x = pack_h2(v[jj * 8 + 0], v[jj * 8 + 1]);
          o.y = pack_h2(v[jj * 8 + 2], v[jj * 8 + 3]);
          o.z = pack_h2(v[jj * 8 + 4], v[jj * 8 + 5]);
          o.w = pack_h2(v[jj * 8 + 6], v[jj * 8 + 7]);
          *reinterpret_cast<uint4*>(buf + row_off + ((chunk ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(3, EPI_THREADS);
      if (leader) mbar_arrive_cluster(t2_full0);
#pragma unroll 1
      for (int n3 = 0; n3 < NT3; ++n3, ++m3) {
        const uint32_t a = m3 & 1;
        mbar_wait(&acc3_full[a], (m3 >> 1) & 1);
        tc_fence_after();
        c.n_tile = n3;
        conv_epilogue_tile_pair<128, NSTG, EPI_THREADS>(p, c, tmem_base + 256 + a * 128 + lane_off, stg, res_full, res_empty,
                                                        acc3_empty0 + a * 8, cc, row_off, sw, hsel, lane, leader, tmO);
      }
    }
    if (leader) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be signalling this CTA's barriers / reading its shared memory
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int CM>
int conv_c23p_launch(const CUtensorMap& tmA, const CUtensorMap& tmB2, const CUtensorMap& tmB3, const CUtensorMap& tmR,
                     const CUtensorMap& tmO, const ConvPersParams& p, int num_sms, cudaStream_t stream) {
  using L = ConvC23PSmem<CM>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static_assert(CM == 64 || CM == 128 || CM == 256, "TMEM budget: conv2 accumulator of CM <= 256 columns");
  auto kern = conv_c23p_kernel<CM>;
  static std::atomic<uint64_t> attr_done{0};
  if (first_launch_on_device(attr_done))
    DIRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int pair_tiles = (p.total_tiles + 1) / 2;
  int grid = 2 * (pair_tiles < num_sms / 2 ? pair_tiles : num_sms / 2);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(L::THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  DIRB_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB2, tmB3, tmR, tmO, p));
  count_launch();
  return 0;
}

}  // namespace dirb
