// Persistent tcgen05 implicit-GEMM convolution + BN + residual + ReLU for sm_100a.
//
// One CTA per SM loops over output tiles (128 pixels x BN channels, BN up to 256); 12 warps (4 + 8 epilogue warps;
// the similarity / whitening epilogues use 4), each with one job:
//
//   warp 0  TMA producer   A (128x64 activation patch) and B (BNx64 weights) into a STAGES-deep smem ring
//   warp 1  MMA issuer     tcgen05.mma (M=128, N=BN, K=16) x 4 per ring slot into one of TWO TMEM accumulators,
//                          so the tensor core starts tile i+1 while tile i is still being drained
//   warp 2  residual TMA   prefetches the residual tile (same box as the output tile) 64 channels at a time into
//                          the staging buffers, up to NBUF chunks ahead of the epilogue
//   warp 3  TMEM allocator
//   warps 4-11 epilogue    (two warps per TMEM lane quarter, each taking 32 of the 64 channels of a chunk)
//                          tcgen05.ld (one accumulator row per thread) -> scale/shift (+ residual read from the
//                          staging buffer) (+ ReLU) -> fp16, written back IN PLACE into the 128-byte-swizzled
//                          staging buffer -> one TMA store per 64-channel chunk (clips ragged edges by itself)
//
// All global traffic is bulk/asynchronous (TMA); no thread ever issues a strided global load or store.  The A
// operand addressing (flat rows or NHWC patches with taps, padding and stride through the tensor map) is the same
// as in gemm_tc.cuh.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace dirb {

enum { PERS_EPI_CONV = 0, PERS_EPI_SIM_DENSE = 1, PERS_EPI_SIM_FILTER = 2, PERS_EPI_SIM_GMAX = 3, PERS_EPI_F32 = 4 };

struct ConvPersParams {
  int a_spatial, taps, kw_taps, cin_blocks, stride, pad;
  int tw, th, nb, tiles_w, tiles_h;
  int n_tiles, total_tiles;
  int m_fastest, m_tiles;      // tile order: 0 = n-tile fastest (convolutions: CTAs share the activation tile),
                               // 1 = m-tile fastest (search: CTAs running together share the streamed database tile)
  int has_res, relu;
  // fused projection shortcut (block 0 of a layer): k-iterations >= k_split read their A tile from a SECOND
  // activation tensor (tensor map tmR, spatial boxes with element stride a2_stride baked into the map) - the
  // block input x of the 1x1 downsample conv - while the weights are the K-concatenation [W3*s3 | Wd*sd].
  int k_split, a2_stride;
  // PERS_EPI_F32 (whitening): fp32-accurate product from fp16 hi/lo splits, K = 3 parts of k_per_part blocks:
  // (A_hi, B_lo), (A_lo, B_hi), (A_hi, B_hi) with A_lo in tensor map tmR and B_lo in tmO; epilogue = column scale,
  // fp32 store to dense[M][dense_ld].
  int k_per_part;
  const float* scale;
  const float* shift;
  const float* scale2;         // conv_c23.cuh: BN of the first (3x3) convolution of the fused pair
  const float* shift2;
  // similarity epilogues (search.cu): D[q][n] = <query q, database row n>
  int M, N;                    // valid queries / database rows of this launch
  float* dense;                // [M][dense_ld] fp32 scores                     (PERS_EPI_SIM_DENSE)
  int64_t dense_ld;
  const float* thr;            // [M] per-query threshold                       (PERS_EPI_SIM_FILTER)
  unsigned long long* cand;    // [M][cand_cap] packed (score bits << 32 | row)
  int* cand_cnt;               // [M]
  int cand_cap;
  // HBM-bound 1x1 convolutions: the producers also request the NEXT tile's activation / residual boxes into L2
  // (cp.async.bulk.prefetch.tensor) while the current tile streams, which deepens the HBM request queue beyond what the
  // shared-memory rings can hold in flight.  0 = off.
  int l2_prefetch;
  // Epilogue organisation of the convolution kernels (A/B knobs, option "epi_mode"): bit 0 = 1: two independent groups of
  // 4 warps (two chunks in flight), 0: all 8 warps on one chunk; bit 1 = 1: a staging buffer goes back to the residual
  // producer as soon as its own TMA store has finished reading it (earliest possible), 0: one store later.
  // Measured on ResNet-101 64 x 1024^2 (profiles/r2_conv_sweep.txt): 0 is the fastest; 1 starves the residual ring (two
  // buffers per group), 2 puts the store's read-completion wait on the leader's critical path, 3 equals 0.  A third
  // organisation (every warp storing its own 32-pixel slab, no CTA-wide barrier per chunk) also measured equal to 0 and
  // was removed again: once the long-latency links are out of the chain (see conv_epilogue_tile_1g) the 1x1 convolutions
  // with a residual run at ~4.9 TB/s in situ whatever the organisation.
  int epi_mode;
  // Device-side launch predicate (retry passes of the search): when non-null the whole grid returns at once unless
  // *gate != 0.  The value was written by the previous kernel of the stream, so it is read after pdl_wait().
  const int* gate;
};

template <int BN, int STAGES, int EPI = 0, int NB = 4>
struct ConvPersSmem {
  static constexpr int NBUF = (EPI == 0) ? NB : 0;  // staging buffers exist only for the convolution epilogue
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_BYTES = 128 * 128;                      // 128 rows x 64 channels fp16
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  // BN scale | shift of the current tile, double-buffered by tile parity: only where the epilogue is the bottleneck
  // (configurations with a residual ring) and the budget allows it
  static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * (NB > 4 ? NB : 4);   // ring, accumulators, residual ring
  static constexpr bool SS = (EPI == 0) && (NB >= 4) &&
                             (STAGES * STAGE_BYTES + NB * STG_BYTES + 16 * BN + 8 * NUM_BARS + 16 + 1024 <= 232448);
  static constexpr int SS_OFF = STG_OFF + NBUF * STG_BYTES;
  static constexpr int SS_BYTES = SS ? 2 * 2 * BN * 4 : 0;
  static constexpr int BAR_OFF = SS_OFF + SS_BYTES;
  static constexpr int TOTAL = BAR_OFF + 8 * NUM_BARS + 16 + 1024;  // + tmem slot + alignment slack
  static constexpr int TMEM_COLS = 2 * BN;                          // two accumulators
};

struct TileCoord {
  int m_tile, n_tile, wo0, ho0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvPersParams& p, int t) {
  TileCoord c;
  if (p.m_fastest) {
    c.m_tile = t % p.m_tiles;
    c.n_tile = t / p.m_tiles;
  } else {
    c.n_tile = t % p.n_tiles;
    c.m_tile = t / p.n_tiles;
  }
  c.wo0 = c.ho0 = c.n0 = 0;
  if (p.a_spatial) {
    const int tx = c.m_tile % p.tiles_w;
    const int ty = (c.m_tile / p.tiles_w) % p.tiles_h;
    const int tb = c.m_tile / (p.tiles_w * p.tiles_h);
    c.wo0 = tx * p.tw;
    c.ho0 = ty * p.th;
    c.n0 = tb * p.nb;
  }
  return c;
}

// Convolution epilogue of one output tile.  The 8 epilogue warps form TWO independent groups of 4 warps (one warp per
// TMEM lane quarter); group g takes the 64-channel chunks whose running index has parity g, so two chunks are in
// flight at any time.  Per chunk a thread reads its accumulator row (64 columns) from TMEM, applies BN scale/shift
// (+ residual from the staging buffer) (+ ReLU), writes fp16 in place into the 128-byte-swizzled staging buffer, and the
// group's leader TMA-stores it.  (Round 1 ran all 8 warps on ONE chunk at a time: two 256-thread barriers, a proxy fence
// and the store hand-off made ~1.2 us per chunk the critical path of every 1x1 convolution with a residual - 4.9 us per
// 128 x 256 tile where the MMAs need 1.5 us and HBM 3.3 us.)
// `cc` counts staging chunks across tiles (same value in every thread); chunk i uses staging buffer i % NBUF, NBUF even,
// so a buffer always belongs to the same group.  `g` = group of this thread, `leader` = first thread of its group.
template <int BN, int NBUF, int EPI_THREADS>
__device__ __forceinline__ void conv_epilogue_tile(const ConvPersParams& p, const TileCoord& c, uint32_t taddr,
                                                   uint8_t* stg, uint64_t* res_full, uint64_t* res_empty,
                                                   uint64_t* acc_empty_a, uint32_t& cc, uint32_t row_off, uint32_t sw,
                                                   int g, int lane, bool leader, const CUtensorMap& tmO) {
  static_assert(NBUF % 2 == 0, "staging buffers are split between the two epilogue groups");
  constexpr int CHUNKS = BN / 64;
  constexpr int STG_BYTES = 128 * 128;
  constexpr int GROUP_THREADS = 128;
  int my_last = -1;                                         // last chunk of this tile that this group reads from TMEM
#pragma unroll
  for (int ch = 0; ch < CHUNKS; ++ch)
    if (static_cast<int>((cc + ch) & 1u) == g) my_last = ch;
  if (my_last < 0) {                                        // (BN = 64: one chunk per tile, the other group only releases)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(acc_empty_a);
  }
#pragma unroll 1
  for (int ch = 0; ch < CHUNKS; ++ch) {
    const uint32_t ccc = cc + ch;
    if (static_cast<int>(ccc & 1u) != g) continue;
    const int b = ccc % NBUF;
    uint8_t* buf = stg + b * STG_BYTES;
    const uint32_t buf_addr = smem_u32(buf);
    if (p.has_res) {
      mbar_wait(&res_full[b], (ccc / NBUF) & 1);            // residual chunk has landed in `buf`
    } else {
      if (leader) bulk_wait_read<NBUF / 2 - 1>();            // this group's store NBUF/2 chunks ago (same buffer) has left it
      named_bar_sync(1 + 2 * g, GROUP_THREADS);
    }
    const int col0 = c.n_tile * BN + ch * 64;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float v[32];
      tmem_ld32(taddr + ch * 64 + half * 32, v);
      tmem_ld_wait();
      if (ch == my_last && half == 1) {                     // last TMEM read of this warp for this tile
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty_a);
      }
      const float4* sc4 = reinterpret_cast<const float4*>(p.scale + col0 + half * 32);
      const float4* sh4 = reinterpret_cast<const float4*>(p.shift + col0 + half * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 s = __ldg(sc4 + q), h = __ldg(sh4 + q);
        v[4 * q + 0] = fmaf(v[4 * q + 0], s.x, h.x);
        v[4 * q + 1] = fmaf(v[4 * q + 1], s.y, h.y);
        v[4 * q + 2] = fmaf(v[4 * q + 2], s.z, h.z);
        v[4 * q + 3] = fmaf(v[4 * q + 3], s.w, h.w);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {                          // 4 x 16-byte chunks (8 channels each) of this half
        const uint32_t chunk = static_cast<uint32_t>(half * 4 + j);
        const uint32_t sp = buf_addr + row_off + ((chunk ^ sw) << 4);
        if (p.has_res) {
          const uint4 r = lds128(sp);
          const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = unpack_h2(rr[e]);
            v[j * 8 + e * 2] += f.x;
            v[j * 8 + e * 2 + 1] += f.y;
          }
        }
        if (p.relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[j * 8 + e] = fmaxf(v[j * 8 + e], 0.0f);
        }
        uint4 o;
        o.x = pack_h2(v[j * 8 + 0], v[j * 8 + 1]);
        o.y = pack_h2(v[j * 8 + 2], v[j * 8 + 3]);
        o.z = pack_h2(v[j * 8 + 4], v[j * 8 + 5]);
        o.w = pack_h2(v[j * 8 + 6], v[j * 8 + 7]);
        sts128(sp, o);
      }
    }
    fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the TMA engine
    named_bar_sync(2 + 2 * g, GROUP_THREADS);
    if (leader) {
      if (p.a_spatial) tma_store_4d(&tmO, buf, col0, c.wo0, c.ho0, c.n0);
      else tma_store_2d(&tmO, buf, col0, c.m_tile * 128);
      bulk_commit();
      if (p.has_res) {
        // hand a staging buffer back to the residual producer once its store has finished reading it: with NG buffers
        // per group the group's stores older than the latest NG - 1 are complete, i.e. the one of chunk ccc - 2 (NG - 1)
        // (NG = 1: this very chunk - the group's only buffer must be free before its next residual can land)
        constexpr int NG = NBUF / 2;
        if (p.epi_mode & 2) {
          bulk_wait_read<0>();                             // this very store has left `buf`: the residual of chunk ccc + NBUF may land
          mbar_arrive(&res_empty[b]);
        } else {
          bulk_wait_read<NG - 1>();
          if (ccc >= 2u * (NG - 1)) mbar_arrive(&res_empty[(ccc - 2u * (NG - 1)) % NBUF]);
        }
      }
    }
  }
  cc += CHUNKS;
}

// One-group organisation: all 8 epilogue warps work on the same 64-channel chunk, warp group `hsel` taking one 32-channel
// half of it.  The per-chunk dependency chain of a warp (TMEM read -> scale/shift -> residual read -> pack -> store ->
// proxy fence -> barrier) is what bounds the 1x1 convolutions with a residual (ncu warp-state samples of round 2: the
// MMA warp spends 57 % of its time waiting for a drained accumulator, the epilogue warps never wait for data), so the
// long-latency links are kept out of it:
//   * the accumulator columns of chunk ch + 1 are requested from TMEM (tcgen05.ld is asynchronous) before chunk ch is
//     processed, two register arrays alternate;
//   * BN scale / shift of the tile's BN channels are staged in shared memory once per tile (`ss_addr`, SS = true) and read
//     back with broadcast LDS instead of 16 LDG.128 per chunk and thread;
//   * staging-buffer accesses are explicit ld.shared / st.shared (see ptx.cuh: lds128).
// The staging buffer of chunk cc - 1 goes back to the residual producer after the store of chunk cc (epi_mode bit 1: the
// buffer of chunk cc itself, after its own store has been read).
// CPT = channels per thread and chunk: 32 with 8 epilogue warps (two per TMEM lane quarter), 16 with 16 warps (four per
// quarter); `slice` = which CPT-channel slice of the 64-channel chunk this warp owns.
template <int BN, int NBUF, int EPI_THREADS, bool SS, int CPT>
__device__ __forceinline__ void conv_epilogue_chunk_1g(const ConvPersParams& p, const TileCoord& c, float (&v)[CPT], int ch,
                                                       uint32_t buf_addr, int b, uint32_t cc, uint32_t ss_addr,
                                                       uint64_t* res_empty, uint32_t row_off, uint32_t sw, int slice,
                                                       bool leader, const CUtensorMap& tmO) {
  const int col0 = c.n_tile * BN + ch * 64;
  if (SS) {
    const uint32_t sa = ss_addr + static_cast<uint32_t>(ch * 64 + slice * CPT) * 4u;
#pragma unroll
    for (int q = 0; q < CPT / 4; ++q) {
      const float4 s = lds128f(sa + q * 16), h = lds128f(sa + BN * 4 + q * 16);
      v[4 * q + 0] = fmaf(v[4 * q + 0], s.x, h.x);
      v[4 * q + 1] = fmaf(v[4 * q + 1], s.y, h.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], s.z, h.z);
      v[4 * q + 3] = fmaf(v[4 * q + 3], s.w, h.w);
    }
  } else {
    const float4* sc4 = reinterpret_cast<const float4*>(p.scale + col0 + slice * CPT);
    const float4* sh4 = reinterpret_cast<const float4*>(p.shift + col0 + slice * CPT);
#pragma unroll
    for (int q = 0; q < CPT / 4; ++q) {
      const float4 s = __ldg(sc4 + q), h = __ldg(sh4 + q);
      v[4 * q + 0] = fmaf(v[4 * q + 0], s.x, h.x);
      v[4 * q + 1] = fmaf(v[4 * q + 1], s.y, h.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], s.z, h.z);
      v[4 * q + 3] = fmaf(v[4 * q + 3], s.w, h.w);
    }
  }
#pragma unroll
  for (int j = 0; j < CPT / 8; ++j) {                    // 16-byte pieces (8 channels each) of this slice
    const uint32_t chunk = static_cast<uint32_t>(slice * (CPT / 8) + j);
    const uint32_t sp = buf_addr + row_off + ((chunk ^ sw) << 4);
    if (p.has_res) {
      const uint4 r = lds128(sp);
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_h2(rr[e]);
        v[j * 8 + e * 2] += f.x;
        v[j * 8 + e * 2 + 1] += f.y;
      }
    }
    if (p.relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[j * 8 + e] = fmaxf(v[j * 8 + e], 0.0f);
    }
    uint4 o;
    o.x = pack_h2(v[j * 8 + 0], v[j * 8 + 1]);
    o.y = pack_h2(v[j * 8 + 2], v[j * 8 + 3]);
    o.z = pack_h2(v[j * 8 + 4], v[j * 8 + 5]);
    o.w = pack_h2(v[j * 8 + 6], v[j * 8 + 7]);
    sts128(sp, o);
  }
  fence_proxy_async_smem();                              // generic-proxy smem writes -> visible to the TMA engine
  named_bar_sync(2, EPI_THREADS);
  if (leader) {
    if (p.a_spatial) tma_store_4d_addr(&tmO, buf_addr, col0, c.wo0, c.ho0, c.n0);
    else tma_store_2d_addr(&tmO, buf_addr, col0, c.m_tile * 128);
    bulk_commit();
    if (p.has_res) {
      if (p.epi_mode & 2) {
        bulk_wait_read<0>();
        mbar_arrive(&res_empty[b]);
      } else if (cc >= 1) {                               // the previous chunk's store has finished reading its
        bulk_wait_read<1>();                              // buffer -> hand that buffer back to the residual producer
        mbar_arrive(&res_empty[(cc - 1) % NBUF]);
      }
    }
  }
}

template <int BN, int NBUF, int EPI_THREADS, bool SS>
__device__ __forceinline__ void conv_epilogue_tile_1g(const ConvPersParams& p, const TileCoord& c, uint32_t taddr,
                                                      uint32_t stg_addr, uint32_t ss_addr, uint64_t* res_full,
                                                      uint64_t* res_empty, uint64_t* acc_empty_a, uint32_t& cc,
                                                      uint32_t row_off, uint32_t sw, int slice, int lane, bool leader,
                                                      const CUtensorMap& tmO) {
  constexpr int CHUNKS = BN / 64;
  constexpr int STG_BYTES = 128 * 128;
  constexpr int CPT = EPI_THREADS >= 512 ? 16 : 32;      // 32 channels per thread and chunk with 8 warps, 16 with 16
  if (SS) {                                              // BN scale | shift of this tile's channels -> shared memory
    const int j = static_cast<int>(threadIdx.x) - 128;
    if (j < BN) {
      sts32f(ss_addr + j * 4, __ldg(p.scale + c.n_tile * BN + j));
      sts32f(ss_addr + (BN + j) * 4, __ldg(p.shift + c.n_tile * BN + j));
    }
    named_bar_sync(6, EPI_THREADS);
  }
  float va[CPT], vb[CPT];
  tmem_ld_cols<CPT>(taddr + slice * CPT, va);
#pragma unroll
  for (int ch = 0; ch < CHUNKS; ++ch, ++cc) {
    const int b = cc % NBUF;
    const uint32_t buf_addr = stg_addr + b * STG_BYTES;
    if (p.has_res) {
      mbar_wait(&res_full[b], (cc / NBUF) & 1);           // residual chunk has landed in the buffer
    } else {
      if (leader) bulk_wait_read<NBUF - 1>();               // the store issued NBUF chunks ago has left the buffer
      named_bar_sync(1, EPI_THREADS);
    }
    tmem_ld_wait();                                        // this chunk's accumulator columns are in registers
    if (ch + 1 < CHUNKS) {
      tmem_ld_cols<CPT>(taddr + (ch + 1) * 64 + slice * CPT, (ch & 1) ? va : vb);
    } else {                                               // last TMEM read of this tile: release the accumulator
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty_a);
    }
    conv_epilogue_chunk_1g<BN, NBUF, EPI_THREADS, SS, CPT>(p, c, (ch & 1) ? vb : va, ch, buf_addr, b, cc, ss_addr, res_empty,
                                                           row_off, sw, slice, leader, tmO);
  }
}

// Epilogue warps: 8 for the convolution epilogue (its per-chunk critical path bounds the memory-bound 1x1
// convolutions), 4 for the light similarity epilogues.
template <int EPI, int EW = 0>
struct PersThreads {
  static constexpr int EPI_WARPS = EW > 0 ? EW : ((EPI == 0) ? 8 : 4);
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
};

// EW = 0: default number of epilogue warps (8 for the convolution epilogue, 4 otherwise); EW = 16 (convolution epilogue
// only): four warps per TMEM lane quarter, each taking 16 of the 64 channels of a chunk - for the convolutions whose
// tile time is the epilogue's dependency chain (1x1 with a residual, short K).
template <int BN, int STAGES, int EPI, int NB = 4, int EW = 0>
__global__ void __launch_bounds__(PersThreads<EPI, EW>::THREADS, 1)
conv_pers_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                 const ConvPersParams p) {
  using L = ConvPersSmem<BN, STAGES, EPI, NB>;
  constexpr int NBUF = NB;
  constexpr int CHUNKS = BN / 64;
  if (EPI != PERS_EPI_CONV && p.gate != nullptr) {   // uniform over the grid: nothing has been allocated yet
    pdl_wait();
    if (*reinterpret_cast<const volatile int*>(p.gate) == 0) return;
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg = smem + L::STG_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* res_full = acc_empty + 2;
  uint64_t* res_empty = res_full + NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + NBUF);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int k_iters = p.taps * p.cin_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI == PERS_EPI_CONV) tma_prefetch_desc(&tmO);
    if (EPI == PERS_EPI_CONV && (p.has_res || p.k_split > 0)) tma_prefetch_desc(&tmR);
    if (EPI == PERS_EPI_F32) { tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmO); }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], PersThreads<EPI, EW>::EPI_WARPS);  // one arrival per epilogue warp
    }
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(&res_full[b], 1);
      mbar_init(&res_empty[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 3) tmem_alloc(tmem_slot, L::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // the next kernel may start its own prologue while this one runs / drains
  pdl_wait();                // everything below reads tensors written by the previous kernel of the stream

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------ A/B producer
      uint32_t g = 0;  // ring slot counter, runs across tiles
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        for (int it = 0; it < k_iters; ++it, ++g) {
          const int s = g % STAGES;
          if (EPI == PERS_EPI_CONV && p.l2_prefetch > 0 && p.taps == 1 && p.k_split == 0) {
            // constant look-ahead: the activation box this CTA will load l2_prefetch ring slots from now is requested into
            // L2 (one request per slot, interleaved with the loads - a burst per tile would queue in front of them)
            const uint32_t j = g + static_cast<uint32_t>(p.l2_prefetch);
            const int tj = blockIdx.x + static_cast<int>(j / k_iters) * static_cast<int>(gridDim.x);
            if (tj < p.total_tiles) {
              const TileCoord cn = decode_tile(p, tj);
              if (p.n_tiles == 1 || cn.n_tile == 0) {           // (the CTAs of the other channel slices share these boxes)
                const int itj = static_cast<int>(j % k_iters);
                if (p.a_spatial) tma_prefetch_4d(&tmA, itj * 64, cn.wo0 * p.stride - p.pad, cn.ho0 * p.stride - p.pad, cn.n0);
                else tma_prefetch_2d(&tmA, itj * 64, cn.m_tile * 128);
              }
            }
          }
          mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
          mbar_expect_tx(&full_bar[s], L::STAGE_BYTES);
          uint8_t* sa = smem + s * L::STAGE_BYTES;
          const int tap = it / p.cin_blocks;
          const int kc = it - tap * p.cin_blocks;
          if (EPI == PERS_EPI_F32) {
            const int part = it / p.k_per_part;
            const int kk = it - part * p.k_per_part;
            // small cross terms first, (hi, hi) last: the tensor core's fp32 accumulation truncates, an error that grows
            // with the accumulator's magnitude per step - the 2/3 of the steps that add 2^-11-sized terms then run
            // while the accumulator is still small (measured: 1.1e-5 -> see tests/test_gpu_ops.py)
            tma_load_2d(sa, part == 1 ? &tmR : &tmA, &full_bar[s], kk * 64, c.m_tile * 128);
            tma_load_2d(sa + L::A_BYTES, part == 0 ? &tmO : &tmB, &full_bar[s], kk * 64, c.n_tile * BN);
            continue;
          }
          if (p.k_split > 0 && it >= p.k_split) {
            tma_load_4d(sa, &tmR, &full_bar[s], (it - p.k_split) * 64, c.wo0 * p.a2_stride, c.ho0 * p.a2_stride, c.n0);
          } else if (p.a_spatial) {
            const int kh = tap / p.kw_taps;
            const int kw = tap - kh * p.kw_taps;
            tma_load_4d(sa, &tmA, &full_bar[s], kc * 64, c.wo0 * p.stride + kw - p.pad,
                        c.ho0 * p.stride + kh - p.pad, c.n0);
          } else {
            tma_load_2d(sa, &tmA, &full_bar[s], it * 64, c.m_tile * 128);
          }
          tma_load_2d(sa + L::A_BYTES, &tmB, &full_bar[s], it * 64, c.n_tile * BN);
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------- MMA issuer
    // The whole warp walks the loop (uniform control flow keeps descriptors and TMEM addresses in uniform
    // registers); one elected lane issues the MMAs and the commits of a ring slot.
    constexpr uint32_t idesc = umma_idesc_f16(128, BN);
    uint32_t g = 0, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const uint32_t a = i & 1;
      mbar_wait(&acc_empty[a], ((i >> 1) & 1) ^ 1);   // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * BN;
      for (int it = 0; it < k_iters; ++it, ++g) {
        const int s = g % STAGES;
        mbar_wait(&full_bar[s], (g / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * L::STAGE_BYTES);
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + L::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
          umma_commit(&empty_bar[s]);
          if (it == k_iters - 1) umma_commit(&acc_full[a]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    if (EPI == PERS_EPI_CONV && lane == 0 && p.has_res) {
      // ------------------------------------------------------------ residual producer
      uint32_t cc = 0;  // staging chunk counter, runs across tiles
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const TileCoord c = decode_tile(p, t);
        for (int ch = 0; ch < CHUNKS; ++ch, ++cc) {
          const int b = cc % NBUF;
          if (p.l2_prefetch > 0) {                            // residual chunk NBUF + 4 chunks from now -> L2
            const uint32_t j = cc + NBUF + 4;
            const int tj = blockIdx.x + static_cast<int>(j / CHUNKS) * static_cast<int>(gridDim.x);
            if (tj < p.total_tiles) {
              const TileCoord cn = decode_tile(p, tj);
              const int chj = static_cast<int>(j % CHUNKS);
              if (p.a_spatial) tma_prefetch_4d(&tmR, cn.n_tile * BN + chj * 64, cn.wo0, cn.ho0, cn.n0);
              else tma_prefetch_2d(&tmR, cn.n_tile * BN + chj * 64, cn.m_tile * 128);
            }
          }
          mbar_wait(&res_empty[b], ((cc / NBUF) & 1) ^ 1);
          mbar_expect_tx(&res_full[b], L::STG_BYTES);
          if (p.a_spatial)
            tma_load_4d(stg + b * L::STG_BYTES, &tmR, &res_full[b], c.n_tile * BN + ch * 64, c.wo0, c.ho0, c.n0);
          else
            tma_load_2d(stg + b * L::STG_BYTES, &tmR, &res_full[b], c.n_tile * BN + ch * 64, c.m_tile * 128);
        }
      }
    }
  } else if (warp >= 4) {
    // -------------------------------------------------------------- epilogue
    constexpr int EPI_THREADS = 32 * PersThreads<EPI, EW>::EPI_WARPS;
    const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
    const int hsel = (warp - 4) >> 2;                  // conv epilogue: which channel slice of a chunk (one-group) / which group
    const int row = quarter * 32 + lane;
    const bool two_groups = (EPI == PERS_EPI_CONV) && (p.epi_mode & 1) && PersThreads<EPI, EW>::EPI_WARPS == 8;
    const bool leader = two_groups ? (threadIdx.x == 128u + 128u * hsel) : (threadIdx.x == 128);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    const uint32_t sw = static_cast<uint32_t>(row & 7);
    uint32_t cc = 0, i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      const TileCoord c = decode_tile(p, t);
      const uint32_t a = i & 1;
      mbar_wait(&acc_full[a], (i >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + a * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      if (EPI == PERS_EPI_CONV) {
        if (two_groups)
          conv_epilogue_tile<BN, NBUF, EPI_THREADS>(p, c, taddr, stg, res_full, res_empty, &acc_empty[a], cc, row_off, sw,
                                                    hsel, lane, leader, tmO);
        else
          conv_epilogue_tile_1g<BN, NBUF, EPI_THREADS, L::SS>(p, c, taddr, smem_u32(stg), smem_u32(smem + L::SS_OFF) + (i & 1) * 2 * BN * 4,
                                                              res_full, res_empty, &acc_empty[a], cc, row_off, sw, hsel, lane,
                                                              leader, tmO);
      } else {
        // ---------------------------------------------------------- similarity epilogues: row = query
        const int64_t qi = static_cast<int64_t>(c.m_tile) * 128 + row;
        const bool valid = qi < p.M;
        if (EPI == PERS_EPI_SIM_FILTER) {
          // Candidate capture in two passes over the accumulator row: pass 1 only builds the match masks of the BN / 32
          // column groups, then ONE atomicAdd reserves the thread's slots for the whole tile, pass 2 re-reads the
          // groups that matched and writes them.  (One atomic per group made its ~1.5 us round trip the critical path
          // of the tile as soon as a few groups per row matched: +118 us on a 125k-row shard.)
          const float t_q = valid ? __ldg(p.thr + qi) : INFINITY;
          uint32_t masks[BN / 32];                            // (dynamic index: 32 bytes of local memory, L1-resident)
          int total = 0;
#pragma unroll 1
          for (int g = 0; g < BN / 32; ++g) {
            float v[32];
            tmem_ld32(taddr + g * 32, v);
            tmem_ld_wait();
            const int nb0 = c.n_tile * BN + g * 32;
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) mask |= (v[j] >= t_q && nb0 + j < p.N) ? (1u << j) : 0u;
            masks[g] = mask;
            total += __popc(mask);
          }
          const bool any = __any_sync(0xffffffffu, total != 0);
          if (any) {                                          // warp-uniform: tcgen05.ld is warp-collective
            int pos = total ? atomicAdd(p.cand_cnt + qi, total) : 0;
            unsigned long long* dst = p.cand + qi * p.cand_cap;
#pragma unroll 1
            for (int g = 0; g < BN / 32; ++g) {
              if (!__any_sync(0xffffffffu, masks[g] != 0)) continue;
              float v[32];
              tmem_ld32(taddr + g * 32, v);
              tmem_ld_wait();
              const int nb0 = c.n_tile * BN + g * 32;
              const uint32_t mask = masks[g];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (mask & (1u << j)) {
                  if (pos < p.cand_cap)
                    dst[pos] = (static_cast<unsigned long long>(__float_as_uint(v[j])) << 32) | static_cast<unsigned int>(nb0 + j);
                  ++pos;
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[a]);
          continue;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          float v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (c0 + 32 == BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
          }
          const int nb0 = c.n_tile * BN + c0;
          if (!valid || nb0 >= p.N) continue;
          if (EPI == PERS_EPI_F32) {
            float* dp = p.dense + qi * p.dense_ld + nb0;
            const bool vec = (nb0 + 32 <= p.N) && ((p.dense_ld & 3) == 0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= (p.scale != nullptr && nb0 + j < p.N) ? __ldg(p.scale + nb0 + j) : 1.0f;
            if (vec) {
              float4* d4 = reinterpret_cast<float4*>(dp);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) d4[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (nb0 + j < p.N) dp[j] = v[j];
            }
          } else if (EPI == PERS_EPI_SIM_GMAX) {
            // maximum of each group of 32 database rows: the k-th largest group maximum is a valid lower bound on
            // the k-th best score (k distinct rows reach it) and costs 1/32 of the dense write + select
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) m = (nb0 + j < p.N) ? fmaxf(m, v[j]) : m;
            p.dense[qi * p.dense_ld + (nb0 >> 5)] = m;
          } else if (EPI == PERS_EPI_SIM_DENSE) {
            float* dp = p.dense + qi * p.dense_ld + nb0;
            if (nb0 + 32 <= p.N) {
              float4* d4 = reinterpret_cast<float4*>(dp);
#pragma unroll
              for (int q4 = 0; q4 < 8; ++q4) d4[q4] = make_float4(v[4 * q4], v[4 * q4 + 1], v[4 * q4 + 2], v[4 * q4 + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (nb0 + j < p.N) dp[j] = v[j];
            }
          }
        }
      }
    }
    if (EPI == PERS_EPI_CONV && leader) bulk_wait<0>();     // all output bytes written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc(tmem_base, L::TMEM_COLS);
  }
}

template <int BN, int STAGES, int EPI, int NB = 4, int EW = 0>
int conv_pers_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmR, const CUtensorMap& tmO,
                     const ConvPersParams& p, int num_sms, cudaStream_t stream) {
  using L = ConvPersSmem<BN, STAGES, EPI, NB>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  auto kern = conv_pers_kernel<BN, STAGES, EPI, NB, EW>;
  static std::atomic<uint64_t> attr_done{0};
  if (first_launch_on_device(attr_done))
    DIRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  DIRB_CUDA(launch_pdl(kern, dim3(grid), dim3(PersThreads<EPI, EW>::THREADS), L::TOTAL, stream, tmA, tmB, tmR, tmO, p));
  count_launch();
  return 0;
}

}  // namespace dirb
