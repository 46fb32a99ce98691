// Similarity + top-k over one row shard of the descriptor database.
//
// Replaces the reference's dense  scores = matmul(q, db)  (dirtorch/utils/common.py:30-38, test_dir.py:145) +
// per-query sort (datasets/generic.py:207,221) for the first k ranks, without materialising the Q x N matrix:
//
//   1. queries -> fp16                                                             (f32_to_f16)
//   2. seed:   tcgen05 GEMM of the queries against the first S rows.  The epilogue keeps the maximum of every group
//              of 32 consecutive rows (PERS_EPI_SIM_GMAX; 1/32 of the dense score traffic); per query the k-th largest
//              group maximum t_S has >= k rows at or above it, so it is a lower bound on the final k-th score.
//              (fewer than k groups: dense scores, PERS_EPI_SIM_DENSE)                 (kth_dense_kernel)
//   3. filter: tcgen05 GEMM against all N rows; the epilogue appends (score,row) to the query's candidate list
//              only when score >= t_S - 2*eps16                                     (PERS_EPI_SIM_FILTER)
//      (for N <= S step 3 is a scan of the dense scores instead: dense_compact_kernel)
//   4. select: exact k-th largest candidate score t (radix select, cand_kth_kernel); survivors = candidates with
//              score >= t - 2*eps16                                                    (cand_survivors_kernel)
//   5. rescore the survivors exactly (fp64 accumulation of the fp32 rows, rescore_kernel) and sort
//      (score desc, index asc: sort_topk_kernel).
//
// eps16 bounds |fp16-path score - exact score| (unit-norm rows: 2 * 2^-11 from the operand roundings + fp32
// accumulation, default 1.2e-3); any row of the true top-k then satisfies the step-3 and step-4 conditions, so the
// result is the exact top-k (same argument twice).  Candidate-buffer overflow raises the threshold from what was
// captured and re-runs the filter pass.  With several shards the phases are split (search_begin / search_finish)
// so that the caller can MIN-reduce the per-shard selection thresholds in between (see cand_kth_kernel).
// The GEMMs run on the persistent warp-specialised kernel of conv_pers.cuh (128 x 256 tiles, K = D).
#include <math.h>

#include <algorithm>
#include <string>
#include <vector>

#include "conv.h"
#include "conv_pers.cuh"

namespace dirb {
namespace {

constexpr int SEL_THREADS = 512;

__device__ __forceinline__ uint32_t f2key(float f) {  // monotone increasing map float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct DenseAcc {
  const float* p;
  __device__ float operator()(int i) const { return p[i]; }
};
struct CandAcc {
  const unsigned long long* p;
  __device__ float operator()(int i) const { return __uint_as_float(static_cast<uint32_t>(p[i] >> 32)); }
};

// k-th largest of n values (block-wide MSB-first radix select, 4 passes of 8 bits). n < k -> -inf.
template <class Acc>
__device__ float block_kth_largest(Acc acc, int n, int k, uint32_t* hist /*[256]*/, uint32_t* bc /*[2]*/) {
  if (n < k) return -INFINITY;
  uint32_t prefix = 0, mask = 0;
  int kk = k;
  for (int pass = 3; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // similarity scores share their high bits, so most keys of a pass land in one or two bins: aggregate the
    // lanes of a warp that hit the same bin and let one of them add the whole group
    const int n_round = (n + 31) & ~31;
    for (int i = threadIdx.x; i < n_round; i += blockDim.x) {
      uint32_t bin = 0xffffffffu;
      if (i < n) {
        const uint32_t key = f2key(acc(i));
        if ((key & mask) == prefix) bin = (key >> shift) & 255u;
      }
      const uint32_t peers = __match_any_sync(0xffffffffu, bin);
      if (bin != 0xffffffffu && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&hist[bin], static_cast<uint32_t>(__popc(peers)));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, b = 255;
      for (; b > 0; --b) {
        if (cum + static_cast<int>(hist[b]) >= kk) break;
        cum += hist[b];
      }
      bc[0] = b;
      bc[1] = kk - cum;
    }
    __syncthreads();
    prefix |= bc[0] << shift;
    mask |= 255u << shift;
    kk = bc[1];
    __syncthreads();
  }
  return key2f(prefix);
}

// thr[q] = (k-th largest of dense[q][0..S)) - band
__global__ void __launch_bounds__(SEL_THREADS) kth_dense_kernel(const float* __restrict__ dense, int64_t ld, int S,
                                                                int k, float band, float* __restrict__ thr) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bc[2];
  const int q = blockIdx.x;
  const float t = block_kth_largest(DenseAcc{dense + q * ld}, S, k, hist, bc);
  if (threadIdx.x == 0) thr[q] = t - band;
}

// N <= S: candidates straight from the dense scores.
__global__ void dense_compact_kernel(const float* __restrict__ dense, int64_t ld, int N, const float* __restrict__ thr,
                                     unsigned long long* __restrict__ cand, int* __restrict__ cnt, int cap) {
  const int q = blockIdx.y;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float v = dense[q * ld + n];
  if (v >= thr[q]) {
    const int pos = atomicAdd(cnt + q, 1);
    if (pos < cap)
      cand[static_cast<int64_t>(q) * cap + pos] = (static_cast<unsigned long long>(__float_as_uint(v)) << 32) | static_cast<unsigned>(n);
  }
}

// Per query: exact k-th and k_shard-th largest candidate scores (fp16-path scores).
//   kth_k[q]  : local k-th best - always a valid lower bound on the global k-th best
//   sel[q]    : local min(k_shard, N)-th best; the MINIMUM of this value over all shards is a valid and much tighter
//               lower bound on the global k-th best as long as the shards certify k rows between them,
//               sum_g min(k_shard, N_g) >= min(k, N_total) (shard g holds min(k_shard, N_g) rows at or above its own
//               value) - the caller picks k_shard accordingly (ceil(k / shards) for evenly filled shards, see
//               dist.py: shard_quota) and min-reduces sel.
// flags[q] bit0 = candidate overflow (cnt > cap): the tighter threshold kth(captured) - band is written to thr[q];
// otherwise thr[q] = +inf so that a re-run of the filter pass leaves this query alone.
__global__ void __launch_bounds__(SEL_THREADS) cand_kth_kernel(const unsigned long long* __restrict__ cand,
                                                               const int* __restrict__ cnt, int cap, int k, int k_shard,
                                                               float band, float* __restrict__ kth_k,
                                                               float* __restrict__ sel, float* __restrict__ thr,
                                                               int* __restrict__ flags, int64_t n_rows) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t bc[2];
  const int q = blockIdx.x;
  const int total = cnt[q];
  const int n = min(total, cap);
  const unsigned long long* c = cand + static_cast<int64_t>(q) * cap;
  const int kk = (n_rows < static_cast<int64_t>(k)) ? static_cast<int>(n_rows) : k;
  const float t = block_kth_largest(CandAcc{c}, n, kk, hist, bc);
  if (total > cap) {
    if (threadIdx.x == 0) {
      thr[q] = t - band;
      flags[q] = 1;
    }
    return;
  }
  float ts = t;
  if (k_shard < kk) {
    __syncthreads();
    ts = block_kth_largest(CandAcc{c}, n, k_shard, hist, bc);
  }
  if (threadIdx.x == 0) {
    kth_k[q] = t;
    sel[q] = ts;
    flags[q] = 0;
    thr[q] = INFINITY;
  }
}

// Survivors = candidates with score >= max(sel[q], kth_k[q]) - band.  flags[q] bit1 = more than cap2 survivors.
__global__ void __launch_bounds__(SEL_THREADS) cand_survivors_kernel(const unsigned long long* __restrict__ cand,
                                                                     const int* __restrict__ cnt, int cap,
                                                                     const float* __restrict__ kth_k,
                                                                     const float* __restrict__ sel, float band,
                                                                     int* __restrict__ surv_idx, int* __restrict__ surv_cnt,
                                                                     int cap2, int* __restrict__ flags) {
  __shared__ int s_n;
  const int q = blockIdx.x;
  const int n = min(cnt[q], cap);
  const unsigned long long* c = cand + static_cast<int64_t>(q) * cap;
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const float t2 = fmaxf(sel[q], kth_k[q]) - band;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long e = c[i];
    if (__uint_as_float(static_cast<uint32_t>(e >> 32)) >= t2) {
      const int pos = atomicAdd(&s_n, 1);
      if (pos < cap2) surv_idx[static_cast<int64_t>(q) * cap2 + pos] = static_cast<int>(e & 0xffffffffu);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    surv_cnt[q] = min(s_n, cap2);
    flags[q] = (s_n > cap2) ? 2 : 0;
  }
}

// One warp per (query, survivor): exact score = fp64-accumulated dot product of the fp32 rows.
__global__ void rescore_kernel(const float* __restrict__ q32, const float* __restrict__ db32, int D,
                               const int* __restrict__ surv_idx, const int* __restrict__ surv_cnt, int cap2,
                               double* __restrict__ surv_score) {
  const int q = blockIdx.y;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= surv_cnt[q]) return;
  const int row = surv_idx[static_cast<int64_t>(q) * cap2 + w];
  const float4* a = reinterpret_cast<const float4*>(q32 + static_cast<int64_t>(q) * D);
  const float4* b = reinterpret_cast<const float4*>(db32 + static_cast<int64_t>(row) * D);
  double acc = 0.0;
  for (int i = lane; i < D / 4; i += 32) {
    const float4 x = __ldg(a + i), y = __ldg(b + i);
    acc += static_cast<double>(x.x) * y.x;
    acc += static_cast<double>(x.y) * y.y;
    acc += static_cast<double>(x.z) * y.z;
    acc += static_cast<double>(x.w) * y.w;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) surv_score[static_cast<int64_t>(q) * cap2 + w] = acc;
}

// Per query: bitonic sort of (score desc, index asc), write the first k.  Also used to merge shard lists.
// in_idx is int32 local rows (+offset) when idx32 != nullptr, else int64 global indices from idx64.
// With shard_k > 0 the input is G shard lists read in place: entry i of query q = element (i % shard_k) of shard
// (i / shard_k), i.e. in[(i / shard_k) * shard_stride + q * shard_k + i % shard_k] (the all-gather buffer layout).
__global__ void __launch_bounds__(1024) sort_topk_kernel(const double* __restrict__ in_score, const int* __restrict__ idx32,
                                                         const int64_t* __restrict__ idx64, const int* __restrict__ cnts,
                                                         int fixed_cnt, int stride, int64_t offset, int k,
                                                         double* __restrict__ out_score, int64_t* __restrict__ out_idx,
                                                         int shard_k = 0, int64_t shard_stride = 0) {
  extern __shared__ uint8_t sm[];
  const int q = blockIdx.x;
  const int n = cnts ? cnts[q] : fixed_cnt;
  int P = 1;
  while (P < n) P <<= 1;
  if (P < 2) P = 2;
  double* sc = reinterpret_cast<double*>(sm);
  int64_t* ix = reinterpret_cast<int64_t*>(sc + P);
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    if (i < n) {
      const int64_t src = shard_k > 0 ? (static_cast<int64_t>(i / shard_k) * shard_stride + static_cast<int64_t>(q) * shard_k + i % shard_k)
                                      : (static_cast<int64_t>(q) * stride + i);
      sc[i] = in_score[src];
      ix[i] = idx32 ? (static_cast<int64_t>(idx32[src]) + offset) : idx64[src];
      if (ix[i] < 0) sc[i] = -INFINITY;  // empty slots of a shard list
    } else {
      sc[i] = -INFINITY;
      ix[i] = INT64_MAX;
    }
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int st = size >> 1; st > 0; st >>= 1) {
      for (int i = threadIdx.x; i < P / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (st - 1));
        const int hi = lo + st;
        const bool up = ((lo & size) == 0);  // "up" = this block sorted best-first
        const double a = sc[lo], b = sc[hi];
        const int64_t ia = ix[lo], ib = ix[hi];
        const bool a_first = (a > b) || (a == b && ia < ib);
        if (a_first != up) {
          sc[lo] = b; sc[hi] = a;
          ix[lo] = ib; ix[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const bool ok = (i < n) && (ix[i] >= 0) && (ix[i] != INT64_MAX);
    out_score[static_cast<int64_t>(q) * k + i] = ok ? sc[i] : -INFINITY;
    out_idx[static_cast<int64_t>(q) * k + i] = ok ? ix[i] : -1;
  }
}

// Exact dense scores for small evaluation sets: grid (ceil(N/8), ceil(Q/4)), warp = one db row x 4 queries.
__global__ void scores_exact_kernel(const float* __restrict__ q, int Q, const float* __restrict__ db, int64_t N, int D,
                                    float* __restrict__ out) {
  const int64_t n = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int q0 = blockIdx.y * 4;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  double acc[4] = {0, 0, 0, 0};
  const float4* b = reinterpret_cast<const float4*>(db + n * D);
  for (int i = lane; i < D / 4; i += 32) {
    const float4 y = __ldg(b + i);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (q0 + j < Q) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(q + static_cast<int64_t>(q0 + j) * D) + i);
        acc[j] += static_cast<double>(x.x) * y.x + static_cast<double>(x.y) * y.y + static_cast<double>(x.z) * y.z +
                  static_cast<double>(x.w) * y.w;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && q0 + j < Q) out[static_cast<int64_t>(q0 + j) * N + n] = static_cast<float>(v);
  }
}

// alpha query expansion: out_i = normalize(q_i + sum_j db[idx_ij] * s_ij^alpha)   (test_dir.py:38-42; the mean's
// 1/(k+1) cancels in the normalisation).  partial: un-normalised neighbour sum only.
__global__ void aqe_kernel(const float* __restrict__ q, int D, const float* __restrict__ db, const int64_t* __restrict__ nn,
                           const double* __restrict__ ns, int k, double alpha, int partial, int64_t row_offset,
                           int64_t n_rows, float* __restrict__ out) {
  extern __shared__ float wts[];  // [k] weights, then [k] local rows (as int64 pairs of floats)
  __shared__ float sh[32];
  int64_t* rows = reinterpret_cast<int64_t*>(wts + ((k + 1) & ~1));
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    int64_t r = nn[static_cast<int64_t>(i) * k + j];
    if (r >= 0 && (partial || n_rows > 0)) {         // sharded: global index -> local row of this shard, or "not mine"
                                                     // (partial sums are per shard by definition, also for an EMPTY shard)
      r -= row_offset;
      if (r < 0 || r >= n_rows) r = -1;
    }
    rows[j] = r;
    wts[j] = (r >= 0) ? static_cast<float>(pow(ns[static_cast<int64_t>(i) * k + j], alpha)) : 0.f;
  }
  __syncthreads();
  float ss = 0.f;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = partial ? 0.f : q[static_cast<int64_t>(i) * D + c];
    for (int j = 0; j < k; ++j) {
      const int64_t r = rows[j];
      if (r >= 0) acc += db[r * D + c] * wts[j];
    }
    out[static_cast<int64_t>(i) * D + c] = acc;
    ss += acc * acc;
  }
  if (!partial) {
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = ss;
    __syncthreads();
    float t = (l < (blockDim.x >> 5)) ? sh[l] : 0.f;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const float inv = 1.0f / sqrtf(t);
    for (int c = threadIdx.x; c < D; c += blockDim.x) out[static_cast<int64_t>(i) * D + c] *= inv;
  }
}

}  // namespace
}  // namespace dirb

using namespace dirb;

struct dirb200_index {
  int device = 0, dim = 0;
  const float* db32 = nullptr;
  const __half* db16 = nullptr;
  int64_t N = 0, offset = 0;
  bool has_db = false;
  double eps16 = 1.2e-3;
  int64_t sample_rows = 0;
  int cand_cap = 0;          // 0 = auto
  // workspaces (grown on demand)
  void* ws = nullptr;
  size_t ws_bytes = 0;
  int* h_flags = nullptr;  // pinned
  int h_flags_n = 0;
  int64_t stats[5] = {0, 0, 0, 0, 0};
  float* sel_own = nullptr;        // selection thresholds of the single-shard entry point
  size_t sel_bytes = 0;
  // state of a search between dirb200_index_search_begin and _finish
  struct Pending {
    bool active = false;
    int Q = 0, k = 0, cap = 0, cap2 = 0, retries = 0, mark_i = 0;
    int64_t S = 0, launches0 = 0;
    float band = 0;
    int *cnt = nullptr, *sidx = nullptr, *scnt = nullptr, *flags = nullptr;
    unsigned long long* cand = nullptr;
    double* sscore = nullptr;
    float *kth_k = nullptr, *sel = nullptr;
  } pend;
  int profile = 0;                 // option "profile": time the phases of a search with CUDA events
  cudaEvent_t ev[10] = {};
  double phase_ms[9] = {};
};

// Similarity GEMM on the persistent tcgen05 kernel (conv_pers.cuh): A = queries [Q][D], B = database rows [rows][D],
// 128 x 256 tiles, K = D.  Tiles are ordered database-tile-fastest, so CTAs running at the same time share the
// query tile (L2) and stream disjoint database rows.
struct SimArgs {
  float* dense = nullptr; int64_t dense_ld = 0;
  const float* thr = nullptr; unsigned long long* cand = nullptr; int* cand_cnt = nullptr; int cand_cap = 0;
};
static int sim_gemm(int epi, const __half* q16, int Q, const __half* db16, int64_t rows, int D, const SimArgs& a,
                    cudaStream_t stream) {
  constexpr int BN = 256;
  CUtensorMap tmA, tmB;
  DIRB_TRY(encode_tmap_2d(&tmA, q16, D, Q, (uint64_t)D * 2, 64, 128));
  DIRB_TRY(encode_tmap_2d(&tmB, db16, D, rows, (uint64_t)D * 2, 64, BN));
  ConvPersParams p{};
  p.a_spatial = 0;
  p.taps = 1; p.kw_taps = 1; p.cin_blocks = D / 64; p.stride = 1; p.pad = 0;
  p.tw = 128; p.th = 1; p.nb = 1; p.tiles_w = 1; p.tiles_h = 1;
  p.M = Q;
  p.N = static_cast<int>(rows);
  p.n_tiles = static_cast<int>(ceil_div(rows, BN));
  p.m_tiles = static_cast<int>(ceil_div(Q, 128));
  p.m_fastest = 1;
  const int64_t total = static_cast<int64_t>(p.m_tiles) * p.n_tiles;
  DIRB_REQUIRE(total < (int64_t(1) << 31), DIRB200_ENOTSUP, "too many tiles");
  p.total_tiles = static_cast<int>(total);
  p.dense = a.dense; p.dense_ld = a.dense_ld;
  p.thr = a.thr; p.cand = a.cand; p.cand_cnt = a.cand_cnt; p.cand_cap = a.cand_cap;
  if (epi == PERS_EPI_SIM_DENSE)
    return conv_pers_launch<BN, 4, PERS_EPI_SIM_DENSE>(tmA, tmB, tmA, tmA, p, num_sms(), stream);
  if (epi == PERS_EPI_SIM_GMAX)
    return conv_pers_launch<BN, 4, PERS_EPI_SIM_GMAX>(tmA, tmB, tmA, tmA, p, num_sms(), stream);
  return conv_pers_launch<BN, 4, PERS_EPI_SIM_FILTER>(tmA, tmB, tmA, tmA, p, num_sms(), stream);
}

extern "C" {

int dirb200_index_create(int device, int dim, dirb200_index** out) {
  DIRB_REQUIRE(out != nullptr, DIRB200_EINVAL, "null out");
  DIRB_TRY(dirb200_device_check(device));
  DIRB_REQUIRE(dim > 0 && dim % 64 == 0, DIRB200_ENOTSUP, "descriptor dim must be a multiple of 64 (got %d)", dim);
  auto* h = new dirb200_index();
  h->device = device;
  h->dim = dim;
  *out = h;
  return 0;
}

int dirb200_index_set_db(dirb200_index* h, const float* db32_dev, const void* db16_dev, int64_t N,
                         int64_t index_offset) {
  DIRB_REQUIRE(h && N >= 0 && (N == 0 || (db32_dev && db16_dev)), DIRB200_EINVAL, "bad db arguments");   // an empty shard has no buffers
  DIRB_REQUIRE(N < (int64_t(1) << 31) - 256, DIRB200_ENOTSUP, "shard too large (%lld rows)", (long long)N);
  h->db32 = db32_dev;
  h->db16 = static_cast<const __half*>(db16_dev);
  h->N = N;
  h->offset = index_offset;
  h->has_db = true;
  return 0;
}

int dirb200_index_set_option(dirb200_index* h, const char* key, double value) {
  DIRB_REQUIRE(h && key, DIRB200_EINVAL, "null");
  const std::string k(key);
  if (k == "eps16") h->eps16 = value;
  else if (k == "sample_rows") h->sample_rows = static_cast<int64_t>(value);
  else if (k == "cand_cap") h->cand_cap = static_cast<int>(value);
  else if (k == "profile") h->profile = value != 0;
  else DIRB_REQUIRE(false, DIRB200_EKEY, "unknown index option '%s'", key);
  return 0;
}

int dirb200_index_last_stats(dirb200_index* h, int64_t stats[5]) {
  DIRB_REQUIRE(h && stats, DIRB200_EINVAL, "null");
  for (int i = 0; i < 5; ++i) stats[i] = h->stats[i];
  return 0;
}

int dirb200_index_last_profile(dirb200_index* h, double out9[9]) {
  DIRB_REQUIRE(h && out9, DIRB200_EINVAL, "null");
  for (int i = 0; i < 9; ++i) out9[i] = h->phase_ms[i];
  return 0;
}

int dirb200_index_destroy(dirb200_index* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  for (auto e : h->ev) if (e) cudaEventDestroy(e);
  if (h->ws) cudaFree(h->ws);
  if (h->sel_own) cudaFree(h->sel_own);
  if (h->h_flags) cudaFreeHost(h->h_flags);
  delete h;
  return 0;
}

static void mark_phase(dirb200_index* h, cudaStream_t stream) {
  if (!h->profile || h->pend.mark_i >= 10) return;
  if (!h->ev[h->pend.mark_i]) cudaEventCreate(&h->ev[h->pend.mark_i]);
  cudaEventRecord(h->ev[h->pend.mark_i++], stream);
}

// Phase 1: fp16 queries, seed pass, filter pass (with overflow retries), local k-th / k_shard-th candidate scores.
// sel_dev[Q] receives the local k_shard-th best fp16-path score per query; with several shards the caller
// MIN-reduces it over the shards before phase 2 (k_shard = ceil(k / shards)); with one shard k_shard = k.
int dirb200_index_search_begin(dirb200_index* h, const float* q32, int Q, int k, int k_shard, float* sel_dev,
                               void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DIRB_REQUIRE(h && q32 && sel_dev, DIRB200_EINVAL, "null argument");
  DIRB_REQUIRE(h->has_db, DIRB200_ESTATE, "index has no database attached");
  DIRB_REQUIRE(Q > 0 && k > 0 && k <= 1024, DIRB200_ENOTSUP, "need 0 < Q and 0 < k <= 1024 (got Q=%d k=%d)", Q, k);
  DIRB_REQUIRE(k_shard >= 1 && k_shard <= k, DIRB200_EINVAL, "k_shard must be in [1, k]");
  DIRB_CUDA(cudaSetDevice(h->device));
  const int D = h->dim;
  const int64_t N = h->N;
  auto& P = h->pend;
  P = dirb200_index::Pending();
  P.active = true;
  P.Q = Q;
  P.k = k;
  P.launches0 = launches_total();
  P.cap2 = (k > 256) ? 2048 : 1024;
  const int cap2 = P.cap2;
  if (N == 0) {   // empty shard: nothing can be selected; +inf never lowers the MIN over the shards
    std::vector<float> inf(Q, INFINITY);
    DIRB_CUDA(cudaMemcpyAsync(sel_dev, inf.data(), static_cast<size_t>(Q) * 4, cudaMemcpyHostToDevice, stream));
    DIRB_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  // ---- sizes
  int64_t S = h->sample_rows > 0 ? h->sample_rows : std::max<int64_t>(8192, ceil_div(N, 16));
  S = std::min<int64_t>(std::min<int64_t>(S, 65536), N);
  S = std::max<int64_t>(S, std::min<int64_t>(N, 4 * k));
  const bool small = (N <= S);
  if (small) S = N;
  if (!small && S / 32 < k) S = std::min<int64_t>(N, std::max<int64_t>(S, 32 * static_cast<int64_t>(k)));
  // seed threshold from group maxima (1/32 of the dense traffic) whenever there are at least k groups
  const bool use_gmax = !small && (S / 32 >= k);
  const int64_t S_ld = ceil_div(S, 128) * 128;   // dense row stride (16-byte aligned rows)
  const double expect = small ? (2.0 * k + 64) : (1.5 * k * static_cast<double>(N) / S);
  int cap = static_cast<int>(std::min<double>(std::max<double>(4096, 4 * expect), 1 << 18));
  if (h->cand_cap > 0) cap = std::max(h->cand_cap, 2 * k);
  cap = (cap + 255) / 256 * 256;
  const float band = static_cast<float>(2.0 * h->eps16);
  P.S = S;
  P.cap = cap;
  P.band = band;

  // ---- workspace carve-up
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_q16 = carve(static_cast<size_t>(Q) * D * 2);
  const size_t o_dense = carve(static_cast<size_t>(Q) * S_ld * 4);
  const size_t o_thr = carve(static_cast<size_t>(Q) * 4);
  const size_t o_cnt = carve(static_cast<size_t>(Q) * 4);
  const size_t o_cand = carve(static_cast<size_t>(Q) * cap * 8);
  const size_t o_sidx = carve(static_cast<size_t>(Q) * cap2 * 4);
  const size_t o_sscore = carve(static_cast<size_t>(Q) * cap2 * 8);
  const size_t o_scnt = carve(static_cast<size_t>(Q) * 4);
  const size_t o_flags = carve(static_cast<size_t>(Q) * 4);
  const size_t o_kthk = carve(static_cast<size_t>(Q) * 4);
  if (off > h->ws_bytes) {
    if (h->ws) DIRB_CUDA(cudaFree(h->ws));
    h->ws = nullptr;
    DIRB_CUDA(cudaMalloc(&h->ws, off));
    h->ws_bytes = off;
  }
  if (Q > h->h_flags_n) {
    if (h->h_flags) cudaFreeHost(h->h_flags);
    DIRB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h->h_flags), static_cast<size_t>(Q) * 4));
    h->h_flags_n = Q;
  }
  uint8_t* w = static_cast<uint8_t*>(h->ws);
  __half* q16 = reinterpret_cast<__half*>(w + o_q16);
  float* dense = reinterpret_cast<float*>(w + o_dense);
  float* thr = reinterpret_cast<float*>(w + o_thr);
  P.cnt = reinterpret_cast<int*>(w + o_cnt);
  P.cand = reinterpret_cast<unsigned long long*>(w + o_cand);
  P.sidx = reinterpret_cast<int*>(w + o_sidx);
  P.sscore = reinterpret_cast<double*>(w + o_sscore);
  P.scnt = reinterpret_cast<int*>(w + o_scnt);
  P.flags = reinterpret_cast<int*>(w + o_flags);
  P.kth_k = reinterpret_cast<float*>(w + o_kthk);
  P.sel = sel_dev;
  int* cnt = P.cnt;
  unsigned long long* cand = P.cand;

  mark_phase(h, stream);  // 0
  // ---- 1. queries to fp16
  DIRB_TRY(f32_to_f16(q32, static_cast<int64_t>(Q) * D, q16, stream));
  mark_phase(h, stream);  // 1: convert
  // ---- 2. seed pass over the first S rows
  {
    SimArgs a;
    a.dense = dense;
    a.dense_ld = S_ld;
    DIRB_TRY(sim_gemm(use_gmax ? PERS_EPI_SIM_GMAX : PERS_EPI_SIM_DENSE, q16, Q, h->db16, S, D, a, stream));
    mark_phase(h, stream);  // 2: seed GEMM
    const int n_vals = use_gmax ? static_cast<int>(ceil_div(S, 32)) : static_cast<int>(S);
    const int kk = static_cast<int>(std::min<int64_t>(k, n_vals));
    kth_dense_kernel<<<Q, SEL_THREADS, 0, stream>>>(dense, S_ld, n_vals, kk, band, thr);
    count_launch();
    mark_phase(h, stream);  // 3: k-th of the seed scores
  }
  DIRB_CUDA(cudaMemsetAsync(cnt, 0, static_cast<size_t>(Q) * 4, stream));
  for (;;) {
    // ---- 3. candidates
    if (small) {
      dim3 g(static_cast<unsigned>(ceil_div(N, 256)), static_cast<unsigned>(Q));
      dense_compact_kernel<<<g, 256, 0, stream>>>(dense, S_ld, static_cast<int>(N), thr, cand, cnt, cap);
      count_launch();
    } else {
      SimArgs a;
      a.thr = thr;
      a.cand = cand;
      a.cand_cnt = cnt;
      a.cand_cap = cap;
      DIRB_TRY(sim_gemm(PERS_EPI_SIM_FILTER, q16, Q, h->db16, N, D, a, stream));
    }
    if (P.retries == 0) mark_phase(h, stream);  // 4: filter pass
    // ---- 4. local k-th / k_shard-th candidate scores
    cand_kth_kernel<<<Q, SEL_THREADS, 0, stream>>>(cand, cnt, cap, k, std::min<int>(k_shard, static_cast<int>(std::min<int64_t>(k, N))), band,
                                                   P.kth_k, sel_dev, thr, P.flags, N);
    count_launch();
    if (P.retries == 0) mark_phase(h, stream);  // 5: candidate selection
    DIRB_CUDA(cudaMemcpyAsync(h->h_flags, P.flags, static_cast<size_t>(Q) * 4, cudaMemcpyDeviceToHost, stream));
    DIRB_CUDA(cudaStreamSynchronize(stream));
    bool cand_over = false;
    for (int i = 0; i < Q; ++i) cand_over |= (h->h_flags[i] & 1) != 0;
    if (!cand_over) break;
    DIRB_REQUIRE(P.retries < 4, DIRB200_EOVERFLOW, "candidate buffer overflow not resolved after %d retries", P.retries);
    ++P.retries;
    // Overflowed queries restart from an empty list with the tightened threshold that cand_kth wrote.  Finished
    // queries have thr = +inf, so the re-run appends nothing for them and the (idempotent) selection reproduces
    // their values from the unchanged list.
    for (int i = 0; i < Q; ++i)
      if (h->h_flags[i] & 1) DIRB_CUDA(cudaMemsetAsync(cnt + i, 0, 4, stream));
  }
  mark_phase(h, stream);  // 6: flags round trip
  return 0;
}

// Phase 2: survivors (candidates within the band of max(sel, local k-th)), exact re-scoring, ordered output.
int dirb200_index_search_finish(dirb200_index* h, const float* q32, const float* sel_dev, double* scores_dev,
                                int64_t* idx_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DIRB_REQUIRE(h && q32 && sel_dev && scores_dev && idx_dev, DIRB200_EINVAL, "null argument");
  auto& P = h->pend;
  DIRB_REQUIRE(P.active, DIRB200_ESTATE, "dirb200_index_search_finish without a matching _begin");
  DIRB_CUDA(cudaSetDevice(h->device));
  const int Q = P.Q, k = P.k, cap2 = P.cap2;
  P.active = false;
  if (h->N == 0) {
    std::vector<double> s(static_cast<size_t>(Q) * k, -INFINITY);
    std::vector<int64_t> ix(static_cast<size_t>(Q) * k, -1);
    DIRB_CUDA(cudaMemcpyAsync(scores_dev, s.data(), s.size() * 8, cudaMemcpyHostToDevice, stream));
    DIRB_CUDA(cudaMemcpyAsync(idx_dev, ix.data(), ix.size() * 8, cudaMemcpyHostToDevice, stream));
    DIRB_CUDA(cudaStreamSynchronize(stream));
    return 0;
  }
  cand_survivors_kernel<<<Q, SEL_THREADS, 0, stream>>>(P.cand, P.cnt, P.cap, P.kth_k, sel_dev, P.band, P.sidx, P.scnt, cap2,
                                                       P.flags);
  {
    dim3 g(static_cast<unsigned>(ceil_div(cap2, 8)), static_cast<unsigned>(Q));
    rescore_kernel<<<g, 256, 0, stream>>>(q32, h->db32, h->dim, P.sidx, P.scnt, cap2, P.sscore);
    mark_phase(h, stream);  // 7: survivors + exact re-scoring
    const size_t smem = static_cast<size_t>(cap2) * 16;
    DIRB_CUDA(cudaFuncSetAttribute(sort_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
    sort_topk_kernel<<<Q, 1024, smem, stream>>>(P.sscore, P.sidx, nullptr, P.scnt, 0, cap2, h->offset, k, scores_dev, idx_dev);
    count_launch(3);
    DIRB_CUDA(cudaGetLastError());
    mark_phase(h, stream);  // 8: sort
  }
  int64_t cand_total = 0, surv_total = 0;
  {
    std::vector<int> hc(Q), hs(Q);
    DIRB_CUDA(cudaMemcpyAsync(hc.data(), P.cnt, static_cast<size_t>(Q) * 4, cudaMemcpyDeviceToHost, stream));
    DIRB_CUDA(cudaMemcpyAsync(hs.data(), P.scnt, static_cast<size_t>(Q) * 4, cudaMemcpyDeviceToHost, stream));
    DIRB_CUDA(cudaMemcpyAsync(h->h_flags, P.flags, static_cast<size_t>(Q) * 4, cudaMemcpyDeviceToHost, stream));
    DIRB_CUDA(cudaStreamSynchronize(stream));
    bool surv_over = false;
    for (int i = 0; i < Q; ++i) {
      cand_total += hc[i];
      surv_total += hs[i];
      surv_over |= (h->h_flags[i] & 2) != 0;
    }
    DIRB_REQUIRE(!surv_over, DIRB200_EOVERFLOW,
                 "more than %d rows within 2*eps16=%g of the k-th score for some query (near-duplicate rows?)", cap2,
                 (double)P.band);
  }
  if (h->profile && P.mark_i == 9) {
    for (int i = 0; i < 8; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
      h->phase_ms[i] = ms;
    }
  }
  h->stats[0] = P.S;
  h->stats[1] = cand_total;
  h->stats[2] = surv_total;
  h->stats[3] = P.retries;
  h->stats[4] = launches_total() - P.launches0;
  return 0;
}

// Single-shard search = phase 1 with k_shard = k followed directly by phase 2.
int dirb200_index_search(dirb200_index* h, const float* q32, int Q, int k, double* scores_dev, int64_t* idx_dev,
                         void* stream_) {
  DIRB_REQUIRE(h && q32 && scores_dev && idx_dev, DIRB200_EINVAL, "null argument");
  DIRB_REQUIRE(Q > 0, DIRB200_ENOTSUP, "need Q > 0");
  if (static_cast<size_t>(Q) * 4 > h->sel_bytes) {
    if (h->sel_own) cudaFree(h->sel_own);
    h->sel_own = nullptr;
    DIRB_CUDA(cudaMalloc(reinterpret_cast<void**>(&h->sel_own), static_cast<size_t>(Q) * 4));
    h->sel_bytes = static_cast<size_t>(Q) * 4;
  }
  DIRB_TRY(dirb200_index_search_begin(h, q32, Q, k, k, h->sel_own, stream_));
  return dirb200_index_search_finish(h, q32, h->sel_own, scores_dev, idx_dev, stream_);
}

int dirb200_topk_merge(const double* scores_dev, const int64_t* idx_dev, int G, int Q, int k, int64_t shard_stride,
                       double* out_scores_dev, int64_t* out_idx_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DIRB_REQUIRE(scores_dev && idx_dev && out_scores_dev && out_idx_dev, DIRB200_EINVAL, "null argument");
  DIRB_REQUIRE(G >= 1 && Q >= 1 && k >= 1 && static_cast<int64_t>(G) * k <= 4096, DIRB200_ENOTSUP,
               "merge supports G*k <= 4096 (got G=%d k=%d)", G, k);
  // shard g holds [Q][k] at element offset g*shard_stride; the sort kernel gathers the G lists of a query in place
  if (shard_stride <= 0) shard_stride = static_cast<int64_t>(Q) * k;
  const int n = G * k;
  int P = 2;
  while (P < n) P <<= 1;
  DIRB_CUDA(cudaFuncSetAttribute(sort_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
  sort_topk_kernel<<<Q, 1024, static_cast<size_t>(P) * 16, stream>>>(scores_dev, nullptr, idx_dev, nullptr, n, n, 0, k,
                                                                    out_scores_dev, out_idx_dev, k, shard_stride);
  count_launch();
  DIRB_CUDA(cudaGetLastError());
  return 0;
}

int dirb200_scores_exact(const float* q_dev, int Q, const float* db_dev, int64_t N, int D, float* out_dev,
                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DIRB_REQUIRE(q_dev && db_dev && out_dev && D % 4 == 0, DIRB200_EINVAL, "bad arguments");
  if (Q == 0 || N == 0) return 0;
  dim3 g(static_cast<unsigned>(ceil_div(N, 8)), static_cast<unsigned>(ceil_div(Q, 4)));
  scores_exact_kernel<<<g, 256, 0, stream>>>(q_dev, Q, db_dev, N, D, out_dev);
  count_launch();
  DIRB_CUDA(cudaGetLastError());
  return 0;
}

int dirb200_aqe_expand(const float* q_dev, int Q, int D, const float* db32_dev, const int64_t* nn_idx_dev,
                       const double* nn_scores_dev, int k, double alpha, int partial, int64_t row_offset, int64_t n_rows,
                       float* out_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DIRB_REQUIRE(q_dev && nn_idx_dev && nn_scores_dev && out_dev && (db32_dev || (partial && n_rows == 0)), DIRB200_EINVAL,
               "null argument");
  DIRB_REQUIRE(k >= 1 && alpha >= 0, DIRB200_EINVAL, "k and alpha must be non-negative (test_dir.py:25)");
  DIRB_REQUIRE(k <= 2048, DIRB200_ENOTSUP, "at most 2048 neighbours per query (got %d)", k);
  if (Q == 0) return 0;
  const size_t smem = static_cast<size_t>((k + 1) & ~1) * 4 + static_cast<size_t>(k) * 8;
  aqe_kernel<<<Q, 256, smem, stream>>>(q_dev, D, db32_dev, nn_idx_dev, nn_scores_dev, k, alpha, partial, row_offset,
                                       n_rows, out_dev);
  count_launch();
  DIRB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
