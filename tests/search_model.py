"""NumPy model of the exact top-k search protocol of deep-image-retrieval_b200/csrc/search.cu + dist.py.
TEST INFRASTRUCTURE: it restates the ALGORITHM (thresholds, bands, the cross-shard MIN exchange, merge), not the
kernels, so that its exactness can be checked on the CPU over thousands of random and adversarial cases
(tests/test_properties.py).  Scores on the "fast" path are emulated as the kernels compute them: operands rounded to
fp16, products accumulated in fp32."""
import numpy as np

EPS16 = 1.2e-3
BAND = np.float32(2.0 * EPS16)


def fast_scores(q, db):
    q16, d16 = q.astype(np.float16).astype(np.float32), db.astype(np.float16).astype(np.float32)
    out = np.zeros((q.shape[0], db.shape[0]), np.float32)
    for c in range(0, q.shape[1], 16):
        out += q16[:, c:c + 16] @ d16[:, c:c + 16].T
    return out


def exact_scores(q, db):
    # one fp64 reduction per (query, row) pair whose order does not depend on which other rows are in the call
    # (BLAS blocking would make the bits of a pair's score depend on the batch, i.e. break exact ties differently)
    return (q.astype(np.float64)[:, None, :] * db.astype(np.float64)[None, :, :]).sum(axis=-1)


def kth_largest(v, k):
    return -np.inf if v.shape[0] < k else np.partition(v, v.shape[0] - k)[v.shape[0] - k]


class Overflow(Exception):
    """DIRB200_EOVERFLOW: candidate / survivor buffers could not hold the rows inside the band."""


class ShardModel:
    """One row shard: dirb200_index_search_begin / _finish.  cand_cap / surv_cap model the fixed-size candidate and
    survivor buffers: an overflowing candidate list keeps an ARBITRARY subset of `cand_cap` entries (the kernels
    append with atomics), its k-th best gives a tighter threshold and the pass is re-run (at most 4 times)."""

    def __init__(self, db, offset, sample_rows, cand_cap=None, surv_cap=None, rng=None):
        self.db, self.offset, self.sample_rows = db, int(offset), int(sample_rows)
        self.cand_cap, self.surv_cap, self.rng = cand_cap, surv_cap, rng or np.random.RandomState(0)
        self.retries = 0

    def begin(self, q, k, k_shard):
        """dirb200_index_search_begin: seed half + filter half with the purely local filter threshold."""
        self.begin_seed(q, k, k_shard)
        return self.begin_filter(None)

    def begin_seed(self, q, k, k_shard):
        """Seed half (search.cu: search_begin_seed).  Returns this shard's k_shard-th seed bound per query (minus the
        band; +inf for an empty shard) - what the sharded search over peer memory MIN-reduces before the filter pass."""
        n = self.db.shape[0]
        self.q, self.k, self.k_shard = q, k, k_shard
        if n == 0:
            self.cand = None
            return np.full(q.shape[0], np.inf, np.float32)
        self.fast = fast_scores(q, self.db)
        s = min(max(self.sample_rows, min(n, 4 * k)), n)          # seed rows (search.cu: S)
        small = n <= s
        if not small and s // 32 < k:
            s = min(n, max(s, 32 * k))
        use_gmax = (not small) and (s // 32 >= k)
        thr = np.empty(q.shape[0], np.float32)
        seed2 = np.empty(q.shape[0], np.float32)
        for i in range(q.shape[0]):
            seed = self.fast[i, :n if small else s]
            if use_gmax:                                          # maxima of groups of 32 consecutive rows
                pad = (-seed.shape[0]) % 32
                vals = np.concatenate([seed, np.full(pad, -np.inf, np.float32)]).reshape(-1, 32).max(axis=1)
            else:
                vals = seed
            thr[i] = kth_largest(vals, min(k, vals.shape[0])) - BAND
            seed2[i] = kth_largest(vals, min(k_shard, vals.shape[0])) - BAND     # kth_dense_kernel: k2 = min(k_shard, n_vals)
        self.thr = thr
        return seed2

    def begin_filter(self, seed_min):
        """Filter half (search_begin_filter).  seed_min: MIN over the shards of their begin_seed values (thr_min_kernel:
        filter threshold = max(local k-th seed bound, that minimum)), or None for the stand-alone / collective protocol."""
        q, k, k_shard = self.q, self.k, self.k_shard
        n = self.db.shape[0]
        if n == 0:
            return np.full(q.shape[0], np.inf, np.float32)
        thr = self.thr if seed_min is None else np.maximum(self.thr, seed_min.astype(np.float32))
        thr_in = thr.copy()
        self.cand = [np.nonzero(self.fast[i] >= thr[i])[0] for i in range(q.shape[0])]
        kk = min(k, n)
        if self.cand_cap is not None:
            cap = max(self.cand_cap, 2 * k)
            for _ in range(5):
                over = [i for i, c in enumerate(self.cand) if c.shape[0] > cap]
                if not over:
                    break
                if self.retries == 4:
                    raise Overflow("candidates")
                self.retries += 1
                for i in over:                                    # what was captured: any `cap` of the matching rows
                    kept = self.rng.choice(self.cand[i], cap, replace=False)
                    thr[i] = kth_largest(self.fast[i, kept], kk) - BAND
                    thr_in[i] = thr[i]
                    self.cand[i] = np.nonzero(self.fast[i] >= thr[i])[0]
        self.kth_k = np.array([kth_largest(self.fast[i, c], kk) for i, c in enumerate(self.cand)], np.float32)
        ks = min(k_shard, kk)
        sel = np.array([kth_largest(self.fast[i, c], ks) for i, c in enumerate(self.cand)], np.float32)
        # cand_kth_kernel: fewer than ks rows above a threshold that came from the other shards -> report that bound
        # (a valid lower bound on the global k-th best) instead of -inf
        clamp = np.isneginf(sel) & ~np.isneginf(thr_in)
        sel[clamp] = thr_in[clamp] + BAND
        return sel

    def finish(self, sel):
        k, nq = self.k, self.q.shape[0]
        scores = np.full((nq, k), -np.inf)
        idx = np.full((nq, k), -1, np.int64)
        self.survivors = 0
        if self.cand is None:
            return scores, idx
        for i in range(nq):
            t2 = np.float32(max(sel[i], self.kth_k[i])) - BAND
            rows = self.cand[i][self.fast[i, self.cand[i]] >= t2]
            if self.surv_cap is not None and rows.shape[0] > self.surv_cap:
                raise Overflow("survivors")
            self.survivors += rows.shape[0]
            ex = exact_scores(self.q[i:i + 1], self.db[rows])[0]
            order = np.lexsort((rows, -ex))[:k]
            scores[i, :order.shape[0]] = ex[order]
            idx[i, :order.shape[0]] = rows[order] + self.offset
        return scores, idx


def sharded_search(q, db, k, bounds, sample_rows=64, quota=None, cand_cap=None, surv_cap=None, seed=0):
    """dist.ShardedIndex.search over the row ranges `bounds` = [(start, end), ...]."""
    shards = [ShardModel(db[a:b], a, sample_rows, cand_cap, surv_cap, np.random.RandomState(seed + j))
              for j, (a, b) in enumerate(bounds)]
    from dirb200.dist import shard_quota                         # the product's own rule for the selection depth
    k_shard = shard_quota(k, [b - a for a, b in bounds]) if quota is None else quota
    sels = [s.begin(q, k, k_shard) for s in shards]
    sel = np.minimum.reduce(sels)                                 # all_reduce(MIN)
    lists = [s.finish(sel) for s in shards]
    sc = np.concatenate([l[0] for l in lists], axis=1)            # all_gather + merge
    ix = np.concatenate([l[1] for l in lists], axis=1)
    out_s = np.full((q.shape[0], k), -np.inf)
    out_i = np.full((q.shape[0], k), -1, np.int64)
    for i in range(q.shape[0]):
        valid = ix[i] >= 0
        order = np.lexsort((ix[i][valid], -sc[i][valid]))[:k]
        out_s[i, :order.shape[0]] = sc[i][valid][order]
        out_i[i, :order.shape[0]] = ix[i][valid][order]
    return out_s, out_i, sum(s.survivors for s in shards)


def sharded_search_peer(q, db, k, bounds, sample_rows=64, cand_cap=None, surv_cap=None, seed=0):
    """dirb200_index_search_sharded (four phases over peer memory): MIN of the shards' k_shard-th seed bounds before the
    filter pass, MIN of their selection thresholds before the re-scoring, gather + merge of the lists.
    Returns (scores, indices, survivors re-scored, candidates captured)."""
    shards = [ShardModel(db[a:b], a, sample_rows, cand_cap, surv_cap, np.random.RandomState(seed + j))
              for j, (a, b) in enumerate(bounds)]
    from dirb200.dist import shard_quota
    k_shard = shard_quota(k, [b - a for a, b in bounds])
    seed_min = np.minimum.reduce([s.begin_seed(q, k, k_shard) for s in shards])     # tab_push (seed) + thr_min_kernel
    sel = np.minimum.reduce([s.begin_filter(seed_min) for s in shards])             # tab_push (sel) + finish prologue
    lists = [s.finish(sel) for s in shards]                                         # finish epilogue: lists to every peer
    sc = np.concatenate([l[0] for l in lists], axis=1)                              # merge_lists_kernel
    ix = np.concatenate([l[1] for l in lists], axis=1)
    out_s = np.full((q.shape[0], k), -np.inf)
    out_i = np.full((q.shape[0], k), -1, np.int64)
    for i in range(q.shape[0]):
        valid = ix[i] >= 0
        order = np.lexsort((ix[i][valid], -sc[i][valid]))[:k]
        out_s[i, :order.shape[0]] = sc[i][valid][order]
        out_i[i, :order.shape[0]] = ix[i][valid][order]
    cands = sum(sum(c.shape[0] for c in s.cand) for s in shards if s.cand is not None)
    return out_s, out_i, sum(s.survivors for s in shards), cands


def exact_topk(q, db, k):
    ex = exact_scores(q, db)
    out_s = np.full((q.shape[0], k), -np.inf)
    out_i = np.full((q.shape[0], k), -1, np.int64)
    for i in range(q.shape[0]):
        order = np.lexsort((np.arange(db.shape[0]), -ex[i]))[:k]
        out_s[i, :order.shape[0]] = ex[i][order]
        out_i[i, :order.shape[0]] = order
    return out_s, out_i
