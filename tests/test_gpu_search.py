"""Parity of similarity + top-k, merge, exact scores and alpha-QE against the CPU oracle.  -m gpu."""
import numpy as np
import pytest
import torch

import synthdata as synth
from oracle import dir_oracle as O
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ops():
    from dirb200 import ops
    ops.require_gpu(0)
    return ops


def _check_topk(ops, n_db, n_q, dim, k, n_pos=10, sample_rows=0, seed=0, cand_cap=0):
    db, q, pos = synth.make_descriptor_db(n_db, n_q, dim=dim, n_pos=n_pos, db_seed=100 + seed, q_seed=200 + seed)
    index = ops.Index(torch.from_numpy(db).to(DEV), index_offset=0)
    if sample_rows:
        index.set_option("sample_rows", sample_rows)
    if cand_cap:
        index.set_option("cand_cap", cand_cap)
        index.set_option("retries", 3)             # a deliberately tiny buffer needs more than the default single retry pass
    s, i = index.search(torch.from_numpy(q).to(DEV), k)
    torch.cuda.synchronize()
    rs, ri = O.topk(q, db, k)
    kk = min(k, n_db)
    np.testing.assert_array_equal(i.cpu().numpy()[:, :kk], ri)           # indices bit-exact
    np.testing.assert_allclose(s.cpu().numpy()[:, :kk], rs, rtol=0, atol=1e-12)
    if k > n_db:
        assert (i.cpu().numpy()[:, n_db:] == -1).all()
    return index.stats(), pos, i.cpu().numpy()


@pytest.mark.parametrize("n_db,n_q,dim,k", [(100, 4, 2048, 10), (100, 4, 2048, 100), (50, 3, 128, 100),
                                             (1000, 70, 2048, 100), (5000, 130, 256, 20), (20000, 7, 2048, 1)])
def test_topk_small_db(n_db, n_q, dim, k):
    _check_topk(_ops(), n_db, n_q, dim, k)


def test_topk_filtered_pass():
    # N > sample rows -> seed pass + filtered tensor-core pass + exact rescoring
    st, pos, idx = _check_topk(_ops(), 30000, 70, 2048, 100, sample_rows=4096)
    assert st["dense_rows"] == 4096 and st["candidates"] > 0 and st["retries"] == 0
    for qi in range(pos.shape[0]):                                        # every planted positive is retrieved
        assert set(pos[qi]) <= set(idx[qi])


def test_topk_overflow_retry():
    # a tiny seed sample makes a loose threshold -> candidate overflow -> tightened re-run must stay exact
    st, _, _ = _check_topk(_ops(), 60000, 5, 256, 50, sample_rows=256, seed=3, cand_cap=512)
    assert st["retries"] >= 1


def test_topk_ties_lowest_index_first():
    ops = _ops()
    db, q, _ = synth.make_descriptor_db(2000, 3, dim=128, n_pos=0)
    db[100] = db[7]
    db[1500] = db[7]                                                      # three identical rows
    q[0] = db[7]
    s, i = ops.Index(torch.from_numpy(db).to(DEV)).search(torch.from_numpy(q).to(DEV), 5)
    assert i[0, :3].cpu().tolist() == [7, 100, 1500]
    rs, ri = O.topk(q, db, 5)
    np.testing.assert_array_equal(i.cpu().numpy(), ri)


def test_shard_merge_matches_single_index():
    ops = _ops()
    db, q, _ = synth.make_descriptor_db(12000, 33, dim=512, n_pos=6)
    k, G = 50, 4
    qd = torch.from_numpy(q).to(DEV)
    parts = np.array_split(np.arange(12000), G)
    ss, ii = [], []
    for part in parts:
        idx = ops.Index(torch.from_numpy(db[part[0]:part[-1] + 1]).to(DEV), index_offset=int(part[0]))
        s, i = idx.search(qd, k)
        ss.append(s)
        ii.append(i)
    ms, mi = ops.topk_merge(torch.stack(ss).contiguous(), torch.stack(ii).contiguous(), k)
    rs, ri = O.topk(q, db, k)
    np.testing.assert_array_equal(mi.cpu().numpy(), ri)
    np.testing.assert_allclose(ms.cpu().numpy(), rs, rtol=0, atol=1e-12)
    os_, oi = O.merge_topk([s.cpu().numpy() for s in ss], [i.cpu().numpy() for i in ii], k)
    np.testing.assert_array_equal(mi.cpu().numpy(), oi)


def test_scores_exact_and_map(golden):
    ops = _ops()
    g = golden("rank_ap.npz")
    db, q, pos = synth.make_descriptor_db(int(g["n_db"]), int(g["n_q"]), dim=int(g["dim"]), n_pos=int(g["n_pos"]),
                                          db_seed=int(g["db_seed"]), q_seed=int(g["q_seed"]))
    sc = ops.scores_exact(torch.from_numpy(q).to(DEV), torch.from_numpy(db).to(DEV)).cpu().numpy()
    np.testing.assert_allclose(sc, g["scores"], rtol=0, atol=2e-6)
    gnd = synth.oxford_gt(pos, n_junk=int(g["n_junk"]), n_db=int(g["n_db"]), seed=int(g["gt_seed"]))
    for i in range(sc.shape[0]):
        assert abs(O.eval_query_ap(sc[i], gnd[i]["ok"], gnd[i]["junk"]) - g["aps"][i]) < 1e-12


def test_matmul_any_dimension():
    """common.matmul / scores_exact with D not a multiple of 4 (e.g. --whitenv 127): np.dot accepts any D."""
    ops = _ops()
    from dirtorch.utils import common
    r = np.random.RandomState(3)
    for d in (127, 30, 1):
        q = r.standard_normal((5, d)).astype(np.float32)
        db = r.standard_normal((33, d)).astype(np.float32)
        ref = q.astype(np.float64) @ db.astype(np.float64).T
        np.testing.assert_allclose(common.matmul(q, db), ref, rtol=0, atol=1e-5)
        np.testing.assert_allclose(ops.scores_exact(torch.from_numpy(q).to(DEV), torch.from_numpy(db).to(DEV)).cpu().numpy(),
                                   ref, rtol=0, atol=1e-5)


def test_aqe(golden):
    ops = _ops()
    g = golden("aqe.npz")
    db, q, pos = synth.make_descriptor_db(int(g["n_db"]), int(g["n_q"]), dim=int(g["dim"]), n_pos=int(g["n_pos"]),
                                          db_seed=int(g["db_seed"]), q_seed=int(g["q_seed"]))
    dbd, qd = torch.from_numpy(db).to(DEV), torch.from_numpy(q).to(DEV)
    for k, alpha, key in ((2, 0.5, "aqe_k2_a05"), (3, 1.0, "aqe_k3_a1")):
        rs, ri = O.topk(q, db, k)
        out = ops.aqe_expand(qd, dbd, torch.from_numpy(ri).to(DEV), torch.from_numpy(rs).to(DEV), alpha)
        assert rel_l2(out.cpu().numpy(), g[key]) < 1e-5
        assert rel_l2(out.cpu().numpy(), O.expand_descriptors(q, db=db, k=k, alpha=alpha)) < 1e-5


def test_full_size_1m_properties():
    """BASELINE configs[3] geometry on one GPU: 1000 queries x 1M x 2048 (generated on the device).  Checked through
    size-independent properties + a sampled comparison with the CPU oracle."""
    ops = _ops()
    N, Q, D, K = 1_000_000, 1000, 2048, 100
    g = torch.Generator(device="cuda").manual_seed(5)
    db, db16 = ops.l2_normalize(torch.randn((N, D), generator=g, device="cuda"), want_f16=True)
    q = ops.l2_normalize(torch.randn((Q, D), generator=g, device="cuda"))
    # plant: query i is close to database row 1000*i + 7
    rows = torch.arange(Q, device="cuda") * 1000 + 7
    q = ops.l2_normalize((db[rows] + 0.5 * q).contiguous())
    index = ops.Index(db, db16=db16)
    s, i = index.search(q, K)
    torch.cuda.synchronize()
    s_h, i_h = s.cpu().numpy(), i.cpu().numpy()
    assert (i_h[:, 0] == rows.cpu().numpy()).all()                        # planted neighbour is rank 0
    assert (np.diff(s_h, axis=1) <= 0).all()                              # sorted, descending
    assert all(len(set(r)) == K for r in i_h[::50])                       # no duplicates
    s2, i2 = index.search(q, K)
    assert torch.equal(i, i2) and torch.equal(s, s2)                      # deterministic despite atomics
    # idempotence of the shard merge: merging the result with itself changes nothing
    ms, mi = ops.topk_merge(s.unsqueeze(0).contiguous(), i.unsqueeze(0).contiguous(), K)
    assert torch.equal(mi, i)
    # sampled oracle check (fp64 scores of 8 queries against all rows, chunked)
    pick = [0, 1, 137, 500, 501, 777, 998, 999]
    qs = q[pick].cpu().numpy().astype(np.float64)
    best_s = np.full((len(pick), K), -np.inf)
    best_i = np.zeros((len(pick), K), dtype=np.int64)
    for c0 in range(0, N, 125_000):
        blk = db[c0:c0 + 125_000].cpu().numpy().astype(np.float64)
        sc = qs @ blk.T
        cat_s = np.concatenate([best_s, sc], axis=1)
        cat_i = np.concatenate([best_i, np.broadcast_to(np.arange(c0, c0 + blk.shape[0]), sc.shape)], axis=1)
        for r in range(len(pick)):
            o = np.lexsort((cat_i[r], -cat_s[r]))[:K]
            best_s[r], best_i[r] = cat_s[r][o], cat_i[r][o]
    np.testing.assert_array_equal(i_h[pick], best_i)
    np.testing.assert_allclose(s_h[pick], best_s, rtol=0, atol=1e-12)
    st = index.stats()
    assert st["retries"] == 0 and st["candidates"] < 40 * Q * K


def test_config4_geometry_sharded_protocol_single_gpu():
    """BASELINE configs[3]/[4] geometry with the 8 shards on ONE GPU (runs in the 1-GPU driver pass): 1000 queries x
    (8 x 125 000) x 2048, k = 100.  The sharded protocol (phase 1 per shard with k_shard = shard_quota, MIN of the
    thresholds = the all-reduce, phase 2 per shard, merge of the 8 lists = the all-gather + merge) must reproduce the
    single-index search bit for bit (indices AND fp64 scores), and alpha-QE (k=2, alpha=0.5, test_dir.py:24-44) from
    per-shard partial sums must equal the unsharded expansion; 8 queries are checked against the CPU fp64 oracle."""
    ops = _ops()
    from dirb200.dist import shard_quota, shard_rows
    N, Q, D, K, G = 1_000_000, 1000, 2048, 100, 8
    g = torch.Generator(device="cuda").manual_seed(11)
    db, db16 = ops.l2_normalize(torch.randn((N, D), generator=g, device="cuda"), want_f16=True)
    q = ops.l2_normalize(torch.randn((Q, D), generator=g, device="cuda"))
    rows = torch.arange(Q, device="cuda") * 997 + 3                        # plant a close neighbour per query
    q = ops.l2_normalize((db[rows] + 0.7 * q).contiguous())
    whole = ops.Index(db, db16=db16)
    s_ref, i_ref = whole.search(q, K)
    bounds = [shard_rows(N, G, r) for r in range(G)]
    shards = [ops.Index(db[a:b], index_offset=a, db16=db16[a:b]) for a, b in bounds]
    c = shard_quota(K, [b - a for a, b in bounds])
    assert c == 13
    for sh in shards:
        sh.set_option("deferred_check", 1)
    sels = [sh.search_begin(q, K, c) for sh in shards]
    sel = torch.stack(sels).min(dim=0).values.contiguous()
    packed = torch.empty((G, 2, Q, K), dtype=torch.int64, device="cuda")
    for j, sh in enumerate(shards):
        sh.search_finish(q, K, sel, out=packed[j])
    ms, mi = ops.topk_merge_packed(packed, K)
    for sh in shards:
        sh.check()
    assert torch.equal(mi, i_ref) and torch.equal(ms, s_ref)
    surv = sum(sh.stats()["survivors"] for sh in shards)
    cand_nccl = sum(sh.stats()["candidates"] for sh in shards)
    assert surv < 3 * Q * K, surv                                          # ~1.4 k rows per query over ALL shards
    assert all(sh.stats()["retries"] == 0 for sh in shards)
    # alpha-QE from per-shard partial sums (what the all-reduce adds up) == unsharded expansion
    s2, i2 = ms[:, :2].contiguous(), mi[:, :2].contiguous()
    full = ops.aqe_expand(q, db, i2, s2, 0.5)
    part = torch.zeros_like(q)
    for (a, b), sh in zip(bounds, shards):
        part += ops.aqe_expand(q, db[a:b], i2, s2, 0.5, partial=True, row_offset=a, n_rows=b - a)
    out = ops.pool_scales([part, q], "mean", l2=True)
    assert rel_l2(out.cpu().numpy(), full.cpu().numpy()) < 1e-6
    # second search with the expanded queries through the sharded protocol == through the single index
    s3, i3 = whole.search(out, K)
    sels = [sh.search_begin(out, K, c) for sh in shards]
    sel = torch.stack(sels).min(dim=0).values.contiguous()
    for j, sh in enumerate(shards):
        sh.search_finish(out, K, sel, out=packed[j])
    ms3, mi3 = ops.topk_merge_packed(packed, K)
    assert torch.equal(mi3, i3) and torch.equal(ms3, s3)
    # the same protocol through the peer-memory exchange (dirb200_index_search_sharded_phase): thresholds and lists are
    # stored by the kernels into all 8 windows, the MIN and the gather happen inside search_finish / merge_lists; every
    # "rank" ends up with the global result.  Two searches: the double-buffered slots and the epoch counter advance.
    xs = [ops.Exchange(0, G, r, Q, K) for r in range(G)]
    ops.Exchange.open_local(xs)
    for qq, (s_want, i_want) in ((q, (s_ref, i_ref)), (out, (s3, i3)), (q, (s_ref, i_ref))):
        for ph in (1, 2, 3):
            for sh, x in zip(shards, xs):
                sh.search_sharded(x, qq, K, c, phase=ph)
        res = [sh.search_sharded(x, qq, K, c, phase=4) for sh, x in zip(shards, xs)]
        for sh in shards:
            sh.check()
        for ps, pi in res:
            assert torch.equal(pi, i_want) and torch.equal(ps, s_want)
        # the filter threshold of the peer path is the MIN of the shards' 13th seed bounds, not the local 100th: an
        # order of magnitude fewer candidates than the all-reduce path captured above
        assert sum(sh.stats()["candidates"] for sh in shards) < 0.3 * cand_nccl
    for x in xs:
        x.close()
    # sampled oracle check (fp64 scores of 8 queries against all rows, chunked)
    pick = [0, 1, 137, 500, 501, 777, 998, 999]
    qs = q[pick].cpu().numpy().astype(np.float64)
    best_s = np.full((len(pick), K), -np.inf)
    best_i = np.zeros((len(pick), K), dtype=np.int64)
    for c0 in range(0, N, 125_000):
        blk = db[c0:c0 + 125_000].cpu().numpy().astype(np.float64)
        sc = qs @ blk.T
        cat_s = np.concatenate([best_s, sc], axis=1)
        cat_i = np.concatenate([best_i, np.broadcast_to(np.arange(c0, c0 + blk.shape[0]), sc.shape)], axis=1)
        for r in range(len(pick)):
            o = np.lexsort((cat_i[r], -cat_s[r]))[:K]
            best_s[r], best_i[r] = cat_s[r][o], cat_i[r][o]
    np.testing.assert_array_equal(mi.cpu().numpy()[pick], best_i)
    np.testing.assert_allclose(ms.cpu().numpy()[pick], best_s, rtol=0, atol=1e-12)
    # alpha-QE of those queries against the oracle formula: normalize(q + sum_j db_j * s_j^alpha)
    qe = q[pick].cpu().numpy().astype(np.float64)
    for r in range(len(pick)):
        for j in range(2):
            qe[r] += db[int(best_i[r, j])].cpu().numpy().astype(np.float64) * best_s[r, j] ** 0.5
    qe /= np.linalg.norm(qe, axis=1, keepdims=True)
    assert rel_l2(out[pick].cpu().numpy(), qe) < 1e-5


@pytest.mark.parametrize("sizes", [(3000, 3000), (5000, 0, 700), (40, 9000, 2500, 1)], ids=lambda z: "-".join(map(str, z)))
def test_peer_exchange_uneven_and_empty_shards(sizes):
    """Sharded search over the peer-memory exchange with uneven, tiny and EMPTY shards (an empty shard still publishes its
    thresholds and lists - the peers wait for every rank's flag), k larger than some shards, against the CPU oracle; the
    whole-search entry point on a world of one."""
    ops = _ops()
    from dirb200.dist import shard_quota
    n, k, nq = sum(sizes), 50, 37
    db, q, _ = synth.make_descriptor_db(n, nq, dim=256, n_pos=4)
    rs, ri = O.topk(q, db, k)
    dbt, qt = torch.from_numpy(db).to(DEV), torch.from_numpy(q).to(DEV)
    bounds = np.concatenate([[0], np.cumsum(sizes)])
    shards = [ops.Index(dbt[a:b].contiguous(), index_offset=int(a)) for a, b in zip(bounds[:-1], bounds[1:])]
    for sh in shards:
        sh.set_option("deferred_check", 1)
        sh.set_option("sample_rows", 512)
    c = shard_quota(k, sizes)
    G = len(sizes)
    xs = [ops.Exchange(0, G, r, 64, 64) for r in range(G)]
    ops.Exchange.open_local(xs)
    for _ in range(3):
        for ph in (1, 2, 3):
            for sh, x in zip(shards, xs):
                sh.search_sharded(x, qt, k, c, phase=ph)
        res = [sh.search_sharded(x, qt, k, c, phase=4) for sh, x in zip(shards, xs)]
        for sh in shards:
            sh.check()
        for ps, pi in res:
            assert np.array_equal(pi.cpu().numpy(), ri)
            assert np.abs(ps.cpu().numpy() - rs).max() < 1e-12
    # world of one: the whole search in one call == the plain search
    one = ops.Index(dbt)
    x1 = ops.Exchange(0, 1, 0, 64, 64)
    ops.Exchange.open_local([x1])
    one.set_option("deferred_check", 1)
    for _ in range(2):
        ps, pi = one.search_sharded(x1, qt, k, k)
        one.check()
        assert np.array_equal(pi.cpu().numpy(), ri) and np.abs(ps.cpu().numpy() - rs).max() < 1e-12
    with pytest.raises(Exception):
        one.search_sharded(x1, qt, 65, 65)                  # larger than the window


def test_deferred_check_and_unresolved_overflow():
    """Option deferred_check: the search enqueues everything and returns; check() collects the status.  With the retry
    passes disabled an overflowing candidate list must surface as DIRB200_EOVERFLOW - never as a silent wrong list."""
    ops = _ops()
    from dirb200.lib import DirbError
    db, q, _ = synth.make_descriptor_db(60000, 5, dim=256, n_pos=10, db_seed=103, q_seed=203)
    rs, ri = O.topk(q, db, 50)
    qd = torch.from_numpy(q).to(DEV)
    index = ops.Index(torch.from_numpy(db).to(DEV))
    index.set_option("deferred_check", 1)
    s, i = index.search(qd, 50)
    index.check()
    np.testing.assert_array_equal(i.cpu().numpy(), ri)
    # loose seed threshold + tiny candidate buffer: resolved by the gated retry passes ...
    index.set_option("sample_rows", 256)
    index.set_option("cand_cap", 512)
    index.set_option("retries", 3)
    s, i = index.search(qd, 50)
    index.check()
    assert index.stats()["retries"] >= 1
    np.testing.assert_array_equal(i.cpu().numpy(), ri)
    np.testing.assert_allclose(s.cpu().numpy(), rs, rtol=0, atol=1e-12)
    # ... and reported when there are none
    index.set_option("retries", 0)
    index.search(qd, 50)
    with pytest.raises(DirbError) as e:
        index.check()
    assert e.value.status == -4
    index.check()                                   # the error is reported once
    index.set_option("deferred_check", 0)
    s, i = index.search(qd, 50)                     # immediate check: the Python front-end catches the overflow and re-runs
    np.testing.assert_array_equal(i.cpu().numpy(), ri)   # once with more retry passes and a larger candidate buffer
    np.testing.assert_allclose(s.cpu().numpy(), rs, rtol=0, atol=1e-12)
    from dirb200 import lib
    import ctypes as C
    sc = torch.empty((5, 50), dtype=torch.float64, device=DEV)
    ix = torch.empty((5, 50), dtype=torch.int64, device=DEV)
    index.set_option("sample_rows", 256)
    index.set_option("cand_cap", 512)
    index.set_option("retries", 0)
    with pytest.raises(DirbError):                  # the C entry point itself reports it
        lib.call("dirb200_index_search", index._h, C.c_void_p(qd.data_ptr()), 5, 50, C.c_void_p(sc.data_ptr()),
                 C.c_void_p(ix.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))


def test_empty_shard():
    """A shard without rows (more ranks than images): +inf thresholds, an all-(-inf, -1) list, no neighbour sums."""
    ops = _ops()
    db, q, _ = synth.make_descriptor_db(3000, 6, dim=128, n_pos=4)
    qd = torch.from_numpy(q).to(DEV)
    empty = ops.Index(torch.empty((0, 128), dtype=torch.float32, device=DEV), index_offset=3000)
    rest = ops.Index(torch.from_numpy(db).to(DEV), index_offset=0)
    k = 20
    sels = [sh.search_begin(qd, k, k) for sh in (empty, rest)]
    assert bool(torch.isinf(sels[0]).all())
    sel = torch.minimum(sels[0], sels[1]).contiguous()
    outs = [sh.search_finish(qd, k, sel) for sh in (empty, rest)]
    assert bool((outs[0][1] == -1).all()) and bool(torch.isinf(outs[0][0]).all())
    ms, mi = ops.topk_merge(torch.stack([o[0] for o in outs]).contiguous(), torch.stack([o[1] for o in outs]).contiguous(), k)
    rs, ri = O.topk(q, db, k)
    np.testing.assert_array_equal(mi.cpu().numpy(), ri)
    part = ops.aqe_expand(qd, empty.db32, mi[:, :2].contiguous(), ms[:, :2].contiguous(), 0.5, partial=True,
                          row_offset=3000, n_rows=0)
    assert float(part.abs().max()) == 0.0


def test_two_phase_sharded_search_single_gpu():
    """The sharded protocol with all shards on one GPU: phase 1 per shard (k_shard = ceil(k/G)), MIN over the shards'
    selection thresholds (what the all-reduce does), phase 2 per shard, merge == oracle.  Shards re-score far fewer
    rows than a stand-alone search would."""
    ops = _ops()
    db, q, pos = synth.make_descriptor_db(48000, 40, dim=512, n_pos=8, db_seed=77, q_seed=78)
    k, G = 60, 4
    qd = torch.from_numpy(q).to(DEV)
    bounds = [(g * 12000, (g + 1) * 12000) for g in range(G)]
    shards = [ops.Index(torch.from_numpy(db[a:b]).to(DEV), index_offset=a) for a, b in bounds]
    for sh in shards:
        sh.set_option("sample_rows", 2048)
    sels = [sh.search_begin(qd, k, -(-k // G)) for sh in shards]
    sel = torch.stack(sels).min(dim=0).values.contiguous()
    outs = [sh.search_finish(qd, k, sel) for sh in shards]
    ms, mi = ops.topk_merge(torch.stack([o[0] for o in outs]).contiguous(), torch.stack([o[1] for o in outs]).contiguous(), k)
    rs, ri = O.topk(q, db, k)
    np.testing.assert_array_equal(mi.cpu().numpy(), ri)
    np.testing.assert_allclose(ms.cpu().numpy(), rs, rtol=0, atol=1e-12)
    surv_two_phase = sum(sh.stats()["survivors"] for sh in shards)
    alone = 0
    for sh in shards:
        sh.search(qd, k)
        alone += sh.stats()["survivors"]
    assert surv_two_phase < 0.6 * alone, (surv_two_phase, alone)
    # a skewed split (both shards still hold at least k_shard rows) must not break the bound
    tiny = ops.Index(torch.from_numpy(db[:70]).to(DEV), index_offset=0)
    rest = ops.Index(torch.from_numpy(db[70:]).to(DEV), index_offset=70)
    s2 = [tiny.search_begin(qd, k, 30), rest.search_begin(qd, k, 30)]
    sel2 = torch.minimum(s2[0], s2[1]).contiguous()
    o2 = [tiny.search_finish(qd, k, sel2), rest.search_finish(qd, k, sel2)]
    ms2, mi2 = ops.topk_merge(torch.stack([o[0] for o in o2]).contiguous(), torch.stack([o[1] for o in o2]).contiguous(), k)
    np.testing.assert_array_equal(mi2.cpu().numpy(), ri)


def test_two_phase_search_with_a_shard_smaller_than_its_quota():
    """A shard that holds fewer rows than ceil(k / G), all of them close to the query: the naive selection depth
    certifies fewer than k rows and the exchanged threshold would cut true top-k rows (found on the CPU by
    tests/test_properties.py).  dist.shard_quota picks the depth from the shard sizes; merge == oracle."""
    ops = _ops()
    from dirb200.dist import shard_quota
    db, q, _ = synth.make_descriptor_db(5000, 4, dim=256, n_pos=5, db_seed=5, q_seed=6)
    r = np.random.RandomState(9)
    near = q[0][None, :] + 0.02 * r.standard_normal((50, 256)).astype(np.float32)
    near = (near / np.linalg.norm(near, axis=1, keepdims=True)).astype(np.float32)
    full = np.concatenate([near, db]).astype(np.float32)
    k, sizes = 200, [50, 5000]
    c = shard_quota(k, sizes)
    assert c == 150
    qd = torch.from_numpy(q).to(DEV)
    shards = [ops.Index(torch.from_numpy(full[:50].copy()).to(DEV), index_offset=0),
              ops.Index(torch.from_numpy(full[50:].copy()).to(DEV), index_offset=50)]
    sels = [sh.search_begin(qd, k, c) for sh in shards]
    sel = torch.minimum(sels[0], sels[1]).contiguous()
    outs = [sh.search_finish(qd, k, sel) for sh in shards]
    ms, mi = ops.topk_merge(torch.stack([o[0] for o in outs]).contiguous(), torch.stack([o[1] for o in outs]).contiguous(), k)
    rs, ri = O.topk(q, full, k)
    np.testing.assert_array_equal(mi.cpu().numpy(), ri)
    np.testing.assert_allclose(ms.cpu().numpy(), rs, rtol=0, atol=1e-12)


def _targets(pos, n_db, n_junk, seed):
    gnd = synth.oxford_gt(pos, n_junk=n_junk, n_db=n_db, seed=seed)
    offs, rows, flags = [0], [], []
    for g in gnd:
        r = sorted(set(g["ok"]) | set(g["junk"]))
        rows += r
        flags += [1 if x in set(g["ok"]) else 0 for x in r]
        offs.append(len(rows))
    return gnd, np.array(offs, np.int32), np.array(rows, np.int64), np.array(flags, np.uint8)


def test_rank_counts_match_oracle():
    """dirb200_index_rank_count: exact score and number of rows ranking before every positive == the oracle's reading
    of the full ranking, with exact ties (duplicated rows), through the tensor-core pass, through the exact fallback
    (tiny count_cap) and summed over shards; AP from the counts == AP from the score rows."""
    ops = _ops()
    db, q, pos = synth.make_descriptor_db(20000, 9, dim=256, n_pos=8, db_seed=41, q_seed=42)
    db[15000] = db[pos[0, 2]]                      # a duplicate of a positive AFTER it: ranks behind it
    db[3] = db[pos[1, 4]]                          # and one BEFORE a positive (lower index wins the tie)
    gnd, offs, rows, flags = _targets(pos, 20000, 5, 7)
    ref_s, ref_a = O.rank_counts(q, db, offs, rows)
    qd = torch.from_numpy(q).to(DEV)
    index = ops.Index(torch.from_numpy(db).to(DEV))
    sc, above = index.rank_counts(qd, offs, rows, flags)
    m = flags == 1
    np.testing.assert_allclose(sc.cpu().numpy(), ref_s, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(above.cpu().numpy()[m], ref_a[m])
    assert index.stats()["retries"] == 0           # no query needed the exact fallback
    # exact fallback: a candidate capacity far below the depth of the deepest positive (cos 0.15 of 20000 rows)
    index.set_option("count_cap", 1024)
    deep = q.copy()
    deep_db = db.copy()
    deep_db[pos[2, 7]] = db[11]                    # a 'positive' that is just a random row: ~half the database outranks it
    idx2 = ops.Index(torch.from_numpy(deep_db).to(DEV))
    idx2.set_option("count_cap", 1024)
    sc2, above2 = idx2.rank_counts(qd, offs, rows, flags)
    r2s, r2a = O.rank_counts(deep, deep_db, offs, rows)
    np.testing.assert_array_equal(above2.cpu().numpy()[m], r2a[m])
    assert idx2.stats()["retries"] >= 1 and int(r2a[m].max()) > 2000
    # two shards: scores from the owning shard, counts add up
    a_idx = ops.Index(torch.from_numpy(db[:7000]).to(DEV), index_offset=0)
    b_idx = ops.Index(torch.from_numpy(db[7000:]).to(DEV), index_offset=7000)
    t_q = torch.from_numpy(np.repeat(np.arange(9, dtype=np.int32), np.diff(offs))).to(DEV)
    rows_d, flags_d = torch.from_numpy(rows).to(DEV), torch.from_numpy(flags).to(DEV)
    s_sum = a_idx.target_scores(qd, t_q, rows_d) + b_idx.target_scores(qd, t_q, rows_d)
    assert torch.equal(s_sum, sc)
    ab = a_idx.rank_count(qd, offs, rows_d, flags_d, s_sum) + b_idx.rank_count(qd, offs, rows_d, flags_d, s_sum)
    np.testing.assert_array_equal(ab.cpu().numpy()[m], ref_a[m])
    # AP from the counts == AP from the exact score row (generic.py:196-224)
    for i, g in enumerate(gnd):
        sl = slice(offs[i], offs[i + 1])
        at = {int(r): j for j, r in enumerate(rows[sl])}
        js = np.array([ref_s[sl][at[j]] for j in g["junk"]])
        ranks = []
        for p in g["ok"]:
            sp = ref_s[sl][at[p]]
            before = int(((js > sp) | ((js == sp) & (np.array(g["junk"]) < p))).sum())
            ranks.append(int(above.cpu().numpy()[sl][at[p]]) - before)
        ap = O.average_precision(np.sort(np.array(ranks)))
        assert abs(ap - O.eval_query_ap(O.scores_exact(q[i:i + 1], db)[0], g["ok"], g["junk"])) < 1e-12


def test_rank_counts_1m_rows():
    """1M x 2048 synthetic database with planted positives (cosines 0.8 .. 0.15): counts for 16 queries vs the CPU
    oracle (chunked fp64 scores), without any Q x N matrix on either side."""
    ops = _ops()
    N, Q, D, P = 1_000_000, 16, 2048, 10
    g = torch.Generator(device="cuda").manual_seed(21)
    db = ops.l2_normalize(torch.randn((N, D), generator=g, device="cuda"))
    q = ops.l2_normalize(torch.randn((Q, D), generator=g, device="cuda"))
    rows = (torch.arange(Q * P, device="cuda") * 6151 + 17) % N
    cos = torch.linspace(0.8, 0.15, P, device="cuda").repeat(Q)
    noise = torch.randn((Q * P, D), generator=g, device="cuda")
    qq = q.repeat_interleave(P, dim=0)
    noise = noise - (noise * qq).sum(1, keepdim=True) * qq
    noise = noise / noise.norm(dim=1, keepdim=True)
    db[rows] = ops.l2_normalize((cos[:, None] * qq + torch.sqrt(1 - cos * cos)[:, None] * noise).contiguous())
    offs = np.arange(Q + 1, dtype=np.int32) * P
    rows_h = rows.cpu().numpy().astype(np.int64)
    order = np.concatenate([np.argsort(rows_h[i * P:(i + 1) * P]) + i * P for i in range(Q)])
    rows_h = rows_h[order]
    index = ops.Index(db)
    sc, above = index.rank_counts(q, offs, rows_h, np.ones(Q * P, np.uint8))
    assert index.stats()["retries"] == 0
    # oracle: chunked fp64 scores, counts per target
    qs = q.cpu().numpy().astype(np.float64)
    ts = np.array([qs[t // P] @ db[int(rows_h[t])].cpu().numpy().astype(np.float64) for t in range(Q * P)])
    np.testing.assert_allclose(sc.cpu().numpy(), ts, rtol=0, atol=1e-12)
    cnt = np.zeros(Q * P, dtype=np.int64)
    for c0 in range(0, N, 125_000):
        blk = db[c0:c0 + 125_000].cpu().numpy().astype(np.float64)
        s_blk = qs @ blk.T
        idx_blk = np.arange(c0, c0 + blk.shape[0])
        for t in range(Q * P):
            row = s_blk[t // P]
            other = idx_blk != rows_h[t]            # (the block product and the vector product round the target's own score differently)
            cnt[t] += int(((row > ts[t]) & other).sum() + ((row == ts[t]) & (idx_blk < rows_h[t])).sum())
    np.testing.assert_array_equal(above.cpu().numpy(), cnt)
