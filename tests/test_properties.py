"""Property-based tests (hypothesis) of the host-side logic: row sharding, the descriptor store, AP from a ranked
prefix vs AP from the full score row, PIL-exact resize coefficient tables.  CPU only."""
import os
import pickle

import numpy as np
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from dirb200 import dist as ddist
from dirb200 import store as S

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(max_examples=200, **COMMON)
@given(n=st.integers(0, 10**7), world=st.integers(1, 64))
def test_shard_rows_partitions_the_rows(n, world):
    rs = [ddist.shard_rows(n, world, r) for r in range(world)]
    assert rs[0][0] == 0 and rs[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    sizes = [b - a for a, b in rs]
    assert min(sizes) >= 0 and max(sizes) - min(sizes) <= 1                      # balanced to within one row


@settings(max_examples=40, **COMMON)
@given(n=st.integers(0, 300), dim=st.sampled_from([1, 8, 64]), per=st.integers(1, 128), world=st.integers(1, 9),
       f16=st.booleans(), data=st.data())
def test_store_reads_equal_slices(tmp_path_factory, n, dim, per, world, f16, data):
    path = str(tmp_path_factory.mktemp("store"))
    db = np.random.RandomState(n * 31 + dim).standard_normal((n, dim)).astype(np.float32)
    stored = db.astype(np.float16) if f16 else db
    s = S.write_store(path, db, rows_per_shard=per, dtype=np.float16 if f16 else np.float32)
    assert len(s) == n and s.dim == dim and sum(x["n_rows"] for x in s.shards) == n
    a = data.draw(st.integers(0, n))
    b = data.draw(st.integers(a, n))
    assert np.array_equal(s.read_rows(a, b), stored[a:b])
    parts = [s.read_rows(*s.rank_range(r, world)) for r in range(world)]
    assert np.array_equal(np.concatenate(parts) if parts else stored[:0], stored)
    assert all(s.rank_range(r, world) == ddist.shard_rows(n, world, r) for r in range(world))


def _relevants_dataset(tmp, n, gnd):
    from dirtorch.datasets import ImageListRelevants
    names = ["im%04d" % i for i in range(n)]
    with open(os.path.join(tmp, "gnd.pkl"), "wb") as f:
        pickle.dump({"imlist": names, "qimlist": names[:len(gnd)], "gnd": gnd}, f)
    return ImageListRelevants(os.path.join(tmp, "gnd.pkl"), root=tmp)


@settings(max_examples=60, **COMMON)
@given(n=st.integers(8, 120), seed=st.integers(0, 10**6), revisited=st.booleans(), data=st.data())
def test_ap_from_full_ranking_equals_ap_from_scores(tmp_path_factory, n, seed, revisited, data):
    """eval_query_AP_from_ranking over the COMPLETE ranking == eval_query_AP on the score row (distinct scores), and a
    prefix either reproduces it or says 'unknown' - never a different number."""
    r = np.random.RandomState(seed)
    idx = r.permutation(n)
    n_pos, n_junk = int(r.randint(1, 6)), int(r.randint(0, 4))
    pos, junk = [int(v) for v in idx[:n_pos]], [int(v) for v in idx[n_pos:n_pos + n_junk]]
    if revisited:
        cut = int(r.randint(0, n_pos + 1))
        g = {"bbx": (0, 0, 1, 1), "easy": pos[:cut], "hard": pos[cut:], "junk": junk}
    else:
        g = {"bbx": (0, 0, 1, 1), "ok": pos, "junk": junk}
    ds = _relevants_dataset(str(tmp_path_factory.mktemp("gt")), n, [g])
    scores = r.permutation(n).astype(np.float64) / n                             # distinct: no tie ambiguity
    order = np.argsort(-scores)
    full = ds.eval_query_AP(0, scores)
    got = ds.eval_query_AP_from_ranking(0, order)
    k = data.draw(st.integers(1, n))
    part = ds.eval_query_AP_from_ranking(0, order[:k])
    if isinstance(full, dict):
        for m in full:
            assert abs(got[m] - full[m]) < 1e-12
            assert part[m] is None or abs(part[m] - full[m]) < 1e-12
    else:
        assert abs(got - full) < 1e-12
        assert part is None or abs(part - full) < 1e-12


@settings(max_examples=25, **COMMON)
@given(h=st.integers(2, 40), w=st.integers(2, 40), oh=st.integers(1, 60), ow=st.integers(1, 60), seed=st.integers(0, 999))
def test_resize_tables_reproduce_pil_for_any_size(h, w, oh, ow, seed):
    """The coefficient tables the GPU resize uses (dirb200_resize_coeffs, a host function of the library), applied
    in numpy integer arithmetic exactly as the kernels do, reproduce PIL's BILINEAR resize byte for byte."""
    from PIL import Image
    from dirb200 import ops
    img = np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.array(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))

    def apply(a, in_size, out_size):                      # a: (in_size, ...) -> (out_size, ...), along axis 0
        bounds, kk = ops.resize_coeffs(in_size, out_size)
        out = np.empty((out_size,) + a.shape[1:], np.uint8)
        for o in range(out_size):
            x0, n = int(bounds[o, 0]), int(bounds[o, 1])
            acc = (1 << 21) + np.tensordot(kk[o, :n].astype(np.int64), a[x0:x0 + n].astype(np.int64), axes=(0, 0))
            out[o] = np.clip(acc >> 22, 0, 255)
        return out

    horiz = apply(img.transpose(1, 0, 2), w, ow).transpose(1, 0, 2)              # PIL: horizontal pass first
    got = apply(horiz, h, oh)
    assert np.array_equal(got, ref)
