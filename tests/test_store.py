"""Row-sharded on-disk descriptor store (deep-image-retrieval_b200/store.py): format, round trips, reference .npy
compatibility and rank ranges.  CPU only; the GPU leg (ShardedIndex.from_store) is in test_gpu_pipeline.py."""
import json
import os

import numpy as np
import pytest
import torch

from dirb200 import dist as ddist
from dirb200 import store as S


def _db(n=1000, d=64, seed=3):
    x = np.random.RandomState(seed).standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("rows_per_shard", [1, 7, 128, 1000, 4096])
def test_round_trip_and_npy_compat(tmp_path, rows_per_shard):
    db = _db(257 if rows_per_shard == 1 else 1000)
    st = S.write_store(str(tmp_path / "s"), db, rows_per_shard=rows_per_shard, meta={"arch": "resnet101_rmac"})
    assert len(st) == db.shape[0] and st.dim == 64 and st.dtype == np.float32 and st.meta["arch"] == "resnet101_rmac"
    assert len(st.shards) == -(-db.shape[0] // rows_per_shard)
    assert np.array_equal(st.read_all(), db)
    # every shard is a plain .npy; concatenated they are the array the reference writes (test_dir.py:130-134)
    cat = np.concatenate([np.load(os.path.join(st.path, s["file"])) for s in st.shards])
    assert np.array_equal(cat, db)
    # ragged ranges across shard boundaries
    r = np.random.RandomState(0)
    for _ in range(20):
        a, b = sorted(r.randint(0, db.shape[0] + 1, 2))
        assert np.array_equal(st.read_rows(a, b), db[a:b])
    assert st.read_rows(5, 5).shape == (0, 64)
    with pytest.raises(IndexError):
        st.read_rows(0, db.shape[0] + 1)
    # export / import of the reference's single-file format
    st.to_npy(str(tmp_path / "feats.bdescs.npy"))
    assert np.array_equal(np.load(tmp_path / "feats.bdescs.npy"), db)
    st2 = S.DescriptorStore.from_npy(str(tmp_path / "feats.bdescs.npy"), str(tmp_path / "s2"), rows_per_shard=300)
    assert np.array_equal(st2.read_all(), db)


def test_streaming_writer_uneven_appends(tmp_path):
    db = _db(999)
    with S.DescriptorStoreWriter(str(tmp_path / "s"), 64, rows_per_shard=100) as w:
        pos = 0
        for n in [1, 250, 3, 99, 100, 546]:
            w.append(torch.from_numpy(db[pos:pos + n]) if n % 2 else db[pos:pos + n])
            pos += n
        with pytest.raises(ValueError):
            w.append(np.zeros((2, 63), np.float32))
    st = S.DescriptorStore(str(tmp_path / "s"))
    assert [s["n_rows"] for s in st.shards] == [100] * 9 + [99]
    assert np.array_equal(st.read_all(), db)


def test_fp16_store_and_empty(tmp_path):
    db = _db(300)
    st = S.write_store(str(tmp_path / "h"), db, rows_per_shard=128, dtype=np.float16)
    assert st.dtype == np.float16 and np.array_equal(st.read_all(), db.astype(np.float16))
    d32, d16 = st.load_rows_to_device(10, 290, "cpu", chunk_rows=64)
    assert np.array_equal(d16.numpy(), db[10:290].astype(np.float16))
    assert np.array_equal(d32.numpy(), db[10:290].astype(np.float16).astype(np.float32))   # exact widening
    e = S.write_store(str(tmp_path / "e"), np.zeros((0, 64), np.float32))
    assert len(e) == 0 and e.read_all().shape == (0, 64)
    with pytest.raises(TypeError):
        S.write_store(str(tmp_path / "bad"), db.astype(np.float64))


def test_rank_shards_and_reading_world(tmp_path):
    # written by 3 "ranks" with uneven counts, read back under world sizes 1, 2, 8
    db = _db(1003)
    cuts = [0, 400, 401, 1003]
    for r in range(3):
        S.write_rank_shard(str(tmp_path / "s"), db[cuts[r]:cuts[r + 1]], r)
    st = S.finalize_rank_shards(str(tmp_path / "s"), 3, meta={"note": "x"})
    assert [s["row_start"] for s in st.shards] == cuts[:3]
    for world in (1, 2, 8):
        got = []
        for rank in range(world):
            a, b = st.rank_range(rank, world)
            assert (a, b) == ddist.shard_rows(len(st), world, rank)
            d32, d16 = st.load_rows_to_device(a, b, "cpu", chunk_rows=97)
            assert np.array_equal(d16.numpy(), db[a:b].astype(np.float16))     # round-to-nearest-even, as f32_to_f16
            got.append(d32.numpy())
        assert np.array_equal(np.concatenate(got), db)


def test_manifest_validation(tmp_path):
    db = _db(64)
    st = S.write_store(str(tmp_path / "s"), db, rows_per_shard=32)
    assert S.DescriptorStore.is_store(st.path) and not S.DescriptorStore.is_store(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        S.DescriptorStore(str(tmp_path / "nope"))
    m = os.path.join(st.path, S.MANIFEST)
    doc = json.load(open(m))
    bad = dict(doc, n_rows=65)
    json.dump(bad, open(m, "w"))
    with pytest.raises(ValueError):
        S.DescriptorStore(st.path)
    bad = dict(doc, version=99)
    json.dump(bad, open(m, "w"))
    with pytest.raises(ValueError):
        S.DescriptorStore(st.path)
    json.dump(doc, open(m, "w"))
    np.save(os.path.join(st.path, doc["shards"][1]["file"]), db[:5])            # shard file disagrees with manifest
    with pytest.raises(ValueError):
        S.DescriptorStore(st.path).read_all()
