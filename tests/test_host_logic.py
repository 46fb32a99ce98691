"""CPU tests of the host-side logic: dataset / AP bookkeeping against the reference's golden values, the
transform chain, CLI surface, and the multi-rank top-k exchange under gloo (world_size 2)."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

import synthdata as synth
from conftest import REPO
from oracle import dir_oracle as O


def _gt_dataset(tmp_path, g, revisited=False):
    from dirtorch.datasets import ImageListRelevants
    db, q, pos = synth.make_descriptor_db(int(g["n_db"]), int(g["n_q"]), dim=int(g["dim"]), n_pos=int(g["n_pos"]),
                                          db_seed=int(g["db_seed"]), q_seed=int(g["q_seed"]))
    gnd = synth.oxford_gt(pos, n_junk=int(g["n_junk"]), n_db=int(g["n_db"]), seed=int(g["gt_seed"]))
    if revisited:
        gnd = [{"bbx": e["bbx"], "easy": e["ok"][:2], "hard": e["ok"][2:], "junk": e["junk"]} for e in gnd]
    imlist = ["im%04d" % i for i in range(int(g["n_db"]))]
    f = tmp_path / ("gnd%d.pkl" % revisited)
    with open(f, "wb") as fh:
        pickle.dump({"imlist": imlist, "qimlist": imlist[:int(g["n_q"])], "gnd": gnd}, fh)
    return ImageListRelevants(str(f), root=str(tmp_path)), db, q


def test_dataset_ap_matches_reference(golden, tmp_path):
    g = golden("rank_ap.npz")
    ds, db, q = _gt_dataset(tmp_path, g)
    assert ds.nimg == 400 and ds.nquery == 6 and ds.get_key(3) == "im0003.jpg"
    assert ds.get_query_db().nimg == 6
    for i in range(6):
        assert abs(ds.eval_query_AP(i, g["scores"][i]) - g["aps"][i]) < 1e-12
    ds2, _, _ = _gt_dataset(tmp_path, g, revisited=True)
    for i in range(6):
        ap = ds2.eval_query_AP(i, g["scores"][i])
        assert abs(ap["easy"] - g["aps_easy"][i]) < 1e-12
        assert abs(ap["medium"] - g["aps_medium"][i]) < 1e-12
        assert abs(ap["hard"] - g["aps_hard"][i]) < 1e-12
    with pytest.raises(AssertionError):
        ds.eval_query_AP(0, g["scores"][0][:10])           # wrong shape, generic.py:201


def test_ap_from_ranked_prefix_equals_ap_from_scores(golden, tmp_path):
    """The top-k evaluation path: AP from the first entries of the ranking == AP from the full score row."""
    g = golden("rank_ap.npz")
    for revisited in (False, True):
        ds, db, q = _gt_dataset(tmp_path, g, revisited=revisited)
        for i in range(6):
            full = ds.eval_query_AP(i, g["scores"][i])
            order = O.rank_desc(g["scores"][i])
            worst = max(int(np.where(order == p)[0][0]) for p in synth.oxford_gt(
                synth.make_descriptor_db(400, 6, dim=32, n_pos=5, db_seed=21, q_seed=22)[2], n_junk=3, n_db=400, seed=5)[i]["ok"])
            got = ds.eval_query_AP_from_ranking(i, order[:worst + 1])
            if isinstance(full, dict):
                assert all(abs(got[m] - full[m]) < 1e-12 for m in full)
            else:
                assert abs(got - full) < 1e-12
            short = ds.eval_query_AP_from_ranking(i, order[:1])          # prefix too short -> unknown
            assert short is None or (isinstance(short, dict) and any(v is None for v in short.values())) or worst == 0


def test_ap_from_rank_counts_equals_ap_from_scores(golden, tmp_path):
    """eval_query_AP_from_counts (what the GPU counting kernel feeds) == eval_query_AP on the golden score rows, classic
    and revisited protocols; the counts come from the oracle's restatement of the ranking (O.rank_counts)."""
    g = golden("rank_ap.npz")
    for revisited in (False, True):
        ds, db, q = _gt_dataset(tmp_path, g, revisited=revisited)
        offs, rows, flags = [0], [], []
        for i in range(6):
            r, f = ds.rank_targets(i)
            rows.append(r)
            flags.append(f)
            offs.append(offs[-1] + len(r))
        rows = np.concatenate(rows)
        sc, above = O.rank_counts(q, db, offs, rows)
        for i in range(6):
            sl = slice(offs[i], offs[i + 1])
            ap = ds.eval_query_AP_from_counts(i, rows[sl], sc[sl], above[sl])
            ref = ds.eval_query_AP(i, O.scores_exact(q[i:i + 1], db)[0])
            if revisited:
                assert ap.keys() == ref.keys() and all(abs(ap[m] - ref[m]) < 1e-12 for m in ap)
                assert abs(ap["easy"] - g["aps_easy"][i]) < 1e-12
            else:
                assert abs(ap - ref) < 1e-12 and abs(ap - g["aps"][i]) < 1e-12


def test_label_datasets_ap_matches_reference(golden, tmp_path):
    """Label-based AP / top-k (dataset.py:69-101) on labelled image lists (generic.py:44-105) vs the reference."""
    from dirtorch import datasets as D
    g = golden("label_ap.npz")
    labels, qlabels = [str(v) for v in g["labels"]], [str(v) for v in g["qlabels"]]
    with open(tmp_path / "db.txt", "w") as f:
        f.write("\n".join("im%03d.jpg %s" % (i, l) for i, l in enumerate(labels)) + "\n")
    with open(tmp_path / "q.txt", "w") as f:
        f.write("\n".join("q%03d.jpg %s" % (i, l) for i, l in enumerate(qlabels)) + "\n")
    ds = D.create('ImageListLabels("%s", root="%s")' % (tmp_path / "db.txt", tmp_path))
    assert len(ds) == 60 and ds.nclass == int(g["nclass"]) and ds.get_query_db() is ds
    assert ds.get_key(3) == "im003.jpg" and ds.get_label(3) == labels[3] and ds.classes[ds.get_label(3, toint=True)] == labels[3]
    np.testing.assert_array_equal(ds.get_query_groundtruth(0), g["gt0"])
    aps = np.array([ds.eval_query_AP(q, g["scores"][q]) for q in range(60)], dtype=np.float64)
    np.testing.assert_allclose(aps, g["aps"], rtol=0, atol=1e-12)
    assert aps[5] == -1                                                           # a class of one image
    dsq = D.ImageListLabelsQ(str(tmp_path / "db.txt"), str(tmp_path / "q.txt"), root=str(tmp_path))
    qdb = dsq.get_query_db()
    assert dsq.nquery == 7 and len(qdb) == 7 and qdb.get_key(2) == "q002.jpg" and dsq.nclass == int(g["nclass_q"]) == qdb.nclass
    qaps = np.array([dsq.eval_query_AP(q, g["qscores"][q]) for q in range(7)], dtype=np.float64)
    np.testing.assert_allclose(qaps, g["qaps"], rtol=0, atol=1e-12)
    # top-k: the reference's own eval_query_top needs np.bool8 (gone in NumPy 2); checked against its definition
    s = g["scores"][1]
    order = np.argsort(-s)
    top = ds.eval_query_top(1, s, k=(1, 5, 100))
    assert set(top) == {1, 5} and top[5] == float(any(labels[j] == labels[1] for j in order[:5]))
    with pytest.raises(ValueError):
        ds.get_query_groundtruth(0, "bogus")
    # a json list gives the same dataset
    import json
    json.dump({"im%03d.jpg" % i: l for i, l in enumerate(labels)}, open(tmp_path / "db.json", "w"))
    dj = D.ImageListLabels(str(tmp_path / "db.json"), root=str(tmp_path))
    assert dj.labels == labels and dj.cls_idx == ds.cls_idx


def test_transforms_and_loader(tmp_path):
    from PIL import Image
    from dirb200.loader import Scale, create_transforms, get_loader
    from dirtorch.datasets import ImageList
    assert Scale(0.7).get_params((1024, 1024)) == (717, 717)          # int(0.5 + s*w), transforms.py:168
    assert Scale(1.4).get_params((1024, 768)) == (1434, 1075)
    assert Scale(256).get_params((512, 1024)) == (256, 512)
    u8 = synth.make_images_u8(3, 40, 56, seed=2)
    names = []
    for i in range(3):
        Image.fromarray(u8[i]).save(tmp_path / ("a%d.png" % i))
        names.append("a%d.png" % i)
    ds = ImageList(imgs=names, root=str(tmp_path))
    pre = dict(mean=synth.RGB_MEANS, std=synth.RGB_STDS, input_size=224)
    batches = [b[0] for b in get_loader(ds, "", False, preprocess=pre, output=["img"], batch_size=3, threads=1)]
    assert torch.allclose(batches[0], synth.normalise_images(u8), atol=1e-6)     # ToTensor + Normalize, transforms.py:27
    scaled = [b[0] for b in get_loader(ds, "Scale(0.5)", False, preprocess=pre, output=["img"], batch_size=1, threads=1)]
    assert tuple(scaled[0].shape) == (1, 3, 20, 28)
    with pytest.raises(SyntaxError):
        create_transforms("Bogus(3)", to_tensor=True, **pre)


def test_test_time_transforms_match_reference(golden):
    """Scale / Pad / PadSquare / CenterCrop / Identity (+ ToTensor, Normalize) against CRCs of the reference's
    output tensors (tests/golden/make_golden.py: gold_transforms); unknown transforms fail with a clear message."""
    import zlib
    from PIL import Image
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    from dirb200 import loader
    g = golden("transforms.npz")
    chains = ["Pad(70)", "PadSquare()", "PadSquare(60)", "PadSquare(100)", "CenterCrop(32)", "CenterCrop((20,40))",
              "Scale(48), CenterCrop(40)", "Scale(0.7), Pad(64, color=(0.5,0.5,0.5))", "Identity()",
              "CenterCrop(30, padding=4)", "Scale(0.5)", "Scale(1.4)", "Scale(40, largest=True)", "Scale((50, 30))"]
    r = np.random.RandomState(0)
    i = 0
    for (h, w) in [(50, 80), (80, 50), (64, 64), (33, 97)]:
        img = Image.fromarray(r.randint(0, 256, (h, w, 3), dtype=np.uint8))
        for chain in chains:
            t = loader.create_transforms(chain, to_tensor=True, mean=synth.RGB_MEANS, std=synth.RGB_STDS)(img).numpy()
            assert tuple(t.shape) == tuple(g["shapes"][i]), (chain, h, w)
            assert zlib.crc32(np.ascontiguousarray(t).tobytes()) == int(g["crcs"][i]), (chain, h, w)
            i += 1
    assert i == len(g["crcs"])
    with pytest.raises(SyntaxError, match="unsupported transform"):
        loader.create_transforms("RandomCrop(32)", to_tensor=True, mean=0, std=1)


def test_cli_surface_and_no_cpu_path():
    from dirtorch import extract_features, test_dir
    from dirtorch.utils import common
    for fn in ("expand_descriptors", "extract_image_features", "eval_model", "load_model"):
        assert callable(getattr(test_dir, fn))
    assert callable(extract_features.extract_features) and callable(extract_features.load_model)
    for fn in ("tonumpy", "matmul", "pool", "transform", "whiten_features", "torch_set_gpu", "load_checkpoint",
               "switch_model_to_cuda", "variables"):
        assert callable(getattr(common, fn))
    with pytest.raises(RuntimeError):
        common.torch_set_gpu([-1])                                   # the reference would run on CPU here
    with pytest.raises(TypeError):
        common.matmul([1, 2], [3, 4])                                # common.py:37
    x = [torch.zeros(2, 4), torch.zeros(2, 4)]
    with pytest.raises(ValueError):
        common.pool(x, "bogus")                                       # common.py:55
    assert common.pool(x[:1], "gem") is x[0]                          # single chain: identity, common.py:42-43
    import inspect
    sig = inspect.signature(test_dir.eval_model)
    assert list(sig.parameters)[:3] == ["db", "net", "trfs"] and sig.parameters["pooling"].default == "mean"
    assert inspect.signature(test_dir.extract_image_features).parameters["batch_size"].default == 8
    assert inspect.signature(test_dir.expand_descriptors).parameters["alpha"].default == 0
    # CLI flags of the reference (test_dir.py:198-221 / extract_features.py:86-105)
    with pytest.raises(SystemExit):
        test_dir.test_dir_main(["--help"])
    with pytest.raises(SystemExit):
        extract_features.extract_features_main(["--dataset", "x"])   # --checkpoint is required


def test_checkpoint_round_trip_and_model_size(tmp_path, capsys):
    """save_checkpoint / load_checkpoint / model_size (common.py:100-147,178-184) on the duck-typed model."""
    from dirtorch import nets
    from dirtorch.utils import common, evaluation
    net = nets.create_model("resnet50_rmac")
    assert common.model_size(net) == 27757558 and len(net.state_dict()) == 321      # = the reference's own count
    f = str(tmp_path / "sub" / "ck.pt")
    state = {"state_dict": {"module." + k: v for k, v in net.state_dict().items()},
             "model_options": dict(arch="resnet50_rmac", out_dim=2048, pooling="gem", gemp=3), "epoch": 3}
    common.save_checkpoint(state, True, f)
    assert os.path.isfile(f) and os.path.isfile(f + ".best") and "saving to" in capsys.readouterr().out
    ck = common.load_checkpoint(f + ".best")
    assert "(epoch 3)" in capsys.readouterr().out
    assert list(ck["state_dict"]) == list(net.state_dict())                          # 'module.' stripped
    assert all(torch.equal(ck["state_dict"][k], v) for k, v in net.state_dict().items())
    assert abs(evaluation.compute_AP([1, 0, 1, 0], [0.9, 0.8, 0.7, 0.1]) - 5.0 / 6.0) < 1e-12


def test_shard_rows_cover():
    from dirb200.dist import shard_rows
    for n, w in ((1_000_000, 8), (100, 8), (7, 3), (5, 8)):
        rs = [shard_rows(n, w, r) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    assert shard_rows(1_000_000, 8, 3) == (375000, 500000)


def _gloo_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    import synthdata as synth
    from dirb200.dist import all_gather_packed, shard_rows
    from oracle import dir_oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db, q, _ = synth.make_descriptor_db(3000, 11, dim=64, n_pos=4, db_seed=1, q_seed=2)
    k = 16
    s0, s1 = shard_rows(db.shape[0], world, rank)
    ls, li = O.topk(q, db[s0:s1], k)                                   # stands in for dirb200_index_search
    packed = torch.empty((2, q.shape[0], k), dtype=torch.int64)
    packed[0] = torch.from_numpy(ls).view(torch.int64)
    packed[1] = torch.from_numpy(li + s0)                              # global indices = local row + offset
    gathered = all_gather_packed(packed)
    assert tuple(gathered.shape) == (world, 2, q.shape[0], k)
    sc = [gathered[g, 0].view(torch.float64).numpy() for g in range(world)]
    ix = [gathered[g, 1].numpy() for g in range(world)]
    ms, mi = O.merge_topk(sc, ix, k)
    rs, ri = O.topk(q, db, k)
    ok = bool(np.array_equal(mi, ri) and np.allclose(ms, rs, atol=1e-15))
    # alpha-QE partial sums: every rank sums the neighbours it owns, all-reduce, add the query, normalise
    part = np.zeros_like(q, dtype=np.float64)
    for i in range(q.shape[0]):
        for j in range(2):
            r = ri[i, j]
            if s0 <= r < s1:
                part[i] += db[r].astype(np.float64) * rs[i, j] ** 0.5
    t = torch.from_numpy(part)
    dist.all_reduce(t)
    out = t.numpy() + q
    out /= np.linalg.norm(out, axis=1, keepdims=True)
    ref = O.expand_descriptors(q, db=db, k=2, alpha=0.5)
    ok = ok and float(np.abs(out - ref).max()) < 1e-6
    with open(os.path.join(tmp, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_topk_exchange_gloo_world2(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


def _shard_sizes_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    import dirb200  # noqa: F401
    from dirb200 import dist as ddist
    from dirb200 import ops

    class _NoGpuIndex:                                     # the real one uploads to the GPU
        def __init__(self, db32, index_offset=0, db16=None):
            self.db32, self.n, self.dim = db32, db32.shape[0], db32.shape[1]

        def set_option(self, key, value):
            pass
    ops.Index = _NoGpuIndex
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = ddist.ShardedIndex(torch.zeros(([3, 1000, 0][rank], 64)), row_offset=0)
    ok = sh.shard_sizes == [3, 1000, 0] and ddist.shard_quota(100, sh.shard_sizes) == 97
    with open(os.path.join(tmp, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_sharded_index_learns_all_shard_sizes_gloo_world3(tmp_path):
    """ShardedIndex gathers the row counts of all shards at construction (they set the selection depth of the two-phase
    search, dist.shard_quota); exchange plumbing under gloo with a stand-in for the GPU index."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_shard_sizes_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert all(open(tmp_path / ("ok%d" % r)).read() == "1" for r in range(3))


def _peer_handles_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, REPO)
    import torch.distributed as dist
    import dirb200  # noqa: F401
    from dirb200 import dist as ddist
    from dirb200 import ops

    class _NoGpuIndex:
        def __init__(self, db32, index_offset=0, db16=None):
            self.db32, self.n, self.dim = db32, db32.shape[0], db32.shape[1]
            self.calls = []

        def set_option(self, key, value):
            pass

        def search_sharded(self, x, q32, k, k_shard, phase=0):
            self.calls.append((x.rank, q32.shape[0], k, k_shard))
            return "peer"

        def check(self):
            self.calls.append("check")

    class _NoGpuExchange:                                  # the real one allocates a device window and opens CUDA IPC handles
        def __init__(self, device_index, world, rank, max_q, max_k):
            self.world, self.rank, self.max_q, self.max_k, self.opened = world, rank, max_q, max_k, None

        def ipc_handle(self):
            return bytes([self.rank + 1]) * 64

        def open(self, handles):
            self.opened = [bytes(h) for h in handles]
    ops.Index, ops.Exchange = _NoGpuIndex, _NoGpuExchange
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = ddist.ShardedIndex(torch.zeros((10 + rank, 64)), row_offset=0).enable_peer_exchange(max_q=16, max_k=8)
    x = sh.xchg
    ok = x.rank == rank and x.world == world and x.opened == [bytes([g + 1]) * 64 for g in range(world)]
    # searches that fit the window go through the exchange (with the shard quota), larger ones through the collectives
    out = sh.search(torch.zeros((5, 64)), 8)
    ok = ok and out == "peer" and sh.local.calls == [(rank, 5, 8, ddist.shard_quota(8, sh.shard_sizes)), "check"]
    with open(os.path.join(tmp, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_peer_exchange_handles_travel_in_rank_order_gloo_world3(tmp_path):
    """ShardedIndex.enable_peer_exchange: every rank creates its window, the 64-byte IPC handles are all-gathered and each
    rank opens them in rank order; search() then routes through dirb200_index_search_sharded with k_shard = shard_quota.
    Plumbing under gloo with stand-ins for the GPU objects (the device side is covered by the -m gpu tests and
    tools/dist_check.py under NCCL)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_peer_handles_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    assert all(open(tmp_path / ("ok%d" % r)).read() == "1" for r in range(3))


class _FakeNet:
    """Stands in for the GPU network in host-plumbing tests: descriptor = per-channel mean / std of the image."""
    iscuda = False
    arch = "fake"
    descriptor_dim = 6
    preprocess = dict(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225], input_size=224)
    pca = None

    def eval(self):
        return self

    def __call__(self, imgs):
        return torch.cat([imgs.mean(dim=(2, 3)), imgs.std(dim=(2, 3))], dim=1)


def _fake_pool_and_normalize(descs, pooling, gemp):
    x = torch.stack([d.float() for d in descs]).mean(0)
    return x / x.norm(dim=1, keepdim=True)


def _make_image_list(tmp, n):
    from PIL import Image
    from dirtorch.datasets import ImageList
    r = np.random.RandomState(8)
    names = []
    for i in range(n):
        names.append("im%02d.png" % i)
        Image.fromarray(r.randint(0, 256, (24 + i, 30, 3), dtype=np.uint8)).save(os.path.join(tmp, names[-1]))
    return ImageList(imgs=names, root=tmp)


def _extract_store_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch.distributed as dist
    import dirb200  # noqa: F401
    from dirb200 import pipeline
    import test_host_logic as T
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pipeline._pool_and_normalize = T._fake_pool_and_normalize      # the real one runs CUDA kernels
    ds = T._make_image_list(tmp, 7) if rank == 0 else None
    dist.barrier()
    if ds is None:
        from dirtorch.datasets import ImageList
        ds = ImageList(imgs=["im%02d.png" % i for i in range(7)], root=tmp)
    st = pipeline.extract_to_store(ds, T._FakeNet(), ["", "Scale(0.5)"], os.path.join(tmp, "store"), threads=1)
    ok = len(st) == 7 and st.dim == 6 and [s["n_rows"] for s in st.shards] == [3, 4][:world] and st.meta["n_images"] == 7
    with open(os.path.join(tmp, "ok%d" % rank), "w") as f:
        f.write("1" if ok else "0")
    dist.destroy_process_group()


def test_extract_to_store_gloo_world2_matches_single_process(tmp_path, monkeypatch):
    """Sharded extraction plumbing (row ranges, per-rank shard files, manifest) with a stand-in network: two gloo
    ranks write the same store as one process."""
    import socket
    import torch.multiprocessing as mp
    from dirb200 import pipeline, store
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_extract_store_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"
    two = store.DescriptorStore(str(tmp_path / "store")).read_all()
    monkeypatch.setattr(pipeline, "_pool_and_normalize", _fake_pool_and_normalize)
    from dirtorch.datasets import ImageList
    ds = ImageList(imgs=["im%02d.png" % i for i in range(7)], root=str(tmp_path))
    one = pipeline.extract_to_store(ds, _FakeNet(), ["", "Scale(0.5)"], str(tmp_path / "single"), threads=1)
    assert [s["n_rows"] for s in one.shards] == [7] and one.meta["trfs"] == ["", "Scale(0.5)"]
    np.testing.assert_array_equal(one.read_all(), two)
    np.testing.assert_allclose(np.linalg.norm(two, axis=1), 1.0, atol=1e-6)
    # more ranks than images: empty shards are legal
    tiny = pipeline.extract_to_store(ImageList(imgs=[], root=str(tmp_path)), _FakeNet(), "", str(tmp_path / "empty"))
    assert len(tiny) == 0 and tiny.dim == 6


def test_resize_coefficients_reproduce_pil():
    """The host-side coefficient tables of the GPU resize (dirb200_resize_coeffs), applied in numpy integer
    arithmetic, reproduce PIL's Image.resize(BILINEAR) byte for byte (the kernels do exactly this arithmetic)."""
    from PIL import Image
    from dirb200 import ops
    r = np.random.RandomState(0)

    def apply(a, ho, wo):
        h, w, c = a.shape
        bx, kx = ops.resize_coeffs(w, wo)
        tmp = np.zeros((h, wo, c), dtype=np.uint8)
        for xx in range(wo):
            x0, n = bx[xx]
            acc = (1 << 21) + (a[:, x0:x0 + n, :].astype(np.int64) * kx[xx, :n][None, :, None]).sum(axis=1)
            tmp[:, xx, :] = np.clip(acc >> 22, 0, 255)
        by, ky = ops.resize_coeffs(h, ho)
        out = np.zeros((ho, wo, c), dtype=np.uint8)
        for yy in range(ho):
            y0, n = by[yy]
            acc = (1 << 21) + (tmp[y0:y0 + n].astype(np.int64) * ky[yy, :n][:, None, None]).sum(axis=0)
            out[yy] = np.clip(acc >> 22, 0, 255)
        return out

    for (h, w, ho, wo) in ((37, 53, 26, 37), (37, 53, 52, 74), (64, 64, 45, 45), (100, 80, 33, 200), (50, 50, 50, 70)):
        a = r.randint(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(a).resize((wo, ho), Image.BILINEAR))
        assert np.array_equal(apply(a, ho, wo), ref), (h, w, ho, wo)
