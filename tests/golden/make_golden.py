#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports naver/deep-image-retrieval from /root/reference (read-only, nothing is copied),
feeds it the seeded synthetic inputs of ``synthdata.py`` and stores the
reference's outputs as small ``.npz`` files.  Inputs are NOT stored when they can be
regenerated from a seed; the seeds/shapes are recorded in each file.
"""
import importlib.util
import os
import pickle
import sys
import tempfile
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

os.environ.setdefault("DB_ROOT", tempfile.mkdtemp(prefix="dbroot_"))
os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")
sys.path.insert(0, REF)

import torch  # noqa: E402

torch.set_num_threads(os.cpu_count())

spec = importlib.util.spec_from_file_location("synth", os.path.join(REPO, "synthdata.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

import dirtorch.nets as nets  # noqa: E402  (the reference)
from dirtorch.utils import common  # noqa: E402
from dirtorch.nets.layers.pooling import GeneralizedMeanPooling  # noqa: E402
import dirtorch.test_dir as ref_test_dir  # noqa: E402
from dirtorch.datasets.generic import ImageListRelevants  # noqa: E402

assert nets.__file__.startswith(REF), nets.__file__


def ref_model(arch, seed, **kw):
    net = nets.create_model(arch, pretrained="", **kw)
    sd = synth.make_state_dict(arch, seed=seed, out_dim=kw.get("out_dim", 2048), gemp=kw.get("gemp", 3))
    if not kw.get("pooling", "gem").startswith("gem"):
        sd.pop("adpool.p", None)                # max/avg pooling has no learnable p (rmac_resnet.py:24-31)
    if kw.get("mode", 1) == 0:                  # FPN mode 0 has no lateral / smoothing convs (rmac_resnet_fpn.py:27-30)
        sd.pop("conv1x5.weight", None)
        sd.pop("conv3c4.weight", None)
    net.load_state_dict(sd, strict=True)
    net.eval()
    return net


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: getattr(v, "shape", v) for k, v in arrays.items()})


@torch.no_grad()
def gold_extract():
    # config 1: Resnet50-GeM, 4 x 224 x 224
    net = ref_model("resnet50_rmac", seed=0)
    x = synth.make_images(4, 224, 224, seed=1234)
    d4 = net(x).numpy()
    d1 = net(x[:1]).numpy()                       # B=1 -> squeezed (2048,)
    xr = synth.make_images(2, 160, 224, seed=99)  # non-square
    dr = net(xr).numpy()
    # intermediate statistics for debugging a mismatching kernel
    feats = {}
    hooks = []
    for lname in ["maxpool", "layer1", "layer2", "layer3", "layer4"]:
        hooks.append(getattr(net, lname).register_forward_hook(
            lambda m, i, o, lname=lname: feats.__setitem__(lname, o.detach())))
    net(x)
    for h in hooks:
        h.remove()
    stats = {("stat_" + k): np.array([v.mean().item(), v.abs().mean().item(), v.abs().max().item()]) for k, v in feats.items()}
    # one exact activation slice per stage: image 0, all channels, one pixel
    slices = {("slice_" + k): v[0, :, v.shape[2] // 2, v.shape[3] // 3].numpy() for k, v in feats.items()}
    save("extract_r50.npz", arch="resnet50_rmac", seed=0, img_seed=1234, img_shape=np.array([4, 224, 224]),
         desc=d4, desc_b1=d1, img_seed_rect=99, img_shape_rect=np.array([2, 160, 224]), desc_rect=dr, **stats, **slices)

    # head options (rmac_resnet.py:24-31,61-66)
    xs = synth.make_images(2, 128, 128, seed=5)
    out = {}
    for tag, kw in [("max", dict(pooling="max")), ("avg", dict(pooling="avg")),
                    ("normfeat", dict(norm_features=True)), ("nofc", dict(without_fc=True)),
                    ("gemp2", dict(gemp=2))]:
        out["desc_" + tag] = ref_model("resnet50_rmac", seed=3, **kw)(xs).numpy()
    save("extract_r50_options.npz", arch="resnet50_rmac", seed=3, img_seed=5, img_shape=np.array([2, 128, 128]), **out)

    net = ref_model("resnet101_rmac", seed=1)
    x = synth.make_images(2, 224, 224, seed=77)
    save("extract_r101.npz", arch="resnet101_rmac", seed=1, img_seed=77, img_shape=np.array([2, 224, 224]),
         desc=net(x).numpy())


@torch.no_grad()
def gold_extract_extra():
    """resnet152_rmac (rmac_resnet.py:86-88) and the center_bias option (rmac_resnet.py:52-56)."""
    net = ref_model("resnet152_rmac", seed=4)
    x = synth.make_images(1, 96, 128, seed=41)
    out = {"desc_r152": net(x).numpy()}
    xs = synth.make_images(2, 128, 160, seed=6)
    for tag, b in (("cb05", 0.5), ("cb2", 2.0)):
        out["desc_" + tag] = ref_model("resnet50_rmac", seed=3, center_bias=b)(xs).numpy()
    save("extract_extra.npz", r152_seed=4, r152_img_seed=41, r152_img_shape=np.array([1, 96, 128]),
         cb_seed=3, cb_img_seed=6, cb_img_shape=np.array([2, 128, 160]), **out)


@torch.no_grad()
def gold_gem():
    r = np.random.RandomState(3)
    x = torch.from_numpy(r.standard_normal((2, 8, 5, 7)).astype(np.float32))
    out = {"x": x.numpy()}
    for p in (3.0, 2.5, 1.0):
        out["p%g" % p] = GeneralizedMeanPooling(p)(x).numpy().reshape(2, 8)
    save("gem.npz", **out)


def gold_pool():
    r = np.random.RandomState(4)
    xs = [torch.from_numpy(r.standard_normal((5, 16)).astype(np.float32)) for _ in range(3)]
    save("pool.npz", x0=xs[0].numpy(), x1=xs[1].numpy(), x2=xs[2].numpy(),
         mean=common.pool(xs, "mean").numpy(), gem3=common.pool(xs, "gem", 3).numpy(),
         gem2=common.pool(xs[:2], "gem", 2).numpy(), single=common.pool(xs[:1], "gem", 3).numpy())


def gold_whiten():
    out = {}
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        pca = synth.make_pca(64, seed=11, dtype=dt)
        X = synth._unit_rows(np.random.RandomState(12).standard_normal((10, 64))).astype(np.float32)
        out["X"] = X
        out["mean_" + tag], out["comp_" + tag], out["var_" + tag] = pca.mean_, pca.components_, pca.explained_variance_
        out["w_p025_" + tag] = common.whiten_features(X, pca, whitenp=0.25)
        out["w_p05_v32_m2_" + tag] = common.whiten_features(X, pca, whitenp=0.5, whitenv=32, whitenm=2.0)
        out["w_nol2_" + tag] = common.whiten_features(X, pca, l2norm=False, whitenp=0.25)
    save("whiten.npz", **out)


def gold_rank_ap():
    db, q, pos = synth.make_descriptor_db(400, 6, dim=32, n_pos=5, db_seed=21, q_seed=22)
    scores = common.matmul(q, db)
    gnd = synth.oxford_gt(pos, n_junk=3, n_db=400, seed=5)
    tmp = tempfile.mkdtemp()
    imlist = ["im%04d" % i for i in range(400)]
    with open(os.path.join(tmp, "gnd.pkl"), "wb") as f:
        pickle.dump({"imlist": imlist, "qimlist": imlist[:6], "gnd": gnd}, f)
    ds = ImageListRelevants(os.path.join(tmp, "gnd.pkl"), root=tmp)
    aps = np.array([ds.eval_query_AP(i, scores[i]) for i in range(6)])
    order = np.stack([np.argsort(scores[i])[::-1] for i in range(6)])
    # revisited (easy/hard) protocol, generic.py:210-224
    gnd2 = [{"bbx": g["bbx"], "easy": g["ok"][:2], "hard": g["ok"][2:], "junk": g["junk"]} for g in gnd]
    with open(os.path.join(tmp, "gnd2.pkl"), "wb") as f:
        pickle.dump({"imlist": imlist, "qimlist": imlist[:6], "gnd": gnd2}, f)
    ds2 = ImageListRelevants(os.path.join(tmp, "gnd2.pkl"), root=tmp)
    aps2 = [ds2.eval_query_AP(i, scores[i]) for i in range(6)]
    save("rank_ap.npz", n_db=400, n_q=6, dim=32, n_pos=5, db_seed=21, q_seed=22, n_junk=3, gt_seed=5,
         scores=scores, order=order, aps=aps,
         aps_easy=np.array([a["easy"] for a in aps2]), aps_medium=np.array([a["medium"] for a in aps2]),
         aps_hard=np.array([a["hard"] for a in aps2]))


def gold_aqe():
    db, q, pos = synth.make_descriptor_db(300, 5, dim=48, n_pos=4, db_seed=31, q_seed=32)
    out = {"n_db": 300, "n_q": 5, "dim": 48, "n_pos": 4, "db_seed": 31, "q_seed": 32}
    out["aqe_k2_a05"] = ref_test_dir.expand_descriptors(q, db=db, k=2, alpha=0.5)
    out["aqe_k3_a1"] = ref_test_dir.expand_descriptors(q, db=db, k=3, alpha=1)
    out["aqe_k0"] = ref_test_dir.expand_descriptors(q, db=db, k=0, alpha=1)
    small = db[:40].copy()
    small[1] = synth._unit_rows((small[0] + 0.3 * small[1])[None])[0]   # make sure each row has a positive neighbour
    sim = small @ small.T
    np.fill_diagonal(sim, 0)
    if (np.sort(sim, axis=1)[:, -2:] > 0).all():
        out["dba_in"] = small
        out["dba_k2_a1"] = ref_test_dir.expand_descriptors(small, db=None, k=2, alpha=1)
    save("aqe.npz", **out)


@torch.no_grad()
def gold_variants():
    """Model variants of SURVEY 8f-4 for which only the oracle exists so far: BasicBlock trunk (resnet18_rmac,
    rmac_resnet.py:74-76) and the FPN head (rmac_resnet_fpn.py:52-90), modes 1 and 0."""
    out = {}
    x = synth.make_images(2, 128, 160, seed=61)
    out["desc_r18"] = ref_model("resnet18_rmac", seed=5)(x).numpy()
    out["desc_r18_b1"] = ref_model("resnet18_rmac", seed=5)(x[:1]).numpy()
    out["desc_r50_fpn"] = ref_model("resnet50_fpn_rmac", seed=6, out_dim=2048)(x).numpy()
    out["desc_r50_fpn0"] = ref_model("resnet50_fpn_rmac", seed=6, out_dim=2048, mode=0)(x).numpy()
    out["desc_r18_fpn"] = ref_model("resnet18_fpn_rmac", seed=7, out_dim=512)(x).numpy()
    save("extract_variants.npz", img_seed=61, img_shape=np.array([2, 128, 160]), r18_seed=5, fpn_seed=6, r18_fpn_seed=7, **out)


def gold_labels():
    """Label-based AP of Dataset.eval_query_AP (dataset.py:83-92 -> sklearn, evaluation.py:41-43) on labelled image
    lists (generic.py:44-105): every image a query, and a separate query list."""
    from dirtorch.datasets.generic import ImageListLabels, ImageListLabelsQ
    r = np.random.RandomState(17)
    n, nq = 60, 7
    labels = ["c%d" % v for v in r.randint(0, 6, n)]
    labels[5] = "lonely"                                      # a class of one: AP = -1 for that query
    qlabels = ["c%d" % v for v in r.randint(0, 7, nq)]         # c6 never occurs in the database
    tmp = tempfile.mkdtemp()
    with open(os.path.join(tmp, "db.txt"), "w") as f:
        f.write("\n".join("im%03d.jpg %s" % (i, l) for i, l in enumerate(labels)) + "\n")
    with open(os.path.join(tmp, "q.txt"), "w") as f:
        f.write("\n".join("q%03d.jpg %s" % (i, l) for i, l in enumerate(qlabels)) + "\n")
    scores = r.standard_normal((n, n)).astype(np.float32)
    qscores = r.standard_normal((nq, n)).astype(np.float32)
    ds = ImageListLabels(os.path.join(tmp, "db.txt"), root=tmp)
    aps = np.array([ds.eval_query_AP(q, scores[q]) for q in range(n)], dtype=np.float64)
    dsq = ImageListLabelsQ(os.path.join(tmp, "db.txt"), os.path.join(tmp, "q.txt"), root=tmp)
    qaps = []
    for q in range(nq):
        try:
            qaps.append(dsq.eval_query_AP(q, qscores[q]))
        except KeyError:                                        # label absent from the database: c_relevant_idx[...] raises
            qaps.append(np.nan)
    save("label_ap.npz", labels=np.array(labels), qlabels=np.array(qlabels), scores=scores, qscores=qscores, aps=aps,
         qaps=np.array(qaps, dtype=np.float64), gt0=ds.get_query_groundtruth(0), nclass=ds.nclass, nclass_q=dsq.nclass)


TRF_CHAINS = ["Pad(70)", "PadSquare()", "PadSquare(60)", "PadSquare(100)", "CenterCrop(32)", "CenterCrop((20,40))",
              "Scale(48), CenterCrop(40)", "Scale(0.7), Pad(64, color=(0.5,0.5,0.5))", "Identity()",
              "CenterCrop(30, padding=4)", "Scale(0.5)", "Scale(1.4)", "Scale(40, largest=True)", "Scale((50, 30))"]
TRF_SIZES = [(50, 80), (80, 50), (64, 64), (33, 97)]


@torch.no_grad()
def gold_transforms():
    """The deterministic test-time transforms of dirtorch/utils/transforms.py (Scale, Pad, PadSquare, CenterCrop,
    Identity + the ToTensor/Normalize tail that transforms.create appends) on seeded images: a CRC of every output
    tensor; and the BATCHED extraction path of test_dir.extract_image_features (same_size=True, batch_size=4,
    test_dir.py:52-53,64-75) through a 'Scale(..), CenterCrop(..)' chain."""
    import zlib
    from PIL import Image
    from dirtorch.utils import transforms as RT
    r = np.random.RandomState(0)
    crcs, shapes = [], []
    for (h, w) in TRF_SIZES:
        img = Image.fromarray(r.randint(0, 256, (h, w, 3), dtype=np.uint8))
        for chain in TRF_CHAINS:
            t = RT.create(chain, to_tensor=True, mean=synth.RGB_MEANS, std=synth.RGB_STDS)(img).numpy()
            crcs.append(zlib.crc32(np.ascontiguousarray(t).tobytes()))
            shapes.append(t.shape)
    sys.path.append(REPO)
    sys.path.append(os.path.join(REPO, "tests"))
    import e2e_data
    from dirtorch.datasets.generic import ImageList
    root = tempfile.mkdtemp(prefix="dbroot_trf_")
    gnd, names, qn, sd = e2e_data.build(root)
    with open(os.path.join(root, "list.txt"), "w") as f:
        f.write("\n".join(names[:10]) + "\n")
    ds = ImageList(os.path.join(root, "list.txt"), os.path.join(root, "oxford5k", "jpg"))
    net = ref_model("resnet50_rmac", seed=0)
    net.iscuda = False
    net.preprocess = dict(mean=synth.RGB_MEANS, std=synth.RGB_STDS, input_size=224)
    chain = "Scale(140), CenterCrop(128)"
    d = ref_test_dir.extract_image_features(ds, chain, net, same_size=True, batch_size=4, iscuda=False, threads=2)
    d1 = ref_test_dir.extract_image_features(ds, chain, net, same_size=False, batch_size=4, iscuda=False, threads=2)
    print("batched vs batch-1 descriptors:", float((d - d1).abs().max()))
    save("transforms.npz", crcs=np.array(crcs, dtype=np.uint32), shapes=np.array(shapes, dtype=np.int32),
         batched_chain=np.array(chain), batched_desc=d.numpy(), n_images=10)


def gold_cli():
    for hard in (False, True):
        _gold_cli(hard)


def _gold_cli(hard):
    """The reference's OWN command lines, unmodified, on the synthetic Oxford-layout DB_ROOT of tests/e2e_data.py:
    `python -m dirtorch.extract_features` and `python -m dirtorch.test_dir` (CPU, --gpu -1), run as subprocesses
    from a neutral working directory with only /root/reference on PYTHONPATH."""
    import json
    import subprocess
    from sklearn.decomposition import PCA
    sys.path.append(REPO)                      # after /root/reference: `dirtorch` stays the reference's
    sys.path.append(os.path.join(REPO, "tests"))
    import e2e_data
    root = tempfile.mkdtemp(prefix="dbroot_cli_")
    gnd, names, qn, sd = e2e_data.build(root, hard=hard)
    env = dict(os.environ, PYTHONPATH=REF, DB_ROOT=root, TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD="1")
    ckpt = os.path.join(root, "ckpt.pt")
    opts = dict(arch="resnet50_rmac", out_dim=2048, pooling="gem", gemp=3)

    def run(mod, *argv):
        r = subprocess.run([sys.executable, "-m", mod] + list(argv), cwd=root, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout

    # 1. raw descriptors through the reference's extract_features CLI (checkpoint without PCA yet)
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}, "model_options": opts}, ckpt)
    with open(os.path.join(root, "list.txt"), "w") as f:
        f.write("\n".join(names) + "\n")
    run("dirtorch.extract_features", "--dataset", 'ImageList("%s/list.txt", "%s/oxford5k/jpg")' % (root, root),
        "--checkpoint", ckpt, "--output", os.path.join(root, "out", "feats.npy"), "--gpu", "-1", "--threads", "2")
    D = np.load(os.path.join(root, "out", "feats.npy"))
    # 2. PCA learnt on them, stored the way the released checkpoints do (ckpt['pca'][name]), test_dir CLI
    pca = PCA(n_components=32, whiten=True).fit(D)
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}, "model_options": opts,
                "pca": {"Landmarks_clean": pca}}, ckpt)
    common_args = ["--dataset", "Oxford5K", "--checkpoint", ckpt, "--whiten", "Landmarks_clean", "--whitenp", "0.25",
                   "--gpu", "-1", "--threads", "2"]
    out1 = run("dirtorch.test_dir", *common_args, "--save-feats", os.path.join(root, "saved"),
               "--out-json", os.path.join(root, "res1.json"))
    out2 = run("dirtorch.test_dir", *common_args, "--load-feats", os.path.join(root, "saved"), "--aqe", "2", "1",
               "--out-json", os.path.join(root, "res2.json"))
    out3 = run("dirtorch.test_dir", *common_args, "--load-feats", os.path.join(root, "saved"), "--adba", "2", "1",
               "--out-json", os.path.join(root, "res3.json"))
    res = [json.load(open(os.path.join(root, "res%d.json" % i)))["Oxford5K"] for i in (1, 2, 3)]
    lines = [[l for l in o.splitlines() if l.startswith(" * ")] for o in (out1, out2, out3)]
    # 3. per-query APs (eval_model(detailed=True) cannot be printed by the reference's own CLI)
    from dirtorch.datasets.generic import ImageListRelevants
    ds = ImageListRelevants(os.path.join(root, "oxford5k", "gnd_oxford5k.pkl"), root=os.path.join(root, "oxford5k"))   # = Oxford5K(), oxford.py:8-11
    net = ref_test_dir.load_model(ckpt, False)
    net.pca = net.pca["Landmarks_clean"]
    ref_test_dir.args = SimpleNamespace(aqe=None, adba=None)
    det = ref_test_dir.eval_model(ds, net, "", detailed=True, whiten=dict(whitenp=0.25, whitenv=None, whitenm=1.0),
                                  load_feats=os.path.join(root, "saved"), threads=2)
    saved = np.load(os.path.join(root, "saved", "feats.bdescs.npy"))
    W = common.whiten_features(saved, net.pca, whitenp=0.25)
    # The fitted PCA in 13 KB instead of 270 KB: its components lie in the row space of the centred descriptors,
    # components_ = coeff @ (D - mean(D)); the test rebuilds them from the oracle's descriptors (a PCA *refit* there
    # would be ill-conditioned in the trailing components and is sklearn's business, not the path's).
    Dc = D.astype(np.float64) - D.astype(np.float64).mean(0)
    coeff = pca.components_.astype(np.float64) @ np.linalg.pinv(Dc, rcond=1e-10)
    err_c, err_m = np.abs(coeff @ Dc - pca.components_).max(), np.abs(pca.mean_ - D.mean(0)).max()
    print('PCA coefficient reconstruction error', err_c, err_m)
    assert err_c < 1e-5 and err_m < 1e-6
    extra = {}
    if hard:
        # multi-scale protocol of the paper through the same command line: three transform chains, GeM pooling of
        # the per-scale descriptors (common.pool, common.py:41-55), F.normalize, whitening, ranking
        out4 = run("dirtorch.test_dir", *common_args, "--trfs", "Scale(0.7)", "", "Scale(1.4)", "--pooling", "gem",
                   "--gemp", "3", "--save-feats", os.path.join(root, "saved_ms"), "--out-json", os.path.join(root, "res4.json"))
        saved_ms = np.load(os.path.join(root, "saved_ms", "feats.bdescs.npy"))
        det_ms = ref_test_dir.eval_model(ds, net, "", detailed=True, whiten=dict(whitenp=0.25, whitenv=None, whitenm=1.0),
                                         load_feats=os.path.join(root, "saved_ms"), threads=2)
        extra = dict(ms_mAP=json.load(open(os.path.join(root, "res4.json")))["Oxford5K"]["mAP"], ms_APs=np.array(det_ms["APs"]),
                     ms_desc_head=saved_ms[:3], ms_console=np.array([l for l in out4.splitlines() if l.startswith(" * ")]))
    if hard:
        # revisited protocol (easy / medium / hard, generic.py:150-170,210-224) through `--dataset ROxford5K`
        gnd_r = [{"bbx": g_["bbx"], "easy": [g_["ok"][0], g_["ok"][2]], "hard": [g_["ok"][1]], "junk": g_["junk"]} for g_ in gnd]
        gnd_r[3]["easy"], gnd_r[3]["hard"] = gnd[3]["ok"], []          # no hard positive: AP-hard = -1, left out of the mean
        with open(os.path.join(root, "oxford5k", "gnd_roxford5k.pkl"), "wb") as f:
            pickle.dump({"imlist": names, "qimlist": [names[i] for i in qn], "gnd": gnd_r}, f)
        out5 = run("dirtorch.test_dir", "--dataset", "ROxford5K", *common_args[2:], "--load-feats", os.path.join(root, "saved"),
                   "--out-json", os.path.join(root, "res5.json"))
        r5 = json.load(open(os.path.join(root, "res5.json")))["ROxford5K"]
        extra.update(rox_easy=r5["mAP-easy"], rox_medium=r5["mAP-medium"], rox_hard=r5["mAP-hard"],
                     rox_console=np.array([l for l in out5.splitlines() if l.startswith(" * ")]))
    save("cli_hard.npz" if hard else "cli_easy.npz", **extra, pca_coeff=coeff, pca_var=pca.explained_variance_, n_images=len(names), queries=np.array(qn), desc_head=D[:3], desc_norms=np.linalg.norm(D, axis=1),
         saved_equals_extract=np.array(np.array_equal(saved, D)), whitened=W,
         mAP=res[0]["mAP"], mAP_aqe_k2_a1=res[1]["mAP"], mAP_adba_k2_a1=res[2]["mAP"], APs=np.array(det["APs"]),
         console=np.array([l for ls in lines for l in ls]))


if __name__ == "__main__":
    if sys.argv[1:] == ["extra"]:               # only the files added after the first batch
        gold_extract_extra()
        sys.exit(0)
    if sys.argv[1:] == ["cli"]:
        gold_cli()
        sys.exit(0)
    if sys.argv[1:] == ["variants"]:
        gold_variants()
        sys.exit(0)
    if sys.argv[1:] == ["labels"]:
        gold_labels()
        sys.exit(0)
    if sys.argv[1:] == ["transforms"]:
        gold_transforms()
        sys.exit(0)
    gold_gem()
    gold_pool()
    gold_whiten()
    gold_rank_ap()
    gold_aqe()
    gold_extract()
    gold_extract_extra()
    gold_cli()
    gold_labels()
    gold_transforms()
    gold_variants()
