"""The extraction / evaluation drivers of the reference, on the B200 path.

Function names, keyword arguments and console output follow ``dirtorch/test_dir.py`` (``expand_descriptors`` :24-44,
``extract_image_features`` :47-94, ``eval_model`` :97-180, ``load_model`` :183-191, CLI :194-259) and
``dirtorch/extract_features.py`` (``extract_features`` :26-68, CLI :82-124).  What runs underneath:
``net(imgs)`` is the libdirb200 network, multi-scale pooling / whitening / similarity / alpha-QE are libdirb200
kernels; only AP bookkeeping (datasets.py) stays on the host, as in the reference.
"""
from __future__ import annotations

import json
import os
import os.path as osp

import numpy as np
import torch
import tqdm

from . import common, datasets, nets, ops
from .common import matmul, pool, tonumpy
from .loader import get_loader
from .store import DescriptorStore


def mkdir(fname, isfile="auto"):
    if isfile == "auto":
        isfile = bool(os.path.splitext(fname)[1])
    directory = os.path.split(fname)[0] if isfile else fname
    if directory and not os.path.isdir(directory):
        os.makedirs(directory)


def _dev_f32(x):
    if isinstance(x, torch.Tensor):
        return x.to("cuda", torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32))).cuda()


def _drop_self(scores, idx, own_rows, k):
    """Neighbour lists (best first, one more entry than needed) of database rows queried against their own
    database -> the first k entries that are not the row itself, order preserved.  A row need not be in its own
    list (exact duplicates with a lower index rank before it); then the list is just cut to k."""
    own = np.asarray(own_rows)[:, None]
    keep = np.argsort(idx == own, axis=1, kind="stable")[:, :k]                              # non-self entries first
    s, i = np.take_along_axis(scores, keep, 1).copy(), np.take_along_axis(idx, keep, 1).copy()
    left = i == own                                   # only when the database has no more than k rows: no neighbour
    i[left], s[left] = -1, 0.0                        # (index -1 = skipped by dirb200_aqe_expand)
    return s, i


def expand_descriptors(descs, db=None, alpha=0, k=0, q_block=4096, to_numpy=True):
    """alpha query expansion (db given) / database augmentation (db=None): test_dir.py:24-44.
    q' = normalize(mean([q] + [db_j * sim_ij^alpha for the k nearest j])).  Returns a host ndarray like the
    reference (to_numpy=False, extension: the device tensor).  q_block (extension): queries searched per pass."""
    assert k >= 0 and alpha >= 0, "k and alpha must be non-negative"
    if k == 0:
        return descs
    if k + (db is None) > 1024:
        # the reference expands with any k (a full argsort per query); the exact top-k search returns at most 1024 rows
        raise ValueError("query expansion / database augmentation with more than %d neighbours is not supported "
                         "(dirb200_index_search returns at most 1024 rows per query); got k=%d" % (1024 - (db is None), k))
    dim = np.shape(descs)[1]
    pad = (-dim) % 64                      # the tensor-core search wants D % 64 == 0; zero columns change no score

    def _padded(x):
        x = _dev_f32(x)
        return torch.nn.functional.pad(x, (0, pad)).contiguous() if pad else x

    q = _padded(descs)
    d = _padded(db) if db is not None else q
    index = ops.Index(d)
    out = torch.empty_like(q)
    # Queries go through in blocks so that database augmentation (an N x N self-search) needs O(block) scratch.
    for c0 in range(0, q.shape[0], q_block):
        qc = q[c0:c0 + q_block].contiguous()
        if db is not None:
            s, i = index.search(qc, min(k, d.shape[0]))
        else:
            # the reference zeroes the diagonal of the self-similarity (test_dir.py:33-34): drop each row from its
            # own neighbour list
            s1, i1 = index.search(qc, min(k + 1, d.shape[0]))
            s1, i1 = _drop_self(s1.cpu().numpy(), i1.cpu().numpy(), np.arange(c0, c0 + qc.shape[0]), k)
            s, i = torch.from_numpy(s1).cuda(), torch.from_numpy(i1).cuda()
        out[c0:c0 + qc.shape[0]] = ops.aqe_expand(qc, d, i.contiguous(), s.contiguous(), float(alpha))
    out = out[:, :dim].contiguous()
    return out.cpu().numpy() if to_numpy else out


_GPU_CHAIN = None


def _gpu_chain_scale(transforms):
    """Transform chains the GPU preprocessing path reproduces bit for bit: '' (ToTensor + Normalize only) and
    'Scale(<float>)' (transforms.py:133-185 with PIL bilinear) -> the scale factor (1.0 for ''), else None."""
    import re
    global _GPU_CHAIN
    if _GPU_CHAIN is None:
        _GPU_CHAIN = re.compile(r"^\s*(?:Scale\(\s*([0-9]*\.[0-9]+|[0-9]+\.[0-9]*)\s*\))?\s*$")
    m = _GPU_CHAIN.match(transforms or "")
    if not m:
        return None
    s = float(m.group(1)) if m.group(1) else 1.0
    return s if 0 < s <= 4 else None


class _RawImages(torch.utils.data.Dataset):
    """Decoded images as uint8 HWC tensors (the decode stays on CPU workers; everything after it runs on the GPU)."""

    def __init__(self, dataset):
        self.dataset = dataset

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, i):
        return torch.from_numpy(np.array(self.dataset.get_image(i).convert("RGB"), dtype=np.uint8))


def extract_image_features(dataset, transforms, net, ret_imgs=False, same_size=False, flip=None,
                           desc="Extract feats...", iscuda=True, threads=8, batch_size=8, gpu_preprocess=True):
    """Descriptors of every image of `dataset`, (N, D) on the GPU (test_dir.py:47-94).
    gpu_preprocess (extension): when the chain is '' or 'Scale(<float>)' and images go through one by one, resize
    (byte-identical to PIL bilinear), ToTensor and Normalize run on the GPU from the decoded uint8 pixels
    (dirb200_resize_bilinear_u8 + dirb200_net_forward_u8) - bit-identical descriptors, no PIL resize / float
    conversion on the host and 4x fewer bytes over PCIe."""
    if not same_size:
        batch_size = 1
    scale = _gpu_chain_scale(transforms) if (gpu_preprocess and not same_size and not ret_imgs and hasattr(net, "forward_u8")) else None
    if hasattr(net, "eval"):
        net.eval()
    if scale is not None:
        loader = torch.utils.data.DataLoader(_RawImages(dataset), batch_size=None, shuffle=False,
                                             num_workers=max(0, threads if threads > 1 else 0), pin_memory=True)
        img_feats = []
        for img in tqdm.tqdm(loader, desc, total=len(dataset)):
            x = img.cuda(non_blocking=True)
            if flip and flip.pop(0):
                x = x.flip(1)                            # horizontal flip (imgs[i].flip(2) on CHW, test_dir.py:70-72)
            h, w = int(x.shape[0]), int(x.shape[1])
            oh, ow = int(0.5 + scale * h), int(0.5 + scale * w)          # Scale.get_params, transforms.py:168
            x = x.unsqueeze(0).contiguous()
            if (oh, ow) != (h, w) and min(oh, ow) != min(h, w):          # Scale.__call__: can_upscale / can_downscale
                x = ops.resize_bilinear_u8(x, (oh, ow))
            d = net.forward_u8(x)
            img_feats.append(d.unsqueeze(0) if d.dim() == 1 else d)
        return torch.cat(img_feats, dim=0)
    loader = get_loader(dataset, trf_chain=transforms, preprocess=net.preprocess, iscuda=iscuda, output=["img"],
                        batch_size=batch_size, threads=threads, shuffle=False)
    tocpu = (lambda x: x.cpu()) if ret_imgs == "cpu" else (lambda x: x)
    img_feats, trf_images = [], []
    for inputs in tqdm.tqdm(loader, desc, total=1 + (len(dataset) - 1) // batch_size):
        imgs = inputs[0]
        for i in range(len(imgs)):
            if flip and flip.pop(0):
                imgs[i] = imgs[i].flip(2)
        imgs = common.variables(inputs[:1], net.iscuda)[0]
        d = net(imgs)
        if ret_imgs:
            trf_images.append(tocpu(imgs))
        if d.dim() == 1:
            d = d.unsqueeze(0)
        img_feats.append(d)
    img_feats = torch.cat(img_feats, dim=0)
    if ret_imgs:
        if same_size:
            trf_images = torch.cat(trf_images, dim=0)
        return trf_images, img_feats
    return img_feats


def _pool_and_normalize(descs, pooling, gemp):
    """pool() over transform chains followed by F.normalize(p=2, dim=1): test_dir.py:121-122."""
    if len(descs) == 1:
        return ops.l2_normalize(_dev_f32(descs[0]))
    if pooling not in ("mean", "gem"):
        raise ValueError("Bad pooling mode: " + str(pooling))
    return ops.pool_scales([_dev_f32(d) for d in descs], pooling, gemp, l2=True)


def _aps_from_topk(db, qdescs, bdescs, k):
    """Per-query AP from the exact top-k ranking (dirb200_index_search); dense score row only where needed."""
    dim = qdescs.shape[1]
    pad = (-dim) % 64
    f = lambda a: torch.nn.functional.pad(_dev_f32(a), (0, pad)).contiguous() if pad else _dev_f32(a)
    qd, bd = f(qdescs), f(bdescs)
    k = min(k, bd.shape[0], 1024)
    _, idx = ops.Index(bd).search(qd, k)
    idx = idx.cpu().numpy()
    aps = []
    for q in range(qd.shape[0]):
        ap = db.eval_query_AP_from_ranking(q, idx[q])
        incomplete = ap is None or (isinstance(ap, dict) and any(v is None for v in ap.values()))
        if incomplete:
            row = ops.scores_exact(qd[q:q + 1].contiguous(), bd).cpu().numpy()[0]
            ap = db.eval_query_AP(q, row)
        aps.append(ap)
    return aps


def _aps_from_counts(db, qdescs, bdescs):
    """Per-query AP without the Q x N score matrix and without any host-side ranking: the GPU returns the exact scores
    of the labelled rows and, for every positive, the number of database rows ranking before it
    (dirb200_index_rank_count); datasets.eval_query_AP_from_counts turns that into the AP of generic.py:196-224."""
    dim = qdescs.shape[1]
    pad = (-dim) % 64
    f = lambda a: torch.nn.functional.pad(_dev_f32(a), (0, pad)).contiguous() if pad else _dev_f32(a)
    qd, bd = f(qdescs), f(bdescs)
    offs, rows, flags = [0], [], []
    for q in range(qd.shape[0]):
        r, fl = db.rank_targets(q)
        rows.append(r)
        flags.append(fl)
        offs.append(offs[-1] + len(r))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    flags = np.concatenate(flags) if flags else np.zeros(0, np.uint8)
    sc, above = ops.Index(bd).rank_counts(qd, np.array(offs, np.int32), rows, flags)
    sc, above = sc.cpu().numpy(), above.cpu().numpy()
    return [db.eval_query_AP_from_counts(q, rows[offs[q]:offs[q + 1]], sc[offs[q]:offs[q + 1]], above[offs[q]:offs[q + 1]])
            for q in range(qd.shape[0])]


def eval_model(db, net, trfs, pooling="mean", gemp=3, detailed=False, whiten=None, aqe=None, adba=None, threads=8,
               batch_size=16, save_feats=None, load_feats=None, dbg=(), rank_topk=0, dense_scores=False):
    """Evaluate a network on a retrieval dataset (test_dir.py:97-180).  aqe / adba: dict(k=..., alpha=...).
    Ranking: datasets with relevance lists (Oxford/Paris style) are graded from GPU rank counts - no Q x N score
    matrix, no host argsort; dense_scores=True (extension) forces the literal matmul + per-query sort of the
    reference, which label datasets always use; rank_topk > 0 (extension) grades from the exact top-k lists."""
    print("\n>> Evaluation...")
    query_db = db.get_query_db()
    bdescs, qdescs = [], []
    if not load_feats:
        trfs_list = [trfs] if isinstance(trfs, str) else trfs
        for trf in trfs_list:
            kw = dict(iscuda=net.iscuda, threads=threads, batch_size=batch_size, same_size="Pad" in trf or "Crop" in trf)
            bdescs.append(extract_image_features(db, trf, net, desc="DB", **kw))
            qdescs.append(bdescs[-1] if db is query_db else extract_image_features(query_db, trf, net, desc="query", **kw))
        bdescs = _pool_and_normalize(bdescs, pooling, gemp)
        qdescs = _pool_and_normalize(qdescs, pooling, gemp)
    elif DescriptorStore.is_store(os.path.join(load_feats, "bdescs")):
        # extension: row-sharded descriptor store (store.py) instead of one .npy per side
        bdescs = DescriptorStore(os.path.join(load_feats, "bdescs")).read_all().astype(np.float32, copy=False)
        qdescs = (DescriptorStore(os.path.join(load_feats, "qdescs")).read_all().astype(np.float32, copy=False)
                  if query_db is not db else bdescs)
    else:
        bdescs = np.load(os.path.join(load_feats, "feats.bdescs.npy"))
        qdescs = np.load(os.path.join(load_feats, "feats.qdescs.npy")) if query_db is not db else bdescs
    if save_feats:
        mkdir(save_feats, isfile=False)
        np.save(os.path.join(save_feats, "feats.bdescs.npy"), tonumpy(bdescs))
        if query_db is not db:
            np.save(os.path.join(save_feats, "feats.qdescs.npy"), tonumpy(qdescs))
    # From here on the descriptors stay in HBM (the reference hops GPU -> numpy -> GPU -> numpy between the steps,
    # test_dir.py:136-145): whitening, database augmentation, query expansion and the grading all take device tensors.
    same = qdescs is bdescs
    if whiten is not None:
        bdescs = common.whiten_features_gpu(bdescs, net.pca, **whiten)
        qdescs = bdescs if same else common.whiten_features_gpu(qdescs, net.pca, **whiten)
    if adba is not None:
        # (the reference augments the database rows only: queries drawn from the database keep their un-augmented rows)
        bdescs = expand_descriptors(bdescs, to_numpy=False, **adba)
    if aqe is not None:
        qdescs = expand_descriptors(qdescs, db=bdescs, to_numpy=False, **aqe)
    res = {}
    qn, bn = _dev_f32(qdescs), _dev_f32(bdescs)
    unit = bool(float((qn.norm(dim=1) - 1).abs().max()) < 1e-3 and float((bn.norm(dim=1) - 1).abs().max()) < 1e-3)
    if rank_topk and hasattr(db, "eval_query_AP_from_ranking"):
        # Large databases: rank with the exact top-k engine instead of materialising the Q x N score matrix; a
        # query whose positives are not all inside the top-k falls back to its exact dense score row.
        aps = _aps_from_topk(db, qn, bn, int(rank_topk))
        scores = None
    elif not dense_scores and unit and hasattr(db, "eval_query_AP_from_counts"):
        # default for relevance-list datasets: exact rank counts of the labelled rows from one tensor-core pass
        # (the error bound of the fp16 pass assumes unit-norm rows, which pooling / whitening / QE all produce)
        aps = _aps_from_counts(db, qn, bn)
        scores = None
    else:
        scores = matmul(qn, bn)
        aps = None
    del bdescs, qdescs
    try:
        if aps is None:
            aps = [db.eval_query_AP(q, s) for q, s in enumerate(tqdm.tqdm(scores, desc="AP"))]
        if not isinstance(aps[0], dict):
            aps = [float(e) for e in aps]
            if detailed:
                res["APs"] = aps
            res["mAP"] = float(np.mean([e for e in aps if e >= 0]))      # AP = -1: query without relevants
        else:
            for mode in aps[0].keys():
                apst = [float(e[mode]) for e in aps]
                if detailed:
                    res["APs-" + mode] = apst
                res["mAP-" + mode] = float(np.mean([e for e in apst if e >= 0]))
    except NotImplementedError:
        print(" AP not implemented!")
    try:
        if scores is None:
            raise NotImplementedError()
        tops = [db.eval_query_top(q, s) for q, s in enumerate(tqdm.tqdm(scores, desc="top1"))]
        if detailed:
            res["tops"] = tops
        for k in tops[0]:
            res["top%d" % k] = float(np.mean([top[k] for top in tops]))
    except NotImplementedError:
        pass
    return res


def extract_features(db, net, trfs, pooling="mean", gemp=3, detailed=False, whiten=None, threads=8, batch_size=16,
                     output=None, dbg=()):
    """Extract (and optionally whiten) descriptors and save them as .npy (extract_features.py:26-68)."""
    print("\n>> Extracting features...")
    try:
        query_db = db.get_query_db()
    except NotImplementedError:
        query_db = None
    bdescs, qdescs = [], []
    trfs_list = [trfs] if isinstance(trfs, str) else trfs
    for trf in trfs_list:
        kw = dict(iscuda=net.iscuda, threads=threads, batch_size=batch_size, same_size="Pad" in trf or "Crop" in trf)
        bdescs.append(extract_image_features(db, trf, net, desc="DB", **kw))
        if query_db is not None:
            qdescs.append(bdescs[-1] if db is query_db else extract_image_features(query_db, trf, net, desc="query", **kw))
    bdescs = tonumpy(_pool_and_normalize(bdescs, pooling, gemp))
    if query_db is not None:
        qdescs = tonumpy(_pool_and_normalize(qdescs, pooling, gemp))
    if whiten is not None:
        bdescs = common.whiten_features(bdescs, net.pca, **whiten)
        if query_db is not None:
            qdescs = common.whiten_features(qdescs, net.pca, **whiten)
    mkdir(output, isfile=True)
    if query_db is db or query_db is None:
        np.save(output, bdescs)
    else:
        o = osp.splitext(output)
        np.save(o[0] + ".qdescs" + o[1], qdescs)
        np.save(o[0] + ".dbdescs" + o[1], bdescs)
    print("Features extracted.")


class _RowRange:
    """The images [start, end) of a dataset, seen as a dataset (what one rank of a sharded extraction owns)."""

    def __init__(self, dataset, start, end):
        self.dataset, self.start, self.nimg = dataset, int(start), int(end) - int(start)

    def __len__(self):
        return self.nimg

    def get_key(self, i):
        return self.dataset.get_key(self.start + i)

    def get_filename(self, i, root=None):
        return self.dataset.get_filename(self.start + i, root=root)

    def get_image(self, i, resize=None):
        return self.dataset.get_image(self.start + i, resize=resize)


def extract_to_store(db, net, trfs, store_path, pooling="mean", gemp=3, whiten=None, threads=8, batch_size=16,
                     dtype=np.float32, group=None):
    """Multi-GPU extraction (extension; the reference's analogue is nn.DataParallel, common.py:155).  Run under
    torchrun, one process per GPU: rank r extracts the contiguous image range ``shard_rows(len(db), world, r)`` with
    its own network handle, pools over the transform chains, normalises and optionally whitens exactly as
    ``extract_features`` does, and writes its rows as one shard of a descriptor store (store.py); rank 0 writes the
    manifest.  Images are independent, so there is no data-path collective.  Returns the opened store; each rank can
    then search its rows with ``ShardedIndex.from_store``."""
    import torch.distributed as tdist
    from .dist import shard_rows
    from .store import write_distributed
    on = tdist.is_available() and tdist.is_initialized()
    rank, world = (tdist.get_rank(group), tdist.get_world_size(group)) if on else (0, 1)
    start, end = shard_rows(len(db), world, rank)
    part = _RowRange(db, start, end)
    trfs_list = [trfs] if isinstance(trfs, str) else trfs
    dim = net.descriptor_dim if hasattr(net, "descriptor_dim") else net.feat_dim
    if len(part):
        descs = []
        for trf in trfs_list:
            descs.append(extract_image_features(part, trf, net, desc="DB[%d/%d]" % (rank, world), iscuda=net.iscuda,
                                                threads=threads, batch_size=batch_size,
                                                same_size="Pad" in trf or "Crop" in trf))
        rows = tonumpy(_pool_and_normalize(descs, pooling, gemp))
        if whiten is not None:
            rows = common.whiten_features(rows, net.pca, **whiten)
    else:                                                     # more ranks than images: an empty shard with the
        if whiten is not None:                                # dimension the other ranks' rows have after whitening
            ncomp = np.asarray(net.pca.components_).shape[0]
            dim = min(ncomp, whiten.get("whitenv") or ncomp)
        rows = np.zeros((0, dim), np.float32)
    meta = dict(arch=getattr(net, "arch", ""), trfs=list(trfs_list), pooling=pooling, gemp=gemp, whiten=whiten or {},
                n_images=len(db))
    return write_distributed(store_path, rows, group=group, dtype=dtype, meta=meta)


def load_model(path, iscuda):
    """Build the network from a checkpoint dict (test_dir.py:183-191)."""
    checkpoint = common.load_checkpoint(path, iscuda)
    net = nets.create_model(pretrained="", **checkpoint["model_options"])
    net = common.switch_model_to_cuda(net, iscuda, checkpoint)
    net.load_state_dict(checkpoint["state_dict"])
    net.preprocess = checkpoint.get("preprocess", net.preprocess)
    if "pca" in checkpoint:
        net.pca = checkpoint.get("pca")
    return net


def _common_args(parser, whiten_default, whitenp_default):
    parser.add_argument("--dataset", "-d", type=str, required=True, help="Command to load dataset")
    parser.add_argument("--checkpoint", type=str, required=True, help="path to weights")
    parser.add_argument("--trfs", type=str, required=False, default="", nargs="+", help="test transforms (can be several)")
    parser.add_argument("--pooling", type=str, default="gem", help="pooling scheme if several trf chains")
    parser.add_argument("--gemp", type=int, default=3, help="GeM pooling power")
    parser.add_argument("--out-json", type=str, default="", help="path to output json")
    parser.add_argument("--detailed", action="store_true", help="return detailed evaluation")
    parser.add_argument("--threads", type=int, default=8, help="number of thread workers")
    parser.add_argument("--dbg", default=(), nargs="*", help="debugging options")
    parser.add_argument("--whiten", type=str, default=whiten_default, help="applies whitening")
    parser.add_argument("--whitenp", type=float, default=whitenp_default, help="whitening power")
    parser.add_argument("--whitenv", type=int, default=None, help="number of components, default is None (all)")
    parser.add_argument("--whitenm", type=float, default=1.0, help="whitening multiplier")


def _select_pca(net, args):
    if args.whiten:
        net.pca = net.pca[args.whiten]
        return {"whitenp": args.whitenp, "whitenv": args.whitenv, "whitenm": args.whitenm}
    net.pca = None
    return None


def test_dir_main(argv=None):
    """python -m dirtorch.test_dir (test_dir.py:194-259).  --aqe / --adba also accept a float alpha."""
    import argparse
    parser = argparse.ArgumentParser(description="Evaluate a model")
    _common_args(parser, whiten_default="Landmarks_clean", whitenp_default=0.25)
    parser.add_argument("--save-feats", type=str, default="", help="path to output features")
    parser.add_argument("--load-feats", type=str, default="", help="path to load features from")
    parser.add_argument("--gpu", type=int, default=0, nargs="+", help="GPU ids")
    parser.add_argument("--aqe", type=float, nargs="+", help="alpha-query expansion parameters: k alpha")
    parser.add_argument("--adba", type=float, nargs="+", help="alpha-database augmentation parameters: k alpha")
    parser.add_argument("--rank-topk", type=int, default=0,
                        help="(extension) rank with the exact top-k search (k <= 1024) instead of the dense score matrix")
    parser.add_argument("--dense-scores", action="store_true",
                        help="(extension) force the reference's literal matmul + per-query sort instead of GPU rank counts")
    args = parser.parse_args(argv)
    args.iscuda = common.torch_set_gpu(args.gpu)
    aqe = {"k": int(args.aqe[0]), "alpha": args.aqe[1]} if args.aqe is not None else None
    adba = {"k": int(args.adba[0]), "alpha": args.adba[1]} if args.adba is not None else None
    dataset = datasets.create(args.dataset)
    print("Test dataset:", dataset)
    net = load_model(args.checkpoint, args.iscuda)
    whiten = _select_pca(net, args)
    res = eval_model(dataset, net, args.trfs, pooling=args.pooling, gemp=args.gemp, detailed=args.detailed,
                     threads=args.threads, dbg=args.dbg, whiten=whiten, aqe=aqe, adba=adba,
                     save_feats=args.save_feats, load_feats=args.load_feats, rank_topk=args.rank_topk,
                     dense_scores=args.dense_scores)
    # (the reference's '%s = %g' line raises on the list-valued entries that --detailed adds; print scalars only)
    print(" * " + "\n * ".join(["%s = %g" % p for p in res.items() if isinstance(p[1], (int, float))]))
    if args.out_json:
        try:
            data = json.load(open(args.out_json))
        except IOError:
            data = {}
        data[args.dataset] = res
        mkdir(args.out_json)
        open(args.out_json, "w").write(json.dumps(data, indent=1))
        print("saved to " + args.out_json)
    return res


def extract_features_main(argv=None):
    """python -m dirtorch.extract_features (extract_features.py:82-124)."""
    import argparse
    parser = argparse.ArgumentParser(description="Extract features")
    _common_args(parser, whiten_default=None, whitenp_default=0.5)
    parser.add_argument("--output", type=str, default="", help="path to output features")
    parser.add_argument("--gpu", type=int, nargs="+", help="GPU ids")
    args = parser.parse_args(argv)
    args.iscuda = common.torch_set_gpu(args.gpu if args.gpu is not None else [0])
    dataset = datasets.create(args.dataset)
    print("Dataset:", dataset)
    net = load_model(args.checkpoint, args.iscuda)
    whiten = _select_pca(net, args)
    extract_features(dataset, net, args.trfs, pooling=args.pooling, gemp=args.gemp, detailed=args.detailed,
                     threads=args.threads, dbg=args.dbg, whiten=whiten, output=args.output)
