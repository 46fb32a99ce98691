"""Seeded synthetic inputs for the descriptor-extraction and retrieval path.

There is no network access for checkpoints or datasets, so every test, the
smoke run and the benchmark build their inputs here, from seeds only:

* ``make_state_dict``   - a ResNet-50/101 + GeM + FC state dict with the key
  names and shapes of the reference network (``dirtorch/nets/backbones/resnet.py:54-65,
  115-141``, ``dirtorch/nets/rmac_resnet.py:34``, ``dirtorch/nets/layers/pooling.py:54``),
  with non-trivial BatchNorm statistics and a damped last BN in each block so
  that activations stay O(1) through 33 residual blocks (the reference's own
  random init grows to ~4e5 by layer4).
* ``make_images``       - normalised NCHW fp32 image batches.
* ``make_pca``          - a PCA-whitening parameter set (``mean_, components_,
  explained_variance_``) with the attribute names ``dirtorch/utils/common.py:221-232`` reads.
* ``make_descriptor_db``- unit-norm descriptor database + queries with planted
  positives and Oxford-style ground truth (``dirtorch/datasets/generic.py:134-145``).

Everything is generated tensor by tensor from ``numpy.random.RandomState`` seeded
by (seed, tensor name), so the same call reproduces the same values on any box.
"""
from __future__ import annotations

import zlib
from types import SimpleNamespace

import numpy as np
import torch

BLOCKS = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3], "resnet152": [3, 8, 36, 3]}
RGB_MEANS = [0.485, 0.456, 0.406]
RGB_STDS = [0.229, 0.224, 0.225]


def _rs(seed: int, name: str) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


def _conv(seed, name, cout, cin, k):
    w = _rs(seed, name).standard_normal((cout, cin, k, k)).astype(np.float32)
    w *= np.float32(np.sqrt(2.0 / (cin * k * k)))
    return torch.from_numpy(w)


def _bn(seed, name, c, gamma_scale=1.0):
    r = _rs(seed, name)
    return {
        name + ".weight": torch.from_numpy((gamma_scale * r.uniform(0.7, 1.3, c)).astype(np.float32)),
        name + ".bias": torch.from_numpy((0.1 * r.standard_normal(c)).astype(np.float32)),
        name + ".running_mean": torch.from_numpy((0.1 * r.standard_normal(c)).astype(np.float32)),
        name + ".running_var": torch.from_numpy(r.uniform(0.7, 1.3, c).astype(np.float32)),
        name + ".num_batches_tracked": torch.tensor(0, dtype=torch.long),
    }


def arch_to_trunk(arch: str) -> str:
    """'resnet101_rmac' -> 'resnet101'."""
    trunk = arch.split("_")[0]
    if trunk not in BLOCKS:
        raise NameError("unknown model architecture '%s'" % arch)
    return trunk


BASIC_BLOCKS = {"resnet18": [2, 2, 2, 2]}


def make_state_dict(arch: str = "resnet50_rmac", seed: int = 0, out_dim: int = 2048,
                    gemp: float = 3.0, res_gamma: float = 0.3) -> dict:
    """Reference-keyed state dict: Bottleneck (resnet50/101/152) or BasicBlock (resnet18) trunk + GeM(p) + FC head;
    ``*_fpn_rmac`` / ``*_fpn0_rmac`` names add the FPN head tensors (rmac_resnet_fpn.py:27-46)."""
    name = arch.split("_")[0]
    fpn = "_fpn" in arch
    if name in BASIC_BLOCKS:
        blocks, exp = BASIC_BLOCKS[name], 1
    else:
        blocks, exp = BLOCKS[arch_to_trunk(arch)], 4
    sd = {}
    sd["conv1.weight"] = _conv(seed, "conv1.weight", 64, 3, 7)
    sd.update(_bn(seed, "bn1", 64))
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip([64, 128, 256, 512], blocks), start=1):
        for b in range(nblk):
            p = "layer%d.%d." % (li, b)
            stride = 2 if (li > 1 and b == 0) else 1
            if exp == 4:
                sd[p + "conv1.weight"] = _conv(seed, p + "conv1.weight", planes, inplanes, 1)
                sd.update(_bn(seed, p + "bn1", planes))
                sd[p + "conv2.weight"] = _conv(seed, p + "conv2.weight", planes, planes, 3)
                sd.update(_bn(seed, p + "bn2", planes))
                sd[p + "conv3.weight"] = _conv(seed, p + "conv3.weight", planes * 4, planes, 1)
                sd.update(_bn(seed, p + "bn3", planes * 4, gamma_scale=res_gamma))
            else:
                sd[p + "conv1.weight"] = _conv(seed, p + "conv1.weight", planes, inplanes, 3)
                sd.update(_bn(seed, p + "bn1", planes))
                sd[p + "conv2.weight"] = _conv(seed, p + "conv2.weight", planes, planes, 3)
                sd.update(_bn(seed, p + "bn2", planes, gamma_scale=res_gamma))
            if b == 0 and (stride != 1 or inplanes != planes * exp):          # resnet.py:136-141
                sd[p + "downsample.0.weight"] = _conv(seed, p + "downsample.0.weight", planes * exp, inplanes, 1)
                sd.update(_bn(seed, p + "downsample.1", planes * exp, gamma_scale=0.7))
            inplanes = planes * exp
    feat = 512 * exp
    r = _rs(seed, "fc")
    if fpn:
        sd["conv1x5.weight"] = _conv(seed, "conv1x5.weight", 256 * exp, 512 * exp, 1)
        sd["conv3c4.weight"] = _conv(seed, "conv3c4.weight", 256 * exp, 256 * exp, 3)
        sd["adpoolx5.p"] = torch.ones(1) * float(gemp)
        sd["adpoolc4.p"] = torch.ones(1) * float(gemp)
        feat = 768 * exp
    else:
        sd["adpool.p"] = torch.ones(1) * float(gemp)
    sd["fc.weight"] = torch.from_numpy((r.standard_normal((out_dim, feat)) / np.sqrt(float(feat))).astype(np.float32))
    sd["fc.bias"] = torch.from_numpy((0.01 * r.standard_normal(out_dim)).astype(np.float32))
    return sd


def make_images_u8(batch: int, height: int, width: int, seed: int = 1234) -> np.ndarray:
    """uint8 HWC RGB images, (B,H,W,3)."""
    return np.random.RandomState(seed).randint(0, 256, (batch, height, width, 3), dtype=np.uint8)


def normalise_images(u8: np.ndarray) -> torch.Tensor:
    """ToTensor + Normalize(mean,std) of ``dirtorch/utils/transforms.py:27`` -> NCHW fp32."""
    x = torch.from_numpy(u8).permute(0, 3, 1, 2).to(torch.float32) / 255.0
    mean = torch.tensor(RGB_MEANS, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(RGB_STDS, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()


def make_images(batch: int, height: int, width: int, seed: int = 1234, smooth: bool = True) -> torch.Tensor:
    """Normalised NCHW fp32 batch. ``smooth`` adds low-frequency structure so that
    different images give clearly different descriptors (pure noise images all
    look alike to a random-weight CNN)."""
    u8 = make_images_u8(batch, height, width, seed).astype(np.float32)
    if smooth:
        r = np.random.RandomState(seed + 1)
        yy = np.linspace(0, 1, height, dtype=np.float32)[None, :, None, None]
        xx = np.linspace(0, 1, width, dtype=np.float32)[None, None, :, None]
        f = r.uniform(0.5, 6.0, (batch, 1, 1, 3, 2)).astype(np.float32)
        ph = r.uniform(0, 6.28, (batch, 1, 1, 3, 2)).astype(np.float32)
        wave = np.sin(6.28 * f[..., 0] * yy + ph[..., 0]) * np.cos(6.28 * f[..., 1] * xx + ph[..., 1])
        u8 = np.clip(0.35 * u8 + 0.65 * (127.5 + 127.5 * wave), 0, 255)
    return normalise_images(u8.astype(np.uint8))


def make_pca(dim: int = 2048, seed: int = 11, dtype=np.float32, whiten: bool = True):
    """Object with the fields ``common.transform`` reads (sklearn PCA attribute names).

    components_ is a seeded random orthonormal-ish mixing (rows scaled to unit
    norm), explained_variance_ a decaying positive spectrum, mean_ small."""
    r = np.random.RandomState(seed)
    comp = r.standard_normal((dim, dim)).astype(np.float64)
    comp /= np.linalg.norm(comp, axis=1, keepdims=True)
    var = (np.linspace(1.0, 0.02, dim) ** 2).astype(np.float64) / dim
    mean = (0.2 / np.sqrt(dim)) * r.standard_normal(dim)
    return SimpleNamespace(mean_=mean.astype(dtype), components_=comp.astype(dtype),
                           explained_variance_=var.astype(dtype), whiten=whiten)


def _unit_rows(x: np.ndarray) -> np.ndarray:
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def make_descriptor_db(n_db: int, n_query: int, dim: int = 2048, n_pos: int = 10,
                       db_seed: int = 2024, q_seed: int = 7, chunk: int = 65536):
    """Unit-norm fp32 database and queries with ``n_pos`` planted positives per query.

    Returns (db[N,D] fp32, queries[Q,D] fp32, positives[Q,n_pos] int64). Positive r of
    query i is ``normalize(q_i + sigma_r * noise)`` with sigma graded so that the
    cosines span ~0.8 .. 0.15 while unrelated rows have cos ~ N(0, 1/sqrt(D)).
    Rows for query i sit at distinct pseudo-random indices."""
    db = np.empty((n_db, dim), dtype=np.float32)
    r = np.random.RandomState(db_seed)
    for s in range(0, n_db, chunk):
        e = min(n_db, s + chunk)
        db[s:e] = _unit_rows(r.standard_normal((e - s, dim)).astype(np.float32))
    rq = np.random.RandomState(q_seed)
    q = _unit_rows(rq.standard_normal((n_query, dim)).astype(np.float32)).astype(np.float32)
    n_pos = min(n_pos, max(0, n_db // max(1, n_query)))
    pos = np.zeros((n_query, n_pos), dtype=np.int64)
    if n_pos:
        perm = np.random.RandomState(db_seed + 1).permutation(n_db)[: n_query * n_pos].reshape(n_query, n_pos)
        cos_targets = np.linspace(0.8, 0.15, n_pos)
        for i in range(n_query):
            for j in range(n_pos):
                c = cos_targets[j]
                noise = rq.standard_normal(dim).astype(np.float32)
                noise -= noise.dot(q[i]) * q[i]
                noise /= np.linalg.norm(noise)
                v = c * q[i] + np.sqrt(1.0 - c * c) * noise
                db[perm[i, j]] = (v / np.linalg.norm(v)).astype(np.float32)
                pos[i, j] = perm[i, j]
    return db, q, pos


def oxford_gt(pos: np.ndarray, n_junk: int = 0, n_db: int = 0, seed: int = 5) -> list:
    """Oxford-layout ``gnd`` list (``dirtorch/datasets/generic.py:134-145``): ok / junk index lists."""
    r = np.random.RandomState(seed)
    gnd = []
    for i in range(pos.shape[0]):
        ok = [int(v) for v in pos[i]]
        junk = []
        if n_junk and n_db:
            cand = r.permutation(n_db)[: n_junk + len(ok)]
            junk = [int(v) for v in cand if int(v) not in ok][:n_junk]
        gnd.append({"bbx": (0, 0, 1, 1), "ok": ok, "junk": junk})
    return gnd
