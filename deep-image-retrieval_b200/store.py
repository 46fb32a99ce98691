"""On-disk descriptor store: the database side of the search path, persisted in row shards.

The reference persists descriptors as ONE ``.npy`` per call (``extract_features.py:61-67`` writes ``<output>``,
``test_dir.py:130-134`` writes ``feats.bdescs.npy`` / ``feats.qdescs.npy``) and reloads them with ``np.load``
(``test_dir.py:126-128``).  A 1M x 2048 database is 8 GB in that form and has to pass through one process.  The
store keeps the same rows in the same order, cut into contiguous row shards:

    <dir>/manifest.json           {"format": "dirb200-descriptor-store", "version": 1, "n_rows", "dim", "dtype",
                                   "shards": [{"file", "row_start", "n_rows"}, ...], "meta": {...}}
    <dir>/shard-00000.npy ...     plain NumPy arrays (n_rows_i, dim), float32 or float16, C order

Every shard is an ordinary ``.npy``, so ``np.concatenate([np.load(f) for f in shards])`` is exactly the array the
reference would have written, and :func:`DescriptorStore.to_npy` / :func:`DescriptorStore.from_npy` convert both ways.
Shards are written independently (one per rank after a multi-GPU extraction: rank g writes the rows it extracted,
:func:`write_rank_shard`, then rank 0 writes the manifest), and a rank of ANY world size reads just its
``dist.shard_rows`` range through memory maps - shard boundaries on disk need not match the reading world size.

Host-side I/O only: no kernels here.  Rows go disk -> (mmap) -> pinned staging -> HBM in bounded chunks.
"""
from __future__ import annotations

import json
import os

import numpy as np

FORMAT = "dirb200-descriptor-store"
VERSION = 1
MANIFEST = "manifest.json"
_DTYPES = {"float32": np.float32, "float16": np.float16}


def _shard_name(i: int) -> str:
    return "shard-%05d.npy" % i


def _dtype_name(dtype) -> str:
    name = np.dtype(dtype).name
    if name not in _DTYPES:
        raise TypeError("descriptor store dtype must be float32 or float16, not %s" % name)
    return name


def _write_manifest(path, n_rows, dim, dtype, shards, meta):
    doc = {"format": FORMAT, "version": VERSION, "n_rows": int(n_rows), "dim": int(dim), "dtype": _dtype_name(dtype),
           "shards": shards, "meta": meta or {}}
    tmp = os.path.join(path, MANIFEST + ".tmp")
    with open(tmp, "w") as f:
        json.dump(doc, f, indent=1)
    os.replace(tmp, os.path.join(path, MANIFEST))     # the manifest appears atomically, after all shards exist


class DescriptorStoreWriter:
    """Sequential writer: ``append`` row blocks in database order, ``close`` writes the manifest.

    Rows are buffered up to ``rows_per_shard`` and flushed as one ``.npy`` each; nothing but the current shard is
    held in memory."""

    def __init__(self, path: str, dim: int, dtype=np.float32, rows_per_shard: int = 131072, meta: dict = None):
        if rows_per_shard < 1:
            raise ValueError("rows_per_shard must be >= 1")
        self.path, self.dim, self.dtype = path, int(dim), np.dtype(_DTYPES[_dtype_name(dtype)])
        self.rows_per_shard, self.meta = int(rows_per_shard), dict(meta or {})
        os.makedirs(path, exist_ok=True)
        self._buf, self._buffered, self._shards, self._rows, self._closed = [], 0, [], 0, False

    def _flush(self, n):
        block = np.concatenate(self._buf, axis=0) if len(self._buf) != 1 else self._buf[0]
        out, rest = block[:n], block[n:]
        name = _shard_name(len(self._shards))
        np.save(os.path.join(self.path, name), np.ascontiguousarray(out))
        self._shards.append({"file": name, "row_start": self._rows, "n_rows": int(out.shape[0])})
        self._rows += int(out.shape[0])
        self._buf = [rest] if rest.shape[0] else []
        self._buffered = int(rest.shape[0])

    def append(self, rows):
        if self._closed:
            raise ValueError("store already closed")
        rows = np.asarray(rows.detach().cpu().numpy() if hasattr(rows, "detach") else rows)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError("expected (n, %d) rows, got %s" % (self.dim, rows.shape))
        self._buf.append(rows.astype(self.dtype, copy=False))
        self._buffered += rows.shape[0]
        while self._buffered >= self.rows_per_shard:
            self._flush(self.rows_per_shard)

    def close(self):
        if self._closed:
            return DescriptorStore(self.path)
        if self._buffered:
            self._flush(self._buffered)
        _write_manifest(self.path, self._rows, self.dim, self.dtype, self._shards, self.meta)
        self._closed = True
        return DescriptorStore(self.path)

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        if exc_type is None:
            self.close()


def write_store(path: str, descs, rows_per_shard: int = 131072, dtype=None, meta: dict = None):
    """Persist an (N, D) descriptor matrix; returns the opened :class:`DescriptorStore`."""
    descs = np.asarray(descs.detach().cpu().numpy() if hasattr(descs, "detach") else descs)
    with DescriptorStoreWriter(path, descs.shape[1], dtype or descs.dtype, rows_per_shard, meta) as w:
        for s in range(0, descs.shape[0], rows_per_shard):
            w.append(descs[s:s + rows_per_shard])
    return DescriptorStore(path)


def write_rank_shard(path: str, rows, rank: int, dtype=None) -> dict:
    """One rank's rows (the images it extracted, in database order) -> ``shard-<rank>.npy``.  Returns the shard's
    manifest entry without ``row_start``; :func:`finalize_rank_shards` assigns the offsets."""
    rows = np.asarray(rows.detach().cpu().numpy() if hasattr(rows, "detach") else rows)
    os.makedirs(path, exist_ok=True)
    name = _shard_name(rank)
    np.save(os.path.join(path, name), np.ascontiguousarray(rows.astype(_DTYPES[_dtype_name(dtype or rows.dtype)], copy=False)))
    return {"file": name, "n_rows": int(rows.shape[0])}


def finalize_rank_shards(path: str, world: int, meta: dict = None):
    """Called by rank 0 once every rank has written its shard (after a barrier): reads the shard headers, assigns
    contiguous row offsets in rank order and writes the manifest."""
    shards, start, dim, dtype = [], 0, None, None
    for r in range(world):
        a = np.load(os.path.join(path, _shard_name(r)), mmap_mode="r")
        if dim is None:
            dim, dtype = a.shape[1], a.dtype
        if a.ndim != 2 or a.shape[1] != dim or a.dtype != dtype:
            raise ValueError("shard %d has shape %s / dtype %s, expected (*, %d) %s" % (r, a.shape, a.dtype, dim, dtype))
        shards.append({"file": _shard_name(r), "row_start": start, "n_rows": int(a.shape[0])})
        start += int(a.shape[0])
    _write_manifest(path, start, dim, dtype, shards, meta)
    return DescriptorStore(path)


def write_distributed(path: str, rows_local, group=None, dtype=None, meta: dict = None):
    """Collective write under torch.distributed (one process per GPU): every rank writes the rows it owns - in rank
    order they must form the database order, as after ``dist.shard_rows`` - as its own shard file, then rank 0
    assigns the offsets and writes the manifest.  Two barriers, no data-path collective: descriptors never leave
    the rank that extracted them.  Without an initialised process group this is ``write_store`` with one shard."""
    import torch.distributed as tdist
    if not (tdist.is_available() and tdist.is_initialized()):
        write_rank_shard(path, rows_local, 0, dtype)
        return finalize_rank_shards(path, 1, meta)
    rank, world = tdist.get_rank(group), tdist.get_world_size(group)
    write_rank_shard(path, rows_local, rank, dtype)
    tdist.barrier(group)                       # every shard file is complete
    if rank == 0:
        finalize_rank_shards(path, world, meta)
    tdist.barrier(group)                       # the manifest exists
    return DescriptorStore(path)


class DescriptorStore:
    """Read side.  Shards are memory-mapped lazily; ``read_rows`` may span shard boundaries."""

    def __init__(self, path: str):
        self.path = path
        try:
            with open(os.path.join(path, MANIFEST)) as f:
                doc = json.load(f)
        except FileNotFoundError:
            raise FileNotFoundError("%s is not a descriptor store (no %s)" % (path, MANIFEST)) from None
        if doc.get("format") != FORMAT or doc.get("version") != VERSION:
            raise ValueError("%s: unsupported store format %r version %r" % (path, doc.get("format"), doc.get("version")))
        self.n_rows, self.dim = int(doc["n_rows"]), int(doc["dim"])
        self.dtype = np.dtype(_DTYPES[doc["dtype"]])
        self.shards, self.meta = doc["shards"], doc.get("meta", {})
        pos = 0
        for s in self.shards:
            if s["row_start"] != pos:
                raise ValueError("%s: shards are not contiguous at row %d" % (path, pos))
            pos += s["n_rows"]
        if pos != self.n_rows:
            raise ValueError("%s: manifest n_rows %d != sum of shards %d" % (path, self.n_rows, pos))
        self._maps = {}

    @staticmethod
    def is_store(path: str) -> bool:
        return bool(path) and os.path.isfile(os.path.join(path, MANIFEST))

    def __len__(self):
        return self.n_rows

    def _map(self, i):
        m = self._maps.get(i)
        if m is None:
            s = self.shards[i]
            m = np.load(os.path.join(self.path, s["file"]), mmap_mode="r")
            if m.shape != (s["n_rows"], self.dim) or m.dtype != self.dtype:
                raise ValueError("%s: %s is %s %s, manifest says (%d, %d) %s" % (self.path, s["file"], m.shape, m.dtype,
                                                                              s["n_rows"], self.dim, self.dtype))
            self._maps[i] = m
        return m

    def read_rows(self, start: int, end: int, out: np.ndarray = None) -> np.ndarray:
        """Rows [start, end) as one C-contiguous array of the store dtype (into ``out`` if given)."""
        if not (0 <= start <= end <= self.n_rows):
            raise IndexError("rows [%d, %d) outside [0, %d)" % (start, end, self.n_rows))
        if out is None:
            out = np.empty((end - start, self.dim), self.dtype)
        if out.shape != (end - start, self.dim):
            raise ValueError("out has shape %s, expected %s" % (out.shape, (end - start, self.dim)))
        for i, s in enumerate(self.shards):
            lo, hi = max(start, s["row_start"]), min(end, s["row_start"] + s["n_rows"])
            if lo < hi:
                out[lo - start:hi - start] = self._map(i)[lo - s["row_start"]:hi - s["row_start"]]
        return out

    def read_all(self) -> np.ndarray:
        return self.read_rows(0, self.n_rows)

    def rank_range(self, rank: int, world: int):
        """The rows rank `rank` of `world` owns: the same contiguous split as ``dist.shard_rows``."""
        return (self.n_rows * rank) // world, (self.n_rows * (rank + 1)) // world

    def load_rows_to_device(self, start: int, end: int, device, chunk_rows: int = 65536):
        """Rows [start, end) -> (fp32 (n, D), fp16 (n, D)) device tensors, staged through pinned host memory
        ``chunk_rows`` at a time (copy k+1 is read from disk while copy k is in flight)."""
        import torch
        n = end - start
        dev = torch.device(device)
        d32 = torch.empty((n, self.dim), dtype=torch.float32, device=dev)
        d16 = torch.empty((n, self.dim), dtype=torch.float16, device=dev)
        tdt = torch.float32 if self.dtype == np.float32 else torch.float16
        pin = dev.type == "cuda"
        stage = [torch.empty((min(chunk_rows, max(n, 1)), self.dim), dtype=tdt, pin_memory=pin) for _ in range(2)]
        done = [None, None]
        for j, s in enumerate(range(0, n, chunk_rows)):
            e = min(n, s + chunk_rows)
            buf = stage[j & 1]
            if done[j & 1] is not None:
                done[j & 1].synchronize()                      # the copy that last used this staging buffer
            self.read_rows(start + s, start + e, out=buf.numpy()[:e - s])
            src = buf[:e - s].to(dev, non_blocking=True)
            d32[s:e].copy_(src)                                # exact: fp16 -> fp32 widening or identity
            d16[s:e].copy_(src)                                # fp32 -> fp16 round-to-nearest-even (as f32_to_f16)
            if pin:
                done[j & 1] = torch.cuda.Event()
                done[j & 1].record()
        return d32, d16

    def to_npy(self, file: str):
        """Export as the single ``.npy`` the reference reads (``test_dir.py:126-128``): float32, all rows."""
        out = np.lib.format.open_memmap(file, mode="w+", dtype=np.float32, shape=(self.n_rows, self.dim))
        for s in self.shards:
            out[s["row_start"]:s["row_start"] + s["n_rows"]] = self.read_rows(s["row_start"], s["row_start"] + s["n_rows"])
        out.flush()
        del out

    @staticmethod
    def from_npy(file: str, path: str, rows_per_shard: int = 131072, dtype=None, meta: dict = None):
        """Import a reference-format ``.npy`` (memory-mapped, never fully resident)."""
        src = np.load(file, mmap_mode="r")
        with DescriptorStoreWriter(path, src.shape[1], dtype or src.dtype, rows_per_shard, meta) as w:
            for s in range(0, src.shape[0], rows_per_shard):
                w.append(np.asarray(src[s:s + rows_per_shard]))
        return DescriptorStore(path)
