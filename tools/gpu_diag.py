#!/usr/bin/env python
"""Structured diagnostics for the tcgen05 implicit-GEMM convolution (run on the GPU box).

Each check uses weights that make the expected output trivially readable (identity / single-tap delta), so a
wrong smem descriptor, swizzle, TMA coordinate or TMEM lane mapping shows up as a recognisable pattern rather
than "max error large".  Every check runs in its own process (a trapped kernel kills the CUDA context).

    python tools/gpu_diag.py            # run all checks, report to gpurun_out/diag_conv.txt
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

CHECKS = ["identity_1x1", "k256_1x1", "ragged_1x1", "tap_3x3", "tap_3x3_s2", "down_1x1_s2", "random_3x3", "multiimg_3x3",
          "sim_dense"]


def run_check(name):
    import numpy as np
    import torch
    import torch.nn.functional as F
    from dirb200 import ops
    dev = "cuda:0"
    r = np.random.RandomState(0)
    out = {"check": name}

    def conv(x, w, k, stride, pad, impl):
        cout = w.shape[0]
        wp = ops.pack_conv_weight(w).to(dev)
        one = torch.ones(cout, device=dev)
        zero = torch.zeros(cout, device=dev)
        y = ops.conv_bn_act(x.to(dev), wp, cout, k, k, stride, pad, one, zero, None, False, impl)
        torch.cuda.synchronize()
        return y.float().cpu()

    def ref(x, w, stride, pad):
        return F.conv2d(x.float().permute(0, 3, 1, 2), w.half().float(), None, stride=stride, padding=pad).permute(0, 2, 3, 1)

    def compare(y, yr):
        err = (y - yr).abs()
        out["max_err"] = float(err.max())
        out["ref_absmax"] = float(yr.abs().max())
        bad = err > 2e-3 * max(1.0, float(yr.abs().max()))
        out["bad_frac"] = float(bad.float().mean())
        if bad.any():
            b, h, w_, c = bad.shape
            rows = bad.reshape(-1, c).any(dim=1).nonzero().flatten()[:16].tolist()
            cols = bad.reshape(-1, c).any(dim=0).nonzero().flatten()[:16].tolist()
            out["bad_rows_first"] = rows
            out["bad_cols_first"] = cols
            out["bad_row_count"] = int(bad.reshape(-1, c).any(dim=1).sum())
            out["bad_col_count"] = int(bad.reshape(-1, c).any(dim=0).sum())

    if name == "identity_1x1":
        x = torch.from_numpy(r.standard_normal((1, 16, 16, 64)).astype(np.float32)).half()
        w = torch.eye(64).view(64, 64, 1, 1)
        y = conv(x, w, 1, 1, 0, 0)
        compare(y, x.float())
        # which input row does each output row equal?
        xf, yf = x.float().reshape(-1, 64), y.reshape(-1, 64)
        match = []
        for i in [0, 1, 7, 8, 31, 32, 64, 100, 127, 128, 255]:
            d = (xf - yf[i]).abs().sum(dim=1)
            match.append((i, int(d.argmin()), float(d.min())))
        out["row_match(out_row,in_row,dist)"] = match
    elif name == "k256_1x1":
        x = torch.from_numpy(r.standard_normal((2, 16, 16, 256)).astype(np.float32)).half()
        w = torch.from_numpy((r.standard_normal((128, 256, 1, 1)) / 16).astype(np.float32))
        compare(conv(x, w, 1, 1, 0, 0), ref(x, w, 1, 0))
        # partial-K hypotheses: does the output equal the sum over only some k-blocks / k-sixteenths?
        y = conv(x, w, 1, 1, 0, 0)
        for tag, sel in (("first16_of_each_64", [i for i in range(256) if i % 64 < 16]), ("kblock0", list(range(64)))):
            wm = torch.zeros_like(w)
            wm[:, sel] = w[:, sel]
            out["err_vs_" + tag] = float((y - ref(x, wm, 1, 0)).abs().max())
    elif name == "ragged_1x1":
        x = torch.from_numpy(r.standard_normal((2, 14, 14, 64)).astype(np.float32)).half()
        w = torch.from_numpy((r.standard_normal((64, 64, 1, 1)) / 8).astype(np.float32))
        compare(conv(x, w, 1, 1, 0, 0), ref(x, w, 1, 0))
    elif name in ("tap_3x3", "tap_3x3_s2"):
        stride = 2 if name.endswith("s2") else 1
        x = torch.from_numpy(r.standard_normal((1, 16, 16, 64)).astype(np.float32)).half()
        res = {}
        for kh in range(3):
            for kw in range(3):
                w = torch.zeros(64, 64, 3, 3)
                w[:, :, kh, kw] = torch.eye(64)
                y, yr = conv(x, w, 3, stride, 1, 0), ref(x, w, stride, 1)
                res["%d%d" % (kh, kw)] = float((y - yr).abs().max())
        out["per_tap_max_err"] = res
        out["max_err"] = max(res.values())
    elif name == "down_1x1_s2":
        x = torch.from_numpy(r.standard_normal((2, 14, 14, 64)).astype(np.float32)).half()
        w = torch.eye(64).view(64, 64, 1, 1)
        compare(conv(x, w, 1, 2, 0, 0), ref(x, w, 2, 0))
    elif name == "random_3x3":
        x = torch.from_numpy(r.standard_normal((1, 20, 24, 128)).astype(np.float32)).half()
        w = torch.from_numpy((r.standard_normal((128, 128, 3, 3)) / 34).astype(np.float32))
        compare(conv(x, w, 3, 1, 1, 0), ref(x, w, 1, 1))
        out["mma_max_err"] = float((conv(x, w, 3, 1, 1, 1) - ref(x, w, 1, 1)).abs().max())
    elif name == "multiimg_3x3":
        x = torch.from_numpy(r.standard_normal((5, 7, 7, 64)).astype(np.float32)).half()
        w = torch.from_numpy((r.standard_normal((64, 64, 3, 3)) / 24).astype(np.float32))
        compare(conv(x, w, 3, 1, 1, 0), ref(x, w, 1, 1))
    elif name == "sim_dense":
        q = torch.from_numpy(r.standard_normal((70, 128)).astype(np.float32))
        db = torch.from_numpy(r.standard_normal((1000, 128)).astype(np.float32))
        q /= q.norm(dim=1, keepdim=True)
        db /= db.norm(dim=1, keepdim=True)
        idx = ops.Index(db.to(dev))
        s, i = idx.search(q.to(dev), 10)
        torch.cuda.synchronize()
        full = (q.double() @ db.double().T)
        rs, ri = full.topk(10, dim=1)
        out["idx_equal"] = bool((i.cpu() == ri).all())
        out["max_err"] = float((s.cpu() - rs).abs().max())
        out["stats"] = idx.stats()
    print("DIAG " + json.dumps(out))


def main():
    if len(sys.argv) > 1:
        run_check(sys.argv[1])
        return
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    lines = []
    for c in CHECKS:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=180)
            got = [l for l in p.stdout.splitlines() if l.startswith("DIAG ")]
            lines.append(got[0] if got else "DIAG-FAIL %s rc=%d :: %s" % (c, p.returncode, (p.stderr or p.stdout)[-600:].replace("\n", " | ")))
        except subprocess.TimeoutExpired:
            lines.append("DIAG-TIMEOUT %s" % c)
    text = "\n".join(lines)
    print(text)
    with open(os.path.join(REPO, "gpurun_out", "diag_conv.txt"), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
