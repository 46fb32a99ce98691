#!/usr/bin/env python
"""Per-layer table for ResNet-101 @ 1024^2 from an ncu launch list: time vs the tensor / HBM lower bounds."""
import csv, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
fwd = int(sys.argv[3]) if len(sys.argv) > 3 else 3
peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
PT, PB = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if "dirb" in r["Kernel Name"]]
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
seq = [("s2d", 0, 0), ("stem", 2 * B * 512 * 512 * 64 * 147, B * (1024 * 1024 * 12 + 515 * 515 * 32 * 2 + 512 * 512 * 128)), ("maxpool", 0, B * 2 * 64 * (512 * 512 + 256 * 256))]
inpl, h = 64, 256
for li, (pl, nb) in enumerate(zip([64, 128, 256, 512], [3, 4, 23, 3])):
    for b in range(nb):
        s = 2 if (li > 0 and b == 0) else 1
        ho = h // s
        seq.append(("L%d c1 1x1 %d->%d" % (li + 1, inpl, pl), 2 * B * h * h * inpl * pl, B * h * h * (inpl + pl) * 2))
        seq.append(("L%d c2 3x3 %d s%d" % (li + 1, pl, s), 2 * B * ho * ho * pl * pl * 9, B * (h * h + ho * ho) * pl * 2))
        if b == 0:   # conv3 fused with the projection shortcut: K = [conv2 output | block input]
            seq.append(("L%d c3+ds 1x1 [%d|%d]->%d" % (li + 1, pl, inpl, pl * 4), 2 * B * ho * ho * (pl + inpl) * pl * 4,
                        B * (ho * ho * (pl + pl * 4) + h * h * inpl) * 2))
        else:
            seq.append(("L%d c3 1x1 %d->%d +res" % (li + 1, pl, pl * 4), 2 * B * ho * ho * pl * pl * 4, B * ho * ho * (pl + pl * 8) * 2))
        inpl, h = pl * 4, ho
seq += [("head", 0, 0)] * int(os.environ.get("HEAD_LAUNCHES", "1"))   # (round 1: 4 head kernels)
per = len(seq)
fw = rows[fwd * per:(fwd + 1) * per]
assert len(fw) == per, (len(rows), per)
groups = {}
for (name, fl, by), r in zip(seq, fw):
    g = groups.setdefault(name, [0, 0.0, 0.0, 0.0])
    g[0] += 1; g[1] += us(r); g[2] += fl; g[3] += by
tot = sum(g[1] for g in groups.values())
print("| layer type | launches | us | share | TFLOP/s | GB/s (algorithmic) | lower bound us (max of tensor, HBM) | efficiency |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
lb_tot = 0
for k, (n, t, fl, by) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    lb = max(fl / PT, by / PB) * 1e6
    lb_tot += lb if lb else t
    print("| %s | %d | %.0f | %.1f%% | %.0f | %.0f | %.0f (%s) | %s |" % (k, n, t, 100 * t / tot, fl / t / 1e6 if fl else 0, by / t / 1e3 if by else 0, lb,
          "tensor" if fl / PT > by / PB else "hbm", ("%.0f%%" % (100 * lb / t)) if lb else "-"))
print("\ntotal %.0f us for %d images (%.1f us/img, ncu serialised cold-cache timing); sum of lower bounds %.0f us (%.0f%%)" % (tot, B, tot / B, lb_tot, 100 * lb_tot / tot))
