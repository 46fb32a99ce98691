#!/usr/bin/env python
"""Summarise ncu captures (run here, no GPU needed) into profiles/: launch-list shares and per-kernel metrics."""
import csv
import collections
import io
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "MB"),
    ("dram__bytes_write.sum", "MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("lts__t_sector_hit_rate.pct", "%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%"),
    ("sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active", "%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"),
    ("launch__registers_per_thread", ""),
    ("launch__grid_size", ""),
]


def us(v, unit):
    v = float(v.replace(",", ""))
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v)


def launch_list(path, name):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"])
        k = re.sub(r"^void ", "", k)[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us(r["Metric Value"], r["Metric Unit"])
    ours = {k: v for k, v in agg.items() if "dirb" in k}
    tot = sum(v[1] for v in ours.values())
    out = ["# %s: ncu launch list (gpu__time_duration.sum, --clock-control none), library kernels only" % name, "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% |" % (k, n, t, 100 * t / tot))
    out.append("")
    out.append("total library kernel time %.1f us over %d launches" % (tot, sum(v[0] for v in ours.values())))
    return "\n".join(out)


def full_report(rep, name, note=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, u) for k, u in KEYS if k in idx]
    out = ["# %s: ncu --set full, one row per captured launch" % name, "", note, "",
           "| # | kernel | grid | " + " | ".join(k.split(".")[0].replace("__", ".") + (" [%s]" % units[idx[k]] if units[idx[k]] else "") for k, _ in cols) + " |",
           "|---|---|---|" + "---:|" * len(cols)]
    for n, d in enumerate(data):
        kn = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void ", "")[:60]
        out.append("| %d | `%s` | %s | " % (n, kn, d[idx["Grid Size"]]) + " | ".join(d[idx[k]] for k, _ in cols) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    g = os.path.join(REPO, "gpurun_out")
    jobs = sys.argv[1:] or ["launch", "l3", "l1", "sim"]
    if "launch" in jobs:
        open(os.path.join(OUT, "r1_bench_launches.md"), "w").write(launch_list(os.path.join(g, "launches_r1_bench.csv"), "round 1, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search`") + "\n")
    for tag, note in (("l3", "ResNet-101 layer3 blocks at batch 64 x 1024^2 (c2 3x3, c3 1x1+residual, c1 1x1 ...)"),
                      ("l1", "stem (s2d + tcgen05 stem), maxpool and the first layer1 convolutions"),):
        if tag in jobs:
            open(os.path.join(OUT, "r1_conv_%s_full.md" % tag), "w").write(full_report(os.path.join(g, "prof_r1_conv_%s.ncu-rep" % tag), "round 1 conv " + tag, note) + "\n")
    if "sim" in jobs:
        open(os.path.join(OUT, "r1_sim_full.md"), "w").write(full_report(os.path.join(g, "prof_r1_sim.ncu-rep"), "round 1 similarity GEMM (1000 q x 1M x 2048)", "seed (group-max) and filter passes of dirb200_index_search") + "\n")
