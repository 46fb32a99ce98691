#!/usr/bin/env python
"""profiles/r2_traffic.json from ncu CSV logs (metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum):
measured DRAM traffic per launch of the kernels bench.py reports rooflines for.

    python tools/traffic_from_ncu.py gpurun_out/traffic_conv.csv [gpurun_out/traffic_search.csv ...]
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
        "nsecond": 1e-9, "second": 1.0}


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H = rows[h]
    ki, mi, ui, vi = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Unit"), H.index("Metric Value")
    out = {}
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        d = out.setdefault(int(r[0]), {"kernel": r[ki]})
        d[r[mi]] = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
    return [out[k] for k in sorted(out)]


def short(name):
    m = re.match(r"(?:void )?(?:dirb::)?(?:<unnamed>::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def main():
    res = {}
    for path in sys.argv[1:]:
        ls = launches(path)
        by = defaultdict(list)
        for l in ls:
            by[short(l["kernel"])].append(l)
        tag = os.path.basename(path)
        conv = [l for l in ls if re.search(r"conv_pers_kernel<\d+, \d+, 0|conv_halo_kernel|conv_c23", short(l["kernel"]))]
        if conv and "conv" in tag:
            n = len(conv)
            res["conv_stack"] = {
                "dram_bytes_per_launch": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in conv) / n,
                "launches": n, "avg_launch_us": sum(l.get("gpu__time_duration.sum", 0) for l in conv) / n * 1e6,
                "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over %d convolution launches (%s)" % (n, tag)}
            dom_name = max(by, key=lambda k: sum(l.get("gpu__time_duration.sum", 0) for l in by[k]) if re.search("conv_", k) else -1)
            dom = by[dom_name]
            res["dominant"] = {"kernel": dom_name, "launches": len(dom),
                               "dram_bytes_per_launch": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in dom) / len(dom),
                               "avg_launch_us": sum(l.get("gpu__time_duration.sum", 0) for l in dom) / len(dom) * 1e6}
        res.setdefault("kernels", {})[tag] = {
            k: {"launches": len(v),
                "dram_bytes_per_launch": sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in v) / len(v),
                "avg_launch_us": sum(l.get("gpu__time_duration.sum", 0) for l in v) / len(v) * 1e6} for k, v in by.items()}
        # search filter passes: the largest EPI=2 (PERS_EPI_SIM_FILTER) launch of a search log
        filt = [l for l in ls if re.search(r"conv_pers_kernel<256, 4, 2", short(l["kernel"]))]
        m = re.search(r"search_(\d+)q_(\d+)k", tag)
        if filt and m:
            big = max(filt, key=lambda l: l.get("gpu__time_duration.sum", 0))
            res["filter_gemm_%sq_%sk" % (m.group(1), m.group(2))] = {
                "dram_bytes_per_launch": big.get("dram__bytes_read.sum", 0) + big.get("dram__bytes_write.sum", 0),
                "launch_us": big.get("gpu__time_duration.sum", 0) * 1e6, "source": "ncu, " + tag}
    out = os.path.join(REPO, "profiles", "r2_traffic.json")
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out, {k: v for k, v in res.items() if k != "kernels"})


if __name__ == "__main__":
    main()
