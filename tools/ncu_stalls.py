#!/usr/bin/env python
"""profiles/<name>.md from one `ncu --set full --import-source on` capture of a single kernel: headline metrics (raw page)
and the warp-state sampling of the SASS (source page): totals per stall reason and the instructions that hold the most
samples.  Run here (no GPU needed):  python tools/ncu_stalls.py gpurun_out/x.ncu-rep profiles/r2_x.md "title" """
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic"]


def page(name):
    return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(page("raw"))))
H, U, V = raw[0], raw[1], raw[2]
vals = {h: (v, u) for h, u, v in zip(H, U, V)}
src = list(csv.reader(io.StringIO(page("source"))))
kernel = src[0][1]
SH = src[1]
idx = {h: i for i, h in enumerate(SH)}
rows = src[2:]
stalls = [h for h in SH if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]]) for r in rows)
agg = sorted(((s, sum(int(r[idx[s]] or 0) for r in rows)) for s in stalls), key=lambda kv: -kv[1])
with open(out, "w") as f:
    f.write("# %s\n\n`%s`\n\nSource: `ncu --set full --import-source on --clock-control none` (one launch, full clock, cold caches); "
            "summary by `tools/ncu_stalls.py`.\n\n| metric | value |\n|---|---|\n" % (title, kernel))
    for k in KEYS:
        if k in vals:
            f.write("| `%s` | %s %s |\n" % (k, vals[k][0], vals[k][1]))
    f.write("\nWarp-state samples: %d in total (12 warps per CTA: TMA producer, MMA issuer, residual producer, TMEM allocator, "
            "8 epilogue warps).\n\n| stall reason | samples | share |\n|---|---:|---:|\n" % tot)
    for s, n in agg[:8]:
        f.write("| %s | %d | %.1f %% |\n" % (s, n, 100.0 * n / tot))
    f.write("\nInstructions holding the most samples (SASS index, samples, executions, instruction, top reasons):\n\n```\n")
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][idx["# Samples"]]))[:28]
    for i in sorted(top):
        r = rows[i]
        st = sorted(((s[6:], int(r[idx[s]] or 0)) for s in stalls), key=lambda kv: -kv[1])[:2]
        f.write("%5d %5s %8s  %-62s %s\n" % (i, r[idx["# Samples"]], r[idx["Instructions Executed"]], r[idx["Source"]].strip()[:62],
                                              " ".join("%s=%d" % kv for kv in st if kv[1])))
    f.write("```\n")
print("wrote", out)
