#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -3 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1
echo "== bench rc=$?"; tail -1 gpurun_out/bench_full.log > gpurun_out/BENCH_line.json; python -c 'import sys,json; d=json.load(open("gpurun_out/BENCH_line.json")); print("extract", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "u8", d["e2e"]["uint8_input"]["value"], "conv TF", d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"]); print("search", d["search"]["value"], d["search"]["ms_per_step"], "cpu", d["cpu_baseline"]["value"], d["gpu_launches"])'
