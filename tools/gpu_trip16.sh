#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_extract.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_ext.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/pytest_ext.log
timeout 1200 python bench.py --no-search --no-cpu-baseline > gpurun_out/bench_ns.log 2>&1
echo "== bench rc=$?"; tail -1 gpurun_out/bench_ns.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("extract", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv TF", d["roofline"]["achieved"], d["roofline"]["classes_ms"], d["clocks"])'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/launches_r1_b64.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_launch.log 2>&1
echo "== ncu rc=$?"
