#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_extract.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_ext.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/pytest_ext.log
timeout 1200 python bench.py --no-search --no-cpu-baseline > gpurun_out/bench_ns.log 2>&1
echo "== bench rc=$?"; tail -1 gpurun_out/bench_ns.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("extract", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv TF", d["roofline"]["achieved"], d["roofline"]["classes_ms"], d["clocks"])'
for hc in 8 32; do timeout 600 python bench.py --no-search --no-cpu-baseline --host-chunk $hc --steps 3 > gpurun_out/bench_hc$hc.log 2>&1; echo "host_chunk $hc: $(tail -1 gpurun_out/bench_hc$hc.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"]["value"])')"; done
