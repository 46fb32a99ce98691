#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "stem" > gpurun_out/pytest_stem.log 2>&1
echo "== pytest stem rc=$?"; tail -12 gpurun_out/pytest_stem.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -12 gpurun_out/pytest_all.log
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1
echo "== bench rc=$?"; tail -1 gpurun_out/bench_full.log | cut -c1-5000
