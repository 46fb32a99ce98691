#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -15 gpurun_out/pytest_all.log
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1
echo "== bench rc=$?"; tail -3 gpurun_out/bench_full.log | cut -c1-6000
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1
echo "== bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-1500
