#!/bin/bash
# Full GPU test suite + smoke, no bench (see tools/gpu_final.sh for the complete round-end sequence).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -15 gpurun_out/pytest_all.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -1 gpurun_out/smoke.log
