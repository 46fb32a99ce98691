#!/bin/bash
# First GPU trip: structured conv diagnostics, parity tests (process-isolated by kernel family), smoke, short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1
echo "== diag done"; cat gpurun_out/diag_conv.txt | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -k "mma or not tcgen05" -p no:cacheprovider > gpurun_out/pytest_base.log 2>&1
echo "== pytest base rc=$?"; tail -25 gpurun_out/pytest_base.log
timeout 600 python -m pytest tests -m gpu -q -k "tcgen05" -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1
echo "== pytest tcgen05 rc=$?"; tail -25 gpurun_out/pytest_tc.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "== smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --batch 16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench16.log 2>&1
echo "== bench rc=$?"; tail -2 gpurun_out/bench16.log | cut -c1-1500
