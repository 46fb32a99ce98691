#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_extract.py -m gpu -q -p no:cacheprovider -k "odd_and_multiscale" > gpurun_out/pytest_sizes.log 2>&1
echo "== pytest rc=$?"; tail -4 gpurun_out/pytest_sizes.log
timeout 600 python tools/variant_sweep.py 2>&1 | tail -6
