#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -k "two_gpu" > gpurun_out/pytest_2gpu.log 2>&1
echo "== pytest 2gpu rc=$?"; tail -5 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1
echo "== bench 2gpu rc=$?"; tail -1 gpurun_out/bench_2gpu.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("extract", d["value"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]); print("search", d["search"]["value"], d["search"]["ms_per_step"], d["search"]["e2e"]["value"])'
tail -5 gpurun_out/bench_2gpu.log | cut -c1-600
