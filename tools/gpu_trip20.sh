#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_search.log 2>&1
echo "== pytest rc=$?"; tail -5 gpurun_out/pytest_search.log
timeout 600 python tools/search_profile.py 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 --batch 16 > gpurun_out/bench_2gpu.log 2>&1
echo "== bench 2gpu rc=$?"; grep '^{' gpurun_out/bench_2gpu.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("search 2gpu", d["search"]["value"], d["search"]["ms_per_step"], d["search"]["stats"])'
