#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
OUT_DIR=gpurun_out timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/dist8.log 2>&1
echo "== dist_check 8 rc=$?"; grep -E "DIST-" gpurun_out/dist8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_8gpu.log 2>&1
echo "== bench 8gpu rc=$?"; grep '^{' gpurun_out/bench_8gpu.log | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("extract", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"]); print("search", d["search"]["value"], d["search"]["ms_per_step"], d["search"]["e2e"]["value"], d["search"]["stats"])'
tail -3 gpurun_out/bench_8gpu.log | cut -c1-400
