#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -5 gpurun_out/pytest_all.log
timeout 1200 python bench.py --no-cpu-baseline > gpurun_out/bench_full.log 2>&1
echo "== bench rc=$?"; tail -1 gpurun_out/bench_full.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("extract", d["value"], "e2e", d["e2e"]["value"]); print("search", json.dumps(d["search"])[:1500])'
timeout 600 python bench.py --no-cpu-baseline --batch 8 --steps 3 --search-n 100000 --search-q 70 > gpurun_out/bench_c3.log 2>&1
echo "== bench c3 rc=$?"; tail -1 gpurun_out/bench_c3.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("search c3", json.dumps(d["search"])[:1500])'
