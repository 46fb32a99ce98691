#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -12 gpurun_out/pytest_all.log
for c in 8 16 32 64; do
  timeout 300 python bench.py --batch 64 --steps 3 --chunk $c --no-cpu-baseline > gpurun_out/bench64_c$c.log 2>&1
  echo "chunk $c: $(tail -1 gpurun_out/bench64_c$c.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"])' 2>&1 | tail -1)"
done
timeout 600 python tools/cpu_threads.py > gpurun_out/cpu_threads.log 2>&1; cat gpurun_out/cpu_threads.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1b.csv \
   python bench.py --batch 16 --chunk 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches_r1b.csv
