#!/bin/bash
# Trip 2: re-run parity tests, the full-size bench line, a chunk sweep, and ncu (launch list + full capture).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "== pytest rc=$?"; tail -8 gpurun_out/pytest_all.log
timeout 900 python bench.py > gpurun_out/bench64.log 2>&1
echo "== bench64 rc=$?"; tail -1 gpurun_out/bench64.log | cut -c1-2000
for c in 1 2 4 8 16; do
  timeout 300 python bench.py --batch 16 --steps 3 --chunk $c --no-cpu-baseline > gpurun_out/bench16_c$c.log 2>&1
  echo "chunk $c: $(tail -1 gpurun_out/bench16_c$c.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"]["value"])' 2>&1 | tail -1)"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --batch 5 --chunk 5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches_r1.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 340 -c 12 -o gpurun_out/prof_conv_l3 \
   python bench.py --batch 5 --chunk 5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full1.log 2>&1
echo "== ncu full l3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc|conv_mma|maxpool|nchw" -s 330 -c 9 -o gpurun_out/prof_conv_l1 \
   python bench.py --batch 5 --chunk 5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
echo "== ncu full l1 rc=$?"; ls -la gpurun_out
