#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_search.log 2>&1
echo "== pytest rc=$?"; tail -3 gpurun_out/pytest_search.log
timeout 600 python tools/search_profile.py 2>&1 | tail -4
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_r1_bench.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches_r1_bench.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_pers" -s 440 -c 12 -o gpurun_out/prof_r1_conv_l3 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_full_conv.log 2>&1
echo "== ncu full conv rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:"conv_pers|stem_pers|s2d|maxpool" -s 420 -c 9 -o gpurun_out/prof_r1_conv_l1 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_full_conv1.log 2>&1
echo "== ncu full l1 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_pers" -s 6 -c 2 -o gpurun_out/prof_r1_sim \
   python tools/search_profile.py > gpurun_out/ncu_full_sim.log 2>&1
echo "== ncu full sim rc=$?"; ls -la gpurun_out/; du -sh gpurun_out
