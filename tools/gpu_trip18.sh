#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/launches_r1_b64.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launch rc=$?"
# 4th forward starts at launch 318 (106 per forward): layer3 block 5..: 318 + 3 + 9 + 12 + 3*5 = 357
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_pers|conv_halo" -s 350 -c 9 -o gpurun_out/prof_r1_conv_l3 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_full_conv.log 2>&1
echo "== ncu full conv rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:"conv_pers|conv_halo|stem_pers|s2d|maxpool" -s 318 -c 9 -o gpurun_out/prof_r1_conv_l1 \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-search > gpurun_out/ncu_full_conv1.log 2>&1
echo "== ncu full l1 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"conv_pers" -s 6 -c 2 -o gpurun_out/prof_r1_sim \
   python tools/search_profile.py > gpurun_out/ncu_full_sim.log 2>&1
echo "== ncu full sim rc=$?"; du -sh gpurun_out
