#!/bin/bash
# One gpurun trip = a list of named tasks (arguments), run in order on the GPU box; everything lands in gpurun_out/.
#   tests [pytest -k expr]   full -m gpu suite (or a selection)
#   sanitize <tool>          compute-sanitizer over the small operator tests (tools/gpu_sanitize.sh)
#   searchprof               phase profile + ncu launch list of the search at 1M / 125k / 100k rows
#   bench [args]             python bench.py <args>
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc_all=0
while [ $# -gt 0 ]; do
  task=$1; shift
  case $task in
    tests)
      sel=${1:-}; [ -n "$sel" ] && shift
      if [ -n "$sel" ] && [ "$sel" != "-" ]; then
        timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$sel" > gpurun_out/pytest.log 2>&1
      else
        timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest.log 2>&1
      fi
      rc=$?; echo "== tests rc=$rc"; tail -15 gpurun_out/pytest.log; [ $rc -ne 0 ] && rc_all=$rc ;;
    sanitize)
      tool=$1; shift
      bash tools/gpu_sanitize.sh $tool; rc=$?; [ $rc -ne 0 ] && rc_all=$rc ;;
    searchprof)
      timeout 300 python tools/search_profile.py > gpurun_out/search_profile.log 2>&1; echo "== searchprof rc=$?"; cat gpurun_out/search_profile.log
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/search_launches.csv \
        python tools/search_profile.py one > gpurun_out/search_ncu.log 2>&1; echo "== search ncu rc=$?" ;;
    bench)
      args=${1:-}; shift
      timeout 900 python bench.py $args > gpurun_out/bench.log 2> gpurun_out/bench.err; rc=$?; echo "== bench rc=$rc"; tail -c 6000 gpurun_out/bench.log; tail -5 gpurun_out/bench.err; [ $rc -ne 0 ] && rc_all=$rc ;;
    traffic)
      # measured DRAM bytes per launch (metrics-only ncu passes; numbers printed under ncu are never bench values)
      M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
      timeout 600 ncu --metrics $M --clock-control none -k regex:"conv_|stem_|maxpool|head_|s2d" -s 330 -c 240 --csv --log-file gpurun_out/traffic_conv.csv \
        python bench.py --steps 1 --warmup 3 --no-search --no-cpu-baseline --no-latency > gpurun_out/traffic_conv.log 2>&1; echo "== traffic conv rc=$?"
      timeout 300 ncu --metrics $M --clock-control none -k regex:"conv_pers" -s 6 -c 4 --csv --log-file gpurun_out/traffic_search_1000q_1000k.csv \
        python tools/search_profile.py c4 > gpurun_out/traffic_s4.log 2>&1; echo "== traffic c4 rc=$?"
      timeout 300 ncu --metrics $M --clock-control none -k regex:"conv_pers" -s 6 -c 4 --csv --log-file gpurun_out/traffic_search_70q_100k.csv \
        python tools/search_profile.py c3 > gpurun_out/traffic_s3.log 2>&1; echo "== traffic c3 rc=$?"
      timeout 300 ncu --metrics $M --clock-control none -k regex:"conv_pers" -s 6 -c 4 --csv --log-file gpurun_out/traffic_search_1000q_125k.csv \
        python tools/search_profile.py one > gpurun_out/traffic_s8.log 2>&1; echo "== traffic shard rc=$?" ;;
    launchlist)
      # ncu launch list of one default bench run (per-launch times are cold-cache and serialised: shares only)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/r2_b64_launches.csv \
        python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-search --no-latency > gpurun_out/launchlist.log 2>&1; echo "== launchlist rc=$?" ;;
    benchall)
      # one JSON line per BASELINE configuration (1 GPU), kept for profiles/
      for cfg in c2 c3 c4 c5; do
        st=20; [ $cfg = c5 ] && st=5
        timeout 900 python bench.py --config $cfg --steps $st --warmup 3 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; rc=$?
        echo "== bench $cfg rc=$rc"; python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench_$cfg.json").read().strip().splitlines()[-1])
    r = l.get("roofline", {})
    print("   ", l["metric"], round(l["value"], 1), l["unit"], "ms/step", round(l["ms_per_step"], 3), "e2e", round(l["e2e"]["value"], 1),
          "| roofline", r.get("bound"), round(r.get("achieved", 0), 1), r.get("unit"), "frac", round(r.get("frac", 0), 3), "| clocks", l.get("clocks", {}).get("sm_mhz"), l.get("clocks", {}).get("reasons"))
    if "search" in l: print("    search", round(l["search"]["value"], 1), "q/s", l["search"]["ms_per_step"], l["search"].get("phases_ms"))
except Exception as e:
    print("    (no line)", e); print(open("gpurun_out/bench_$cfg.err").read()[-1500:])
PY
        [ $rc -ne 0 ] && rc_all=$rc
      done ;;
    dist)
      # multi-GPU: NCCL parity (tools/dist_check.py) + bench lines under torchrun; $1 = number of GPUs, $2 = bench args
      n=$1; shift; bargs=${1:-}; shift
      TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"
      timeout 600 $TR tools/dist_check.py > gpurun_out/dist_check_$n.log 2>&1; rc=$?; echo "== dist_check x$n rc=$rc"; tail -3 gpurun_out/dist_check_$n.log; [ $rc -ne 0 ] && rc_all=$rc
      timeout 300 $TR tools/dist_profile.py > gpurun_out/dist_profile_$n.log 2>&1; echo "== dist_profile x$n rc=$?"; tail -2 gpurun_out/dist_profile_$n.log
      timeout 900 $TR bench.py --gpus $n $bargs > gpurun_out/bench_${n}gpu.log 2> gpurun_out/bench_${n}gpu.err; rc=$?; echo "== bench x$n rc=$rc"; tail -c 3000 gpurun_out/bench_${n}gpu.log; tail -3 gpurun_out/bench_${n}gpu.err; [ $rc -ne 0 ] && rc_all=$rc ;;
    py)
      script=$1; shift
      name=$(basename ${script%% *} .py); timeout 900 python $script > gpurun_out/$name.log 2>&1; rc=$?; echo "== py $script rc=$rc"; tail -40 gpurun_out/$name.log; [ $rc -ne 0 ] && rc_all=$rc ;;
    *) echo "unknown task $task"; rc_all=2 ;;
  esac
done
exit $rc_all
