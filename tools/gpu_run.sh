#!/bin/bash
# One parameterised GPU lease script (replaces the per-session gpu_*.sh files of round 1):
#   gpurun --timeout 900 -- 'bash tools/gpu_run.sh <tag> <step> [<step> ...]'
# steps: tests | tests:<pytest -k expression> | i8 | i8:<N>:<d>:<M>:<kind> | bench | bench:<extra args> | suite | w4 | w5 | ncu_i8 | ncu_list | probe:<name>
# every step logs to gpurun_out/<tag>_<step>.log and prints its tail
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  name=${step%%:*}; arg=${step#*:}; [ "$arg" == "$step" ] && arg=""
  log=gpurun_out/${tag}_${name}.log
  case $name in
    tests)   if [ -n "$arg" ]; then timeout 600 python -m pytest tests -x -q -m gpu -k "$arg" > $log 2>&1; else timeout 900 python -m pytest tests -x -q -m gpu > $log 2>&1; fi; tail -15 $log ;;
    i8)      timeout 300 python tools/i8_bench.py ${arg//:/ } > $log 2>&1; tail -12 $log ;;
    bench)   timeout 600 python bench.py $arg > gpurun_out/${tag}_bench.json 2> $log; tail -3 $log; python -c "
import json,sys
j=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','kernel_ms_per_step','gpu_launches')}); print('e2e',j.get('e2e')); print('roofline',j.get('roofline')); print({k:v for k,v in j.items() if k.endswith('_arm') or k=='strong'})" ;;
    suite)   timeout 900 python bench.py --suite > gpurun_out/${tag}_suite.json 2> $log; cat gpurun_out/${tag}_suite.json | cut -c1-400 ;;
    w4)      timeout 900 python bench.py --workload 4 --steps 1 --warmup 1 --no-cpu-baseline $arg > gpurun_out/${tag}_w4.json 2> $log; tail -2 $log; cut -c1-1500 gpurun_out/${tag}_w4.json ;;
    w5)      timeout 600 python bench.py --workload 5 --steps 3 --warmup 1 $arg > gpurun_out/${tag}_w5.json 2> $log; tail -2 $log; cut -c1-1500 gpurun_out/${tag}_w5.json ;;
    ncu_list) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --candidates 151552 --no-cpu-baseline > $log 2>&1; tail -3 $log ;;
    ncu_i8)  timeout 600 ncu --set full --clock-control none --import-source on -k regex:trigemm_i8 -s 3 -c 1 -o gpurun_out/${tag}_k2i python tools/i8_bench.py 2048 6 151552 > $log 2>&1; tail -3 $log ;;
    probe)   timeout 120 ./tools/research/$arg > $log 2>&1; tail -20 $log ;;
    *) echo "unknown step $step" ;;
  esac
done
