#!/bin/bash
# Multi-GPU lease: gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_multi.sh <tag> <N> [steps...]'
#   steps: tests | bench | w4[:candidates] | w5
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 "$@"; }
for step in "$@"; do
  name=${step%%:*}; arg=${step#*:}; [ "$arg" == "$step" ] && arg=""
  case $name in
    tests) timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py -x -q -m gpu -k "two_gpu or acqmax_many" > gpurun_out/${tag}_tests.log 2>&1; tail -5 gpurun_out/${tag}_tests.log ;;
    bench) timeout 900 $(declare -f run >/dev/null; echo) python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; tail -2 gpurun_out/${tag}_bench_n$n.err; python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_n$n.json').read().strip().splitlines()[-1])
print({k:j.get(k) for k in ('value','ms_per_step','n_gpus','argmax_is_max_over_ranks')}); print('e2e',j.get('e2e',{}).get('value')); print('strong',j.get('strong'))
PY
    ;;
    w4) c=${arg:-16777216}; timeout 1400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $n --workload 4 --candidates $c --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_w4_n$n.json 2> gpurun_out/${tag}_w4_n$n.err; tail -2 gpurun_out/${tag}_w4_n$n.err; cut -c1-700 gpurun_out/${tag}_w4_n$n.json ;;
    w5) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n --workload 5 --steps 5 --warmup 1 > gpurun_out/${tag}_w5_n$n.json 2> gpurun_out/${tag}_w5_n$n.err; tail -2 gpurun_out/${tag}_w5_n$n.err; cut -c1-1800 gpurun_out/${tag}_w5_n$n.json ;;
  esac
done
