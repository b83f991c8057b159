# Round 2, first GPU call: the INT8 variants written blind at the end of round 1 (8-bit digits, eighth accumulator group), the
# two-stream pipeline A/B, and the re-validation of HEAD.  ~2 min of box time.
set -x
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2a_pytest.log 2>&1; tail -4 gpurun_out/r2a_pytest.log
IBO_EXPERIMENTAL_TESTS=1 timeout 200 python -m pytest tests/test_gpu_int8.py -q -k "d8 or g9 or s6" 2>&1 | tail -15
for v in 1 8 6 9; do
  IBO_INT8=$v timeout 100 python bench.py --int8 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('IBO_INT8=$v', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'], 2), j['kernel_ms_per_step'], 'best', j['best'])"
done
IBO_INT8=1 IBO_I8_PIPE=0 timeout 100 python bench.py --int8 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('no pipeline', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'], 2))"
IBO_INT8=1 IBO_I8_K1_CARVEOUT=1 timeout 100 python bench.py --int8 --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('K1 carve-out hint', 'value', round(j['value']), 'ms/step', round(j['ms_per_step'], 2))"
timeout 80 python tools/research/i8_check_8192.py 2>&1 | tail -2
