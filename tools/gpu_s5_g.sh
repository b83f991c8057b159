# session 5: int8 path with the two-stream pipeline and the register-resident K1: tests, check script, bench (default arm + int8 leg)
set -x
timeout 200 python -m pytest tests/test_gpu_int8.py -x -q 2>&1 | tail -15
timeout 100 python tools/research/i8_check.py 2>&1 | tail -8
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/s5g_bench.json 2> gpurun_out/s5g_bench.err
python -c "
import json
j = json.load(open('gpurun_out/s5g_bench.json'))
print('main', j['value'], j['roofline']['frac'], j['roofline'].get('step'))
print('int8', json.dumps(j.get('int8_emulation'))[:900])"
