# session 5: int8 path -- accuracy with balanced digits, then timing at the headline configuration
set -x
timeout 120 python tools/research/i8_check.py 2>&1 | tail -12
IBO_INT8=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read())
print('INT8 value', j['value'], 'e2e', j['e2e']['value'], 'ms/step', j['ms_per_step'], j['kernel_ms_per_step'], 'best', j['best'])"
