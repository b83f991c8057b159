set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s4g_bench.json 2>gpurun_out/s4g_bench.err; python -c "
import json; j=json.load(open('gpurun_out/s4g_bench.json')); print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['kernel_ms_per_step'], j['maximizeEI_wall_ms'])"
python bench.py --suite 2>/dev/null | grep "config4\|config5\|config1" | cut -c1-330
