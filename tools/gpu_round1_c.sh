python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/probe_build.py
for k in direct mma; do echo "== IBO_KSTAR=$k"; IBO_KSTAR=$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g evals/s  ms/step %.2f  frac %.3f  k1 %.2f k2 %.2f k3 %.2f  e2e %.4g' % (j['value'], j['ms_per_step'], j['roofline']['frac'], j['kernel_ms_per_step']['k1_kstar'], j['kernel_ms_per_step']['k2_trigemm'], j['kernel_ms_per_step']['k3_epilogue'], j['e2e']['value']))"; done
python bench.py --suite 2>&1 | cut -c1-420
