# chunk-size sweep of the scoring pipeline (config #2), run under gpurun
for ct in 37 74 148 296 592; do
  echo "== IBO_CHUNK_TILES=$ct"
  IBO_CHUNK_TILES=$ct python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g evals/s  ms/step %.2f  frac %.3f  k1 %.2f k2 %.2f k3 %.2f  e2e %.4g' % (j['value'], j['ms_per_step'], j['roofline']['frac'], j['kernel_ms_per_step']['k1_kstar'], j['kernel_ms_per_step']['k2_trigemm'], j['kernel_ms_per_step']['k3_epilogue'], j['e2e']['value']))"
done
