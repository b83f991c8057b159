set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/small_batch_latency.py
for mt in 4 2 1; do echo "== IBO_NARROW_MT=$mt"; IBO_NARROW_MT=$mt python tools/small_batch_latency.py | grep -v "\"N\": 50" | cut -c1-120; done
IBO_DIRECT_TIMING=1 python bench.py --suite 2>gpurun_out/suite_s2e.err | grep -E "config1|maximizeEI_N2048|config3|config5" | cut -c1-330
grep ibo_acqmax gpurun_out/suite_s2e.err | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kstar|trigemm|epilogue" --csv --log-file gpurun_out/tiny_launches.csv python tools/tiny_ncu.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tiny_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[::3] + rows[1::3] + rows[2::3]: print(r[4][:40], r[7], r[8], r[-1])
PY
