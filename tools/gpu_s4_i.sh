set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py 2>&1 | grep '"N": 50'
python bench.py --suite 2>/dev/null | grep "config1" | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tiny --csv --log-file gpurun_out/s4_tiny_launches.csv python tools/tiny_ncu.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/s4_tiny_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for r in rows[1:13]: print(r[ki][:40], r[gi], r[vi])
PY
