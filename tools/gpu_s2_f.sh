tools/launch_floor
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kstar|trigemm|epilogue" --csv --log-file gpurun_out/tiny_launches.csv python tools/tiny_ncu.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tiny_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:40], r[7], r[8], r[-1])
PY
