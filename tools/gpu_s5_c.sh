# session 5: K2 narrow-shape sweep on config #5 now that batches are twice as large (~323 points)
set -x
for mt in 0 8 4 2 1; do
  IBO_NARROW_MT=$mt IBO_DIRECT_TIMING=1 python bench.py --workload 5 --steps 3 --warmup 2 2> gpurun_out/s5c_w5_mt$mt.err | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('MT=$mt', j['value'])"
  grep -E "ibo_acqmax" gpurun_out/s5c_w5_mt$mt.err | tail -1
done
IBO_DEBUG_PLAN=1 python bench.py --workload 5 --steps 1 --warmup 1 2>&1 | grep plan_narrow | sort | uniq -c | sort -k1nr | head -30
