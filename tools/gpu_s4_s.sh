IBO_DEBUG_PLAN=1 python tools/small_batch_latency.py 2>&1 | grep "plan_narrow\|4096" | grep -v '"N": 50\|"N": 2048' | tail -14
IBO_DEBUG_PLAN=1 IBO_DIRECT_TIMING=1 python bench.py --workload 5 --steps 1 --warmup 1 2>&1 | grep "plan_narrow" | sort | uniq -c | sort -rn | head -40
