# session 5: full GPU suite + bench after the one-batch-per-iteration DIRECT (with speculation for unclean sides)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/s5d_pytest.log 2>&1; tail -3 gpurun_out/s5d_pytest.log
IBO_DIRECT_TIMING=1 python bench.py --steps 5 --warmup 3 > gpurun_out/s5d_bench_w2.json 2> gpurun_out/s5d_bench.err
python -c "import json; j=json.load(open('gpurun_out/s5d_bench_w2.json')); print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['maximizeEI_wall_ms'])"
grep -E "ibo_acqmax" gpurun_out/s5d_bench.err | tail -2
IBO_DIRECT_TIMING=1 python bench.py --workload 5 --steps 5 --warmup 2 > gpurun_out/s5d_w5_n1.json 2> gpurun_out/s5d_w5.err; cut -c1-200 gpurun_out/s5d_w5_n1.json
grep -E "ibo_acqmax|run_direct" gpurun_out/s5d_w5.err | tail -2
python bench.py --suite > gpurun_out/s5d_suite.json 2>/dev/null; cut -c1-1500 gpurun_out/s5d_suite.json
