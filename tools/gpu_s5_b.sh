# session 5: DIRECT with order-independent child centres riding in the probe batch -- batches per query and wall ms
set -x
python -m pytest tests/test_gpu_golden.py tests/test_gpu_multi.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -3
IBO_DIRECT_TIMING=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2> gpurun_out/s5b_bench.err | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['maximizeEI_wall_ms'])"
grep -E "ibo_acqmax" gpurun_out/s5b_bench.err | tail -4
IBO_DIRECT_TIMING=1 python bench.py --workload 5 --steps 5 --warmup 2 2> gpurun_out/s5b_w5.err | cut -c1-200
grep -E "ibo_acqmax|run_direct" gpurun_out/s5b_w5.err | tail -3
