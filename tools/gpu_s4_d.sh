set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py 2>&1 | grep -v "^$"
IBO_DIRECT_TIMING=1 python bench.py --gpus 1 --workload 5 --steps 3 --warmup 2 > gpurun_out/s4d_w5_n1.json 2> gpurun_out/s4d_w5_n1.err; cut -c1-200 gpurun_out/s4d_w5_n1.json; grep "ibo_acqmax\|run_direct" gpurun_out/s4d_w5_n1.err | tail -4
