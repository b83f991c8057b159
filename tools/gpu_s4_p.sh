set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py 2>&1 | grep '"N": 2048\|"N": 4096'
IBO_DIRECT_TIMING=1 python bench.py --gpus 1 --workload 5 --steps 3 --warmup 2 2> gpurun_out/s4p_w5.err | cut -c1-160; grep "ibo_acqmax" gpurun_out/s4p_w5.err | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline']['frac'], j['kernel_ms_per_step'], j['maximizeEI_wall_ms'])"
