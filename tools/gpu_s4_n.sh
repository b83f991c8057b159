set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py 2>&1 | grep '"N": 2048\|"N": 4096'
IBO_K2_DEEP=0 python tools/small_batch_latency.py 2>&1 | grep '"N": 2048\|"N": 4096' | grep '"M": 4,\|"M": 24\|"M": 64'
python bench.py --suite 2>/dev/null | grep "config5\|maximizeEI_N2048\|config3" | cut -c1-200
IBO_DIRECT_TIMING=1 python bench.py --gpus 1 --workload 5 --steps 3 --warmup 2 2> gpurun_out/s4n_w5.err | cut -c1-160; grep "ibo_acqmax" gpurun_out/s4n_w5.err | tail -1
