set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/small_batch_latency.py
IBO_DIRECT_TIMING=1 python bench.py --suite 2>gpurun_out/suite_s2g.err | grep -E "config1|maximizeEI_N2048|config3|config5" | cut -c1-330
grep ibo_acqmax gpurun_out/suite_s2g.err | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-1700
