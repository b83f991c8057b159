set -x
python -m pytest tests/test_gpu_laplace.py tests/test_gpu_append.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
IBO_DIRECT_TIMING=1 python bench.py --suite 2>gpurun_out/suite_s2c.err | grep -E "config3|config5" | cut -c1-400
grep ibo_acqmax gpurun_out/suite_s2c.err | tail -4
