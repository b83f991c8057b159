set -x
python -m pytest tests/test_gpu_append.py -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py
