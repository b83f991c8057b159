set -x
python -m pytest tests/test_gpu_append.py -x -q -m gpu 2>&1 | tail -15
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --suite 2>/dev/null | grep -E "addData|model_build|config3" | cut -c1-300
