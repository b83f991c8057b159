set -x
python -m pytest tests/test_gpu_hyper.py -x -q 2>&1 | tail -15
python bench.py --suite 2>&1 | grep '"nlml"' | cut -c1-300
