set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/tiny_throughput.py
