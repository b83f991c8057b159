set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python tools/small_batch_latency.py 2>&1 | grep '"N": 50'
python bench.py --suite 2>/dev/null | grep "config1\|config3" | cut -c1-400
