set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py 2>&1 | grep '"N": 2048\|"N": 4096'
python bench.py --suite 2>/dev/null | grep "config5\|maximizeEI_N2048\|config3" | cut -c1-200
