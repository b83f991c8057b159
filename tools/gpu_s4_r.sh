set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --suite 2>/dev/null | grep "model_build\|nlml\|config3\|addData" | cut -c1-220
