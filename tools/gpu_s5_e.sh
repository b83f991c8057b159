# session 5: first run of the experimental int8-emulated K2
set -x
timeout 120 python tools/research/i8_check.py 2>&1 | tail -40
