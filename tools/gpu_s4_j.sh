set -x
python -m pytest tests/test_gpu_tiny.py -x -q -k "prior" 2>&1 | grep -E "^E  |Error" | cut -c1-400 | head -8
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tiny_fused -s 4 -c 1 -o gpurun_out/s4_tiny_full -f python tools/tiny_ncu.py > /dev/null 2>&1
ls -la gpurun_out/s4_tiny_full.ncu-rep
