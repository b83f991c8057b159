set -x
timeout 300 ncu --set full --import-source on --clock-control none --cache-control none -k regex:trigemm -s 4 -c 1 -o gpurun_out/s4_k2mid_full -f python tools/k2_tiny_ncu.py > /dev/null 2>&1
ls -la gpurun_out/s4_k2mid_full.ncu-rep
