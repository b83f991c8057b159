set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/small_batch_latency.py 2>&1 | grep -v "^$" | grep '"N": 4096\|"N": 2048'
IBO_DIRECT_TIMING=1 python bench.py --gpus 1 --workload 5 --steps 3 --warmup 2 2> gpurun_out/s4f_w5.err | cut -c1-160; grep "ibo_acqmax" gpurun_out/s4f_w5.err | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kstar_mma -s 3 -c 1 -o gpurun_out/s4_k1_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --candidates 151552 > gpurun_out/s4_ncu_k1.log 2>&1
tail -2 gpurun_out/s4_ncu_k1.log | cut -c1-200
