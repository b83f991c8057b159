set -x
mkdir -p gpurun_out
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 1 --candidates 151552 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1; tail -2 gpurun_out/r02_launches_bench.log
# one full capture of the shipped K2 (INT8) and one of the DMMA K2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trigemm_i8 -s 3 -c 1 -o gpurun_out/r02_k2i python tools/i8_bench.py 2048 6 151552 > gpurun_out/r02_ncu_k2i.log 2>&1; tail -2 gpurun_out/r02_ncu_k2i.log
timeout 600 ncu --set full --clock-control none -k regex:trigemm_kernel -s 2 -c 1 -o gpurun_out/r02_k2f python tools/i8_bench.py 2048 6 151552 > gpurun_out/r02_ncu_k2f.log 2>&1; tail -2 gpurun_out/r02_ncu_k2f.log
IBO_DIRECT_TIMING=1 timeout 300 python bench.py --workload 5 --steps 2 --warmup 1 2>&1 | grep -E "ibo_acqmax|run_direct" | tail -4
timeout 900 bash tools/sanitize.sh > gpurun_out/r02_sanitizer.log 2>&1; grep "rc=" gpurun_out/r02_sanitizer.log
