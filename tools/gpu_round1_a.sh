set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_r01_a.err; tail -c 3000 gpurun_out/bench_r01_a.json; tail -5 gpurun_out/bench_r01_a.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --candidates 151552 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trigemm -s 2 -c 1 -o gpurun_out/prof_k2_r01 python bench.py --steps 1 --warmup 3 --candidates 37888 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
