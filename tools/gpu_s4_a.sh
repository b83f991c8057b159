# session 4, call A: re-validate HEAD on a B200 and collect the profile artefacts for the current code
set -x
nvidia-smi -L
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 5 --warmup 3 > gpurun_out/s4_bench_w2.json 2> gpurun_out/s4_bench_w2.err; cut -c1-1200 gpurun_out/s4_bench_w2.json; tail -3 gpurun_out/s4_bench_w2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s4_bench_ref.json 2> gpurun_out/s4_bench_ref.err; cut -c1-600 gpurun_out/s4_bench_ref.json
# ncu launch list of the same command (reduced steps; per-launch times are cold-cache / serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s4_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/s4_ncu_bench.log 2>&1
tail -2 gpurun_out/s4_ncu_bench.log | cut -c1-300
# one full capture of the dominant kernel (K2) for dram traffic
timeout 600 ncu --set full --clock-control none --import-source on -k regex:trigemm -s 4 -c 1 -o gpurun_out/s4_k2_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --candidates 151552 > gpurun_out/s4_ncu_k2.log 2>&1
tail -2 gpurun_out/s4_ncu_k2.log | cut -c1-300
python bench.py --suite > gpurun_out/s4_suite.json 2> gpurun_out/s4_suite.err; cut -c1-400 gpurun_out/s4_suite.json
