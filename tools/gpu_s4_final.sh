# final artefacts of the session: tests, smoke, bench (ours + reference arm), launch list, suite, config #5
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench_w2.json 2> gpurun_out/final_bench_w2.err; cut -c1-400 gpurun_out/final_bench_w2.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/final_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final_ncu_bench.log 2>&1
python bench.py --suite > gpurun_out/final_suite.json 2>/dev/null; cut -c1-260 gpurun_out/final_suite.json
python bench.py --workload 5 --steps 5 --warmup 2 > gpurun_out/final_w5_n1.json 2>/dev/null; cut -c1-500 gpurun_out/final_w5_n1.json
