# session 5: re-validate HEAD on one B200 within the remaining GPU budget (tests, smoke, bench, reference arm, config #5)
set -x
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -x -q -m gpu ) > gpurun_out/s5_pytest.log 2>&1; tail -4 gpurun_out/s5_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 120 python bench.py --steps 5 --warmup 3 > gpurun_out/s5_bench_w2.json 2> gpurun_out/s5_bench_w2.err; cut -c1-300 gpurun_out/s5_bench_w2.json
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s5_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/s5_bench_ref.json
timeout 60 python bench.py --workload 5 --steps 5 --warmup 2 > gpurun_out/s5_w5_n1.json 2>/dev/null; cut -c1-300 gpurun_out/s5_w5_n1.json
