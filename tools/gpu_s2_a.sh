# session-2 validation of the restored checkpoint: GPU tests, smoke, bench (both arms), suite
set -x
nvidia-smi -L
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_s2_ref.json 2> gpurun_out/bench_s2_ref.err; cut -c1-600 gpurun_out/bench_s2_ref.json
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; cat gpurun_out/bench_s2.json; tail -5 gpurun_out/bench_s2.err
python bench.py --suite > gpurun_out/suite_s2.jsonl 2> gpurun_out/suite_s2.err; cut -c1-700 gpurun_out/suite_s2.jsonl; tail -5 gpurun_out/suite_s2.err
