set -x
python bench.py --suite > gpurun_out/suite_r01.jsonl 2> gpurun_out/suite_r01.err; cat gpurun_out/suite_r01.jsonl; tail -5 gpurun_out/suite_r01.err
ncu --set full --clock-control none --import-source on -k regex:trigemm -s 2 -c 1 -o gpurun_out/prof_k2_r01b python bench.py --steps 1 --warmup 3 --candidates 37888 --no-cpu-baseline > gpurun_out/ncu_full_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kstar -s 2 -c 1 -o gpurun_out/prof_k1_r01 python bench.py --steps 1 --warmup 3 --candidates 37888 --no-cpu-baseline > gpurun_out/ncu_full_k1.log 2>&1
