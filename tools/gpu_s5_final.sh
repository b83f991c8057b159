# session 5, final: full GPU suite, smoke, default bench (FP64 DMMA arm + int8 leg), ncu launch list and one full capture of the int8 K2
set -x
( time timeout 200 python -m pytest tests -x -q -m gpu ) > gpurun_out/s5f_pytest.log 2>&1; tail -4 gpurun_out/s5f_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 150 python bench.py --steps 5 --warmup 3 > gpurun_out/s5f_bench_w2.json 2> gpurun_out/s5f_bench.err
python -c "
import json
j = json.load(open('gpurun_out/s5f_bench_w2.json'))
print('main', j['value'], j['e2e']['value'], j['roofline']['frac'], j['maximizeEI_wall_ms'])
print('int8', json.dumps(j.get('int8_emulation'))[:700])"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s5f_launches_int8.csv python bench.py --int8 --steps 1 --warmup 1 --no-cpu-baseline --candidates 151552 > gpurun_out/s5f_ncu_list.log 2>&1
tail -1 gpurun_out/s5f_ncu_list.log | cut -c1-300
timeout 150 ncu --set full --clock-control none --import-source on -k regex:trigemm_i8 -s 2 -c 1 -o gpurun_out/s5f_k2i_full -f python bench.py --int8 --steps 1 --warmup 1 --no-cpu-baseline --candidates 151552 > gpurun_out/s5f_ncu_k2i.log 2>&1
tail -2 gpurun_out/s5f_ncu_k2i.log | cut -c1-200
ls -la gpurun_out/ | tail -5
