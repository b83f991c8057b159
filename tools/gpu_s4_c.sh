# 2 GPUs: collectives + sharded DIRECT test, both scaling bench lines at N=2, config #5 sharded
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$LAUNCH bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/s4_scale_w2_n2.json 2> gpurun_out/s4_scale_w2_n2.err; cut -c1-330 gpurun_out/s4_scale_w2_n2.json; tail -3 gpurun_out/s4_scale_w2_n2.err
IBO_DIRECT_TIMING=1 $LAUNCH bench.py --gpus 2 --workload 5 --steps 3 --warmup 2 > gpurun_out/s4_w5_n2.json 2> gpurun_out/s4_w5_n2.err; cut -c1-1500 gpurun_out/s4_w5_n2.json; grep ibo_acqmax gpurun_out/s4_w5_n2.err | tail -4
IBO_DIRECT_TIMING=1 python bench.py --gpus 1 --workload 5 --steps 3 --warmup 2 > gpurun_out/s4_w5_n1.json 2> gpurun_out/s4_w5_n1.err; cut -c1-900 gpurun_out/s4_w5_n1.json; grep ibo_acqmax gpurun_out/s4_w5_n1.err | tail -2
for sm in 128 256 512; do IBO_SHARD_MIN=$sm IBO_DIRECT_TIMING=1 $LAUNCH bench.py --gpus 2 --workload 5 --steps 3 --warmup 1 2>gpurun_out/s4_w5_sm$sm.err | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('shard_min $sm', j['value'], j['single_gpu_unsharded_wall_ms'], j['same_result_as_unsharded'])"; grep ibo_acqmax gpurun_out/s4_w5_sm$sm.err | tail -1; done
$LAUNCH bench.py --gpus 2 --workload 4 --candidates 1048576 --steps 2 --warmup 1 > gpurun_out/s4_scale_w4_n2.json 2> gpurun_out/s4_scale_w4_n2.err; cut -c1-420 gpurun_out/s4_scale_w4_n2.json; tail -3 gpurun_out/s4_scale_w4_n2.err
