# N GPUs of one box: the driver's scaling command for the headline workload, plus configs #4 (reduced) and #5
N=$1
set -x
nvidia-smi -L | wc -l
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$LAUNCH bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/s4e_w2_n$N.json 2> gpurun_out/s4e_w2_n$N.err; grep '^{' gpurun_out/s4e_w2_n$N.json | cut -c1-330; tail -2 gpurun_out/s4e_w2_n$N.err
$LAUNCH bench.py --gpus $N --impl reference --steps 1 --warmup 0 2>/dev/null | grep '^{' | cut -c1-200
IBO_DIRECT_TIMING=1 $LAUNCH bench.py --gpus $N --workload 5 --steps 3 --warmup 2 > gpurun_out/s4e_w5_n$N.json 2> gpurun_out/s4e_w5_n$N.err; grep '^{' gpurun_out/s4e_w5_n$N.json | cut -c1-1200; grep ibo_acqmax gpurun_out/s4e_w5_n$N.err | tail -3
$LAUNCH bench.py --gpus $N --workload 4 --candidates 4194304 --steps 1 --warmup 1 > gpurun_out/s4e_w4_n$N.json 2> gpurun_out/s4e_w4_n$N.err; grep '^{' gpurun_out/s4e_w4_n$N.json | cut -c1-420; tail -2 gpurun_out/s4e_w4_n$N.err
