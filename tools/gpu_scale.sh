# usage: bash tools/gpu_scale.sh N   -- both bench lines on N GPUs of one box (run under gpurun --gpus N)
N=$1
set -x
nvidia-smi -L | head -8
if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; fi
$LAUNCH bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_w2_n$N.json 2> gpurun_out/scale_w2_n$N.err; cut -c1-330 gpurun_out/scale_w2_n$N.json; tail -3 gpurun_out/scale_w2_n$N.err
$LAUNCH bench.py --gpus $N --workload 4 --steps ${STEPS4:-3} --warmup 1 > gpurun_out/scale_w4_n$N.json 2> gpurun_out/scale_w4_n$N.err; cut -c1-420 gpurun_out/scale_w4_n$N.json; tail -3 gpurun_out/scale_w4_n$N.err
