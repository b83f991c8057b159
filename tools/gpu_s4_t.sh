set -x
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$LAUNCH bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | grep '^{' | cut -c1-260
$LAUNCH bench.py --gpus 2 --impl reference --steps 1 --warmup 0 2>/dev/null | grep '^{' | cut -c1-160
$LAUNCH bench.py --gpus 2 --workload 5 --steps 3 --warmup 2 2>/dev/null | grep '^{' | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('w5 n2', j['value'], j['single_gpu_unsharded_wall_ms'], j['same_result_as_unsharded'])"
