#!/bin/bash
# weak-scaling check of bench.py at N GPUs of one box (driver launches it the same way)
N=$1; WL=${2:-c2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $WL --steps 20 --warmup 5 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('N=%d %s | traj-steps/s %.3e | step %.3f ms | fwd %.3f bwd %.3f | launches %d' % (d['n_gpus'], d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['fwd_kernel_ms'], d['roofline']['kernel_ms'], d['gpu_launches']))"
