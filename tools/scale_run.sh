#!/bin/bash
# weak-scaling check of bench.py at N GPUs of one box (driver launches it the same way)
# usage: tools/scale_run.sh N [workload] [batch per GPU] [steps]   -- give c4 / c5 a bounded batch: at their default
# per-GPU batch one c5 step is ~30 s and a 25-step run costs more than 10 minutes on EVERY GPU of the box
N=$1; WL=${2:-c2}; BATCH=${3:-0}; STEPS=${4:-20}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload $WL --batch $BATCH --steps $STEPS --warmup 5 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('N=%d %s | traj-steps/s %.3e | step %.3f ms | fwd %.3f bwd %.3f | launches %d' % (d['n_gpus'], d['config']['workload'], d['value'], d['ms_per_step'], d['roofline']['fwd_kernel_ms'], d['roofline']['kernel_ms'], d['gpu_launches']))"
