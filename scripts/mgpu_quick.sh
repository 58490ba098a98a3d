#!/bin/bash
# multi-GPU parity check + one M-axis bench in one gpurun --gpus N call.  usage: scripts/mgpu_quick.sh N
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multigpu_check world=$N"; timeout 150 $TR --master-port 29511 tests/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -4
echo "== bench --gpus $N --shard samples"
timeout 150 $TR --master-port 29513 bench.py --gpus $N --steps 300 --warmup 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('n_gpus', d['n_gpus'], 'cold', round(d['value']), 'warm', round(d['value_l2_resident']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])"
