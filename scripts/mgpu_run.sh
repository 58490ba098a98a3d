#!/bin/bash
# multi-GPU parity check + bench variants in one gpurun --gpus N call.  usage: scripts/mgpu_run.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multigpu_check (LL exchange)"; timeout 150 $TR --master-port 29511 tests/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -5
echo "== multigpu_check (AVI_COMM_LL=0: pull protocol)"; AVI_COMM_LL=0 timeout 150 $TR --master-port 29512 tests/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'n_gpus', d['n_gpus'], 'cold', round(d['value']), 'warm', round(d['value_l2_resident']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])"; }
for shard in samples rows; do
  timeout 200 $TR --master-port 29513 bench.py --gpus $N --steps 300 --warmup 30 --no-cpu-baseline --shard $shard 2>/dev/null | summ "LL shard=$shard"
  AVI_COMM_LL=0 timeout 200 $TR --master-port 29514 bench.py --gpus $N --steps 300 --warmup 30 --no-cpu-baseline --shard $shard 2>/dev/null | summ "pull shard=$shard"
done
timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>/dev/null | summ "single"
