#!/bin/bash
# multi-GPU: parity check (sharded == unsharded, bitwise identical across ranks) + bench (rows / samples) on N GPUs
N=${1:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multigpu_check.py > $O/mg${N}_check.log 2>&1; echo "multigpu_check N=$N rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/mg${N}_check.log | tail -6
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'n_gpus', d['n_gpus'], 'cold', round(d['value']), 'us', round(1e3*d['ms_per_step'],2), 'warm', round(d['value_l2_resident']), 'us', round(1e3*d['ms_per_step_l2_resident'],2), 'e2e', round(d['e2e']['value']), 'launches/step', d['launches_per_step'], (d.get('roofline') or {}).get('kernel_ms'))"; }
for shard in rows samples; do
  timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 300 --warmup 30 --no-extras --shard $shard 2> $O/mg${N}_bench_$shard.err | tee $O/mg${N}_bench_$shard.json | summ "N=$N shard=$shard"
done
AVI_FUSED_STEP=0 timeout 300 $TR --master-port 29515 bench.py --gpus $N --steps 300 --warmup 30 --no-extras --shard rows 2> $O/mg${N}_bench_rows_staged.err | tee $O/mg${N}_bench_rows_staged.json | summ "N=$N staged shard=rows"
timeout 300 python bench.py --steps 300 --warmup 30 --no-extras --no-cpu-baseline 2>/dev/null | tee $O/mg${N}_bench_single.json | summ "single"
