#!/bin/bash
# N-GPU: parity check (mean-field + full-rank, both axes) and the C3 (full-rank) bench line under row / sample sharding
N=${1:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multigpu_check.py > $O/mg${N}_check_v2.log 2>&1; echo "multigpu_check N=$N rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/mg${N}_check_v2.log | tail -6
for sh in rows samples; do
timeout 300 $TR --master-port 29513 bench.py --gpus $N --config c3 --steps 40 --warmup 5 --no-cpu-baseline --no-extras --shard $sh 2> $O/mg${N}_c3_$sh.err | tee $O/mg${N}_c3_$sh.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('C3 N=$N shard=$sh', round(d['value']), round(d['value_l2_resident']), d['final_elbo'], d['launches_per_step'])"
tail -2 $O/mg${N}_c3_$sh.err | cut -c1-300
done
