#!/bin/bash
# multi-GPU: parity check (sharded == unsharded, bitwise identical across ranks) + bench on N GPUs
#   usage: scripts/r2_mgpu.sh N [extras]     extras: also the other BASELINE.json configs in the n-axis run
N=${1:-2}; EX=${2:-}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multigpu_check.py > $O/mg${N}_check.log 2>&1; echo "multigpu_check N=$N rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM" $O/mg${N}_check.log | tail -4
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'n_gpus', d['n_gpus'], 'cold', round(d['value']), 'us', round(1e3*d['ms_per_step'],2), 'warm', round(d['value_l2_resident']), 'us', round(1e3*d['ms_per_step_l2_resident'],2), 'e2e', round(d['e2e']['value']), 'launches/step', d['launches_per_step'], (d.get('roofline') or {}).get('kernel_ms'))
        for k,v in (d.get('configs') or {}).items(): print('   ', k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','ms_per_step','value_l2_resident','launches_per_step','final_elbo','error','setup_s')}, (v.get('roofline') or {}).get('frac'))"; }
if [ -n "$EX" ]; then XF="--extras"; else XF="--no-extras"; fi
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 300 --warmup 30 $XF --shard rows 2> $O/mg${N}_bench_rows.err | tee $O/mg${N}_bench_rows.json | summ "N=$N shard=rows"
tail -3 $O/mg${N}_bench_rows.err | cut -c1-300
timeout 300 $TR --master-port 29514 bench.py --gpus $N --steps 200 --warmup 20 --no-extras --shard samples 2> $O/mg${N}_bench_samples.err | tee $O/mg${N}_bench_samples.json | summ "N=$N shard=samples"
timeout 200 $TR --master-port 29516 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2>/dev/null | cut -c1-260
