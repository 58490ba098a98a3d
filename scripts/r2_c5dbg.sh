#!/bin/bash
N=2; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
summ() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'n_gpus', d['n_gpus'], 'value', round(d['value']), 'us', round(1e3*d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'launches/step', d['launches_per_step'], (d.get('roofline') or {}).get('kernel_ms'))"; }
timeout 200 python bench.py --config c5 --steps 200 --warmup 20 --no-extras 2>/dev/null | summ "N=1 default"
timeout 200 $TR --master-port 29513 bench.py --config c5 --gpus $N --steps 200 --warmup 20 --no-extras 2>/dev/null | summ "N=2 default"
AVI_DRAW_AHEAD=0 timeout 200 $TR --master-port 29514 bench.py --config c5 --gpus $N --steps 200 --warmup 20 --no-extras 2>/dev/null | summ "N=2 no-draw-ahead"
AVI_FUSED_STEP=0 timeout 200 $TR --master-port 29515 bench.py --config c5 --gpus $N --steps 200 --warmup 20 --no-extras 2>/dev/null | summ "N=2 staged"
AVI_NO_GRAPH=1 timeout 200 $TR --master-port 29516 bench.py --config c5 --gpus $N --steps 200 --warmup 20 --no-extras 2>/dev/null | summ "N=2 no-graph"
