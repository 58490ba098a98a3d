#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 60 python scripts/step_timeline.py 48 10000 1 2>&1 | tee $O/g12_timeline.txt | tail -7
echo "== bench AVI_NO_GRAPH=1"
AVI_NO_GRAPH=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'us', round(1e3*d['ms_per_step'],2), 'warm', round(d['value_l2_resident']), 'us', round(1e3*d['ms_per_step_l2_resident'],2), 'e2e', round(d['e2e']['value']), d['e2e'].get('breakdown'))"
echo "== bench AVI_PDL=0"
AVI_PDL=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'us', round(1e3*d['ms_per_step'],2), 'warm', round(d['value_l2_resident']), 'us', round(1e3*d['ms_per_step_l2_resident'],2), 'e2e', round(d['e2e']['value']), d['e2e'].get('breakdown'))"
