#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/g17_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/g17_tests.log
for ng in 0 1; do
AVI_NO_GRAPH=$ng timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('NO_GRAPH=$ng value', round(d['value']), 'us', round(1e3*d['ms_per_step'],2), 'warm', round(d['value_l2_resident']), 'e2e', round(d['e2e']['value']), {k: round(v,2) for k,v in d['e2e']['breakdown'].items()})"
done
