#!/bin/bash
# A/B of environment toggles inside ONE gpurun call (same box, same clocks): prints cold / warm / e2e steps per second
run() { timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['value_l2_resident']), round(d['e2e']['value']), d['roofline']['kernel_ms'], d['gpu_launches'])"; }
for rep in 1 2; do
  AVI_TC_PAIR=1 run "pair=1"
  AVI_TC_PAIR=0 run "pair=0"
done
