#!/bin/bash
# A/B of environment toggles inside ONE gpurun call (same box, same clocks): prints cold / warm / e2e steps per second
# usage: scripts/ab_bench.sh "VAR=a" "VAR=b" ...
run() { env $1 timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'cold', round(d['value']), 'warm', round(d['value_l2_resident']), 'e2e', round(d['e2e']['value']), d['roofline']['kernel_ms'], d['gpu_launches'])"; }
for rep in 1 2; do
  for cfg in "$@"; do run "$cfg"; done
done
