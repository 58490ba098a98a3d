#!/bin/bash
# A/B of environment switches on the C3 bench line: scripts/c3_ab.sh "VAR=val ..." "VAR=val ..." ...
O=gpurun_out; mkdir -p $O
for cfg in "$@"; do
  env $cfg AVI_TC_DEBUG=0 timeout 300 python bench.py --config c3 --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg:', round(d['value']), round(d['value_l2_resident']), d['final_elbo'], d['launches_per_step'])"
done
