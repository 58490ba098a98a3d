#!/bin/bash
# A/B of environment switches on any bench configuration: scripts/cfg_ab.sh <config> <steps> "VAR=val ..." ...
CFG=$1; K=$2; shift 2
for cfg in "$@"; do
  env $cfg timeout 600 python bench.py --config $CFG --steps $K --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$CFG $cfg:', round(d['value']), round(d['value_l2_resident']), d['final_elbo'], d['launches_per_step'])"
done
