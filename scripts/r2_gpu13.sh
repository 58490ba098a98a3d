#!/bin/bash
O=gpurun_out; mkdir -p $O
for dbg in 0 16 32 48; do
  echo "== AVI_TC_DBG=$dbg"
  AVI_TC_DBG=$dbg timeout 60 python scripts/step_prof.py 10000 > $O/g13_prof_warm_dbg$dbg.txt 2>&1; grep -E "^ ?(7|8|9|11|12|13) " $O/g13_prof_warm_dbg$dbg.txt
done
