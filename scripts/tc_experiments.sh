#!/bin/bash
# In-kernel phase timeline of the tcgen05 kernels under the AVI_TC_DBG timing experiments (one gpurun call).
# bits: 1 no operand loads, 2 no MMAs, 4 sleeping epilogue waiters, 8 free-running MMA issue
run() {
  echo "== AVI_TC_NT=$1 AVI_TC_DBG=$2"
  local t0=$(date +%s.%N)
  AVI_NO_GRAPH=1 AVI_TC_NT=$1 AVI_TC_DBG=$2 AVI_TC_PROF=6 timeout 60 python scripts/prof_phases.py 2>&1 | tail -4
  echo "   rc=$? $(echo "$(date +%s.%N) - $t0" | bc) s"
}
run 0 0; run 0 1; run 0 4; run 0 8; run 0 9; run 256 0; run 256 1
