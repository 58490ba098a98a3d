#!/bin/bash
# `ncu --set full` of the nine kernels of one full-rank iteration (C3), cold (ncu's per-launch cache flush) -> gpurun_out/r2_c3_full.ncu-rep
O=gpurun_out; mkdir -p $O
timeout 600 ncu --clock-control none --set full --import-source on --launch-skip 49 -c 7 -f -o $O/r2_c3_full_v2 python scripts/profile_steps.py 9 fullrank > $O/r2_c3_full_v2.log 2>&1; echo "ncu full rc=$?"
ls -la $O/r2_c3_full_v2.ncu-rep
