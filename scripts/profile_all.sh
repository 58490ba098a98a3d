#!/bin/bash
# One gpurun call: ncu launch list of the bench command + `--set full` captures of the kernels of one iteration
# (cold = ncu's default cache flush per kernel, warm = --cache-control none) + the sampling kernel at M = 32768.
# Reports land in gpurun_out/; scripts/ncu_summary.py turns them into the text summaries under profiles/.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
K='regex:k_gemm_tc|k_sample|k_mf_finalize_update'
$NCU --set full --import-source on -k "$K" --launch-skip 3 -c 8 -f -o $O/full_cold python scripts/profile_steps.py 4 > $O/full_cold.log 2>&1
$NCU --set full --import-source on --cache-control none -k "$K" --launch-skip 3 -c 8 -f -o $O/full_warm python scripts/profile_steps.py 4 > $O/full_warm.log 2>&1
$NCU --set full --import-source on -k regex:k_sample --launch-skip 1 -c 2 -f -o $O/full_sample_large python scripts/profile_sample.py 32768 > $O/full_sample_large.log 2>&1
ls -la $O
