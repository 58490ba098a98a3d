#!/bin/bash
# Round-2 ncu evidence in one gpurun call: launch list of the bench command + `--set full` captures of the whole-iteration
# kernel (cold = ncu's cache flush per launch, warm = --cache-control none) and of the C3 (full-rank) iteration's kernels.
# Reports land in gpurun_out/; scripts/ncu_summary.py turns them into the text summaries under profiles/.
O=gpurun_out; mkdir -p $O
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/r2_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 $NCU --set full --import-source on -k regex:k_glm_mf_step --launch-skip 3 -c 3 -f -o $O/r2_step_cold python scripts/profile_steps.py 8 > $O/r2_step_cold.log 2>&1; echo "cold rc=$?"
timeout 600 $NCU --set full --import-source on --cache-control none -k regex:k_glm_mf_step --launch-skip 3 -c 3 -f -o $O/r2_step_warm python scripts/profile_steps.py 8 > $O/r2_step_warm.log 2>&1; echo "warm rc=$?"
timeout 600 $NCU --metrics gpu__time_duration.sum --launch-skip 40 -c 40 --csv --log-file $O/r2_launches_c3.csv python scripts/profile_steps.py 6 fullrank > $O/r2_c3_under_ncu.log 2>&1; echo "c3 rc=$?"
python scripts/ncu_summary.py $O/r2_step_cold.ncu-rep $O/r2_step_warm.ncu-rep > $O/r2_ncu_full_step_kernel_v2_summary.txt 2>&1
ls -la $O | grep r2_
