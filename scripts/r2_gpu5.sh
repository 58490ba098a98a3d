#!/bin/bash
O=gpurun_out; mkdir -p $O
AVI_FUSED_KERNEL=2 timeout 150 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g5_fused_tests_k2.log 2>&1; echo "fused tests kernel=2 rc=$?"; tail -12 $O/g5_fused_tests_k2.log
AVI_FUSED_KERNEL=1 timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g5_fused_tests_k1.log 2>&1; echo "fused tests kernel=1 rc=$?"; tail -2 $O/g5_fused_tests_k1.log
for rows in 10000 1250; do
  AVI_FUSED_KERNEL=2 timeout 60 python scripts/step_prof.py $rows > $O/g5_prof_k2_$rows.txt 2>&1; cat $O/g5_prof_k2_$rows.txt
done
for k in 1 2; do
AVI_FUSED_KERNEL=$k timeout 200 python bench.py --steps 200 --warmup 20 --no-extras > $O/g5_bench_k$k.json 2> $O/g5_bench_k$k.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g5_bench_k$k.json") if l.startswith("{")][-1])
    print("kernel=$k value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "parity", d.get("parity"))
except Exception as e: print("kernel=$k parse failed", e); print(open("gpurun_out/g5_bench_k$k.err").read()[-1500:])
PY
done
