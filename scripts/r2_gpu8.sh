#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_fused_step.py tests/test_gpu_parity.py -x -q > $O/g8_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/g8_tests.log
for rows in 10000 1250; do
  timeout 60 python scripts/step_prof.py $rows > $O/g8_prof_$rows.txt 2>&1; cat $O/g8_prof_$rows.txt
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-extras > $O/g8_bench.json 2> $O/g8_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g8_bench.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "parity", d.get("parity"), "sample", d["roofline"].get("sample_kernel_hbm"))
except Exception as e: print("parse failed", e); print(open("gpurun_out/g8_bench.err").read()[-1500:])
PY
