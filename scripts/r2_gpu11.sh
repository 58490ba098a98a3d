#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g11_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/g11_tests.log
for pt in 1 0; do
  echo "== AVI_PARAM_TOUCH=$pt"
  AVI_PARAM_TOUCH=$pt timeout 90 python scripts/step_prof.py 10000 12 cold > $O/g11_prof_cold_pt$pt.txt 2>&1; grep -E "^#|^ ?(1|2|3|7|8|11|13|15|16|19|2[0-6]) " $O/g11_prof_cold_pt$pt.txt
  AVI_PARAM_TOUCH=$pt timeout 60 python scripts/step_prof.py 10000 > $O/g11_prof_warm_pt$pt.txt 2>&1; grep -E "^#|^ ?(1|2|3|7|8|11|13|15|16|19|2[0-6]) " $O/g11_prof_warm_pt$pt.txt
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g11_bench.json 2> $O/g11_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g11_bench.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "launches", d["launches_per_step"])
except Exception as e: print("parse failed", e); print(open("gpurun_out/g11_bench.err").read()[-1500:])
PY
