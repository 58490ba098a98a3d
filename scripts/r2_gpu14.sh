#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g14_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/g14_tests.log
for bmn in 1 0; do
  echo "== AVI_TC_BMN=$bmn"
  AVI_TC_BMN=$bmn timeout 60 python scripts/step_prof.py 10000 > $O/g14_prof_warm_bmn$bmn.txt 2>&1; grep -E "^ ?(7|8|9|11|13|15|16|17|19|21|23) " $O/g14_prof_warm_bmn$bmn.txt
done
timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g14_bench.json 2> $O/g14_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g14_bench.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "parity", d.get("parity"))
    a=d["alt_precision"]; print("x3 value", round(a["value"]), "warm", round(a["value_l2_resident"]), "e2e", round(a["e2e"]["value"]), a.get("parity"))
except Exception as e: print("parse failed", e); print(open("gpurun_out/g14_bench.err").read()[-1500:])
PY
