#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g2_fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -3 $O/g2_fused_tests.log
for rows in 10000 1250; do
  timeout 200 python scripts/step_prof.py $rows > $O/g2_prof_$rows.txt 2>&1; cat $O/g2_prof_$rows.txt
  AVI_STEP_NC=1 timeout 200 python scripts/step_prof.py $rows > $O/g2_prof_nc_$rows.txt 2>&1; echo "--- NC experiment"; cat $O/g2_prof_nc_$rows.txt
done
AVI_FUSED_STEP=1 timeout 600 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g2_bench.json 2> $O/g2_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/g2_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "alt", {k: (round(v) if isinstance(v,float) else v) for k,v in d.get("alt_precision",{}).items() if k in ("value","value_l2_resident","error")})
PY
