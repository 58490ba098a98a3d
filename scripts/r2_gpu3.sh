#!/bin/bash
O=gpurun_out; mkdir -p $O
for amn in 0 1; do
  AVI_TC_AMN=$amn timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g3_fused_tests_amn$amn.log 2>&1; echo "fused tests amn=$amn rc=$?"; tail -4 $O/g3_fused_tests_amn$amn.log
done
for rows in 10000 1250; do
  timeout 200 python scripts/step_prof.py $rows > $O/g3_prof_$rows.txt 2>&1; cat $O/g3_prof_$rows.txt
done
for amn in 0 1; do
AVI_TC_AMN=$amn timeout 600 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g3_bench_amn$amn.json 2> $O/g3_bench_amn$amn.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/g3_bench_amn$amn.json") if l.startswith("{")][-1])
print("amn=$amn value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "alt", {k: (round(v) if isinstance(v,float) else v) for k,v in d.get("alt_precision",{}).items() if k in ("value","value_l2_resident","error")})
PY
done
