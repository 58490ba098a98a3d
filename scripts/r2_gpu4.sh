#!/bin/bash
O=gpurun_out; mkdir -p $O
AVI_TC_AMN=1 timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g4_fused_tests_amn1.log 2>&1; echo "fused tests amn=1 rc=$?"; tail -4 $O/g4_fused_tests_amn1.log
AVI_TC_AMN=0 timeout 600 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g4_fused_tests_amn0.log 2>&1; echo "fused tests amn=0 rc=$?"; tail -2 $O/g4_fused_tests_amn0.log
for rows in 10000 1250; do
  timeout 200 python scripts/step_prof.py $rows > $O/g4_prof_$rows.txt 2>&1; cat $O/g4_prof_$rows.txt
done
for amn in 0 1; do
AVI_TC_AMN=$amn timeout 600 python bench.py --steps 200 --warmup 20 --no-extras > $O/g4_bench_amn$amn.json 2> $O/g4_bench_amn$amn.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/g4_bench_amn$amn.json") if l.startswith("{")][-1])
print("amn=$amn value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "parity", d.get("parity"), "alt", {k: (round(v) if isinstance(v,float) else v) for k,v in d.get("alt_precision",{}).items() if k in ("value","value_l2_resident","error","parity")})
PY
done
# ncu: launch list of the bench command + one --set full capture of the iteration kernel (source-level)
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k regex:k_glm_mf_step --launch-skip 6 -c 1 -f -o $O/r2_step_full python scripts/profile_steps.py 10 > $O/g4_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 $O/g4_ncu_full.log
