#!/bin/bash
# final single-GPU run of the round: phase stamps, smoke, default bench (all configs + CPU baseline), reference arm
O=gpurun_out; mkdir -p $O
for rows in 10000 1250; do
  timeout 90 python scripts/step_prof.py $rows 12 cold > $O/r2_fused_phase_stamps_final_rows${rows}_cold.txt 2>&1
  timeout 60 python scripts/step_prof.py $rows > $O/r2_fused_phase_stamps_final_rows${rows}_warm.txt 2>&1
done
grep -E "^#|^ ?(1|3|7|8|11|13|15|16|19|21|2[2-6]) " $O/r2_fused_phase_stamps_final_rows10000_warm.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_final.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["e2e"]["breakdown"].items()})
print("roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac","traffic","kernel_ms")}, "clocks", d["clocks"], "launches", d["gpu_launches"], d["launches_per_step"])
print("sample", d["roofline"].get("sample_kernel_hbm"))
print("cpu", d["cpu_baseline"]); print("parity", d["parity"])
a=d["alt_precision"]; print("x3", round(a["value"]), round(a["value_l2_resident"]), round(a["e2e"]["value"]), a.get("parity"), a["roofline"]["frac"])
for k,v in d["configs"].items(): print(k, {kk: (round(vv,1) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ("value","ms_per_step","value_l2_resident","launches_per_step","final_elbo","error","setup_s")}, (v.get("roofline") or {}).get("frac"))
r=json.loads([l for l in open("gpurun_out/r2_bench_reference_arm.json") if l.startswith("{")][-1]); print("ref", r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"].get("reference_shaped_steps_per_s"))
PY
