#!/bin/bash
# quick regression of a kernel change: determinism repro, fused-kernel tests, warm stamps, C2 bench without the extras
O=gpurun_out; mkdir -p $O
timeout 120 python scripts/repro_first_call.py 3 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/chk_tests.log 2>&1; echo "fused tests rc=$?"; tail -2 $O/chk_tests.log
timeout 60 python scripts/step_prof.py 10000 > $O/chk_prof_warm.txt 2>&1; grep -E "^ ?(1|3|7|8|11|12|13|15|16|19|21|23) " $O/chk_prof_warm.txt
timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/chk_bench.json 2> $O/chk_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/chk_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["e2e"]["breakdown"].items()}, "final_elbo", d["final_elbo"])
a=d["alt_precision"]; print("x3", round(a["value"]), round(a["value_l2_resident"]), round(a["e2e"]["value"]), a["final_elbo"])
PY
