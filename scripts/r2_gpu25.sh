#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 120 python scripts/repro_first_call.py 4 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > $O/g25_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/g25_tests.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g25_bench.json 2> $O/g25_bench.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/g25_bench.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), {k: round(v,2) for k,v in d["e2e"]["breakdown"].items()}, "final_elbo", d["final_elbo"])
PY
