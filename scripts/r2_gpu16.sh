#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/g16_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/g16_tests.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g16_bench.json 2> $O/g16_bench.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g16_bench.json") if l.startswith("{")][-1])
    print("value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "launches", d["launches_per_step"])
    a=d["alt_precision"]; print("x3 value", round(a["value"]), "warm", round(a["value_l2_resident"]), "e2e", round(a["e2e"]["value"]), a["e2e"].get("breakdown"))
except Exception as e: print("parse failed", e); print(open("gpurun_out/g16_bench.err").read()[-1500:])
PY
AVI_HOST_LAMBDA=0 timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('HOST_LAMBDA=0: e2e', round(d['e2e']['value']), d['e2e'].get('breakdown'))"
