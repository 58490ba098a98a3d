#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g7_fused_tests.log 2>&1; echo "fused tests rc=$?"; tail -12 $O/g7_fused_tests.log
AVI_DRAW_AHEAD=0 timeout 200 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g7_fused_tests_noahead.log 2>&1; echo "fused tests (no draw-ahead) rc=$?"; tail -2 $O/g7_fused_tests_noahead.log
for rows in 10000 1250; do
  timeout 60 python scripts/step_prof.py $rows > $O/g7_prof_$rows.txt 2>&1; cat $O/g7_prof_$rows.txt
done
for a in 0 1; do
AVI_DRAW_AHEAD=$a timeout 300 python bench.py --steps 200 --warmup 20 --no-extras > $O/g7_bench_a$a.json 2> $O/g7_bench_a$a.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g7_bench_a$a.json") if l.startswith("{")][-1])
    print("ahead=$a value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "us", round(1e3*d["ms_per_step_l2_resident"],2), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "parity", d.get("parity"))
except Exception as e: print("ahead=$a parse failed", e); print(open("gpurun_out/g7_bench_a$a.err").read()[-1500:])
PY
done
AVI_NO_GRAPH=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('NO_GRAPH value', round(d['value']), 'warm', round(d['value_l2_resident']), 'e2e', round(d['e2e']['value']), d['e2e'].get('breakdown'))"
