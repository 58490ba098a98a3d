#!/bin/bash
# round 2, first GPU call: the whole-iteration kernel -- tests, then A/B bench against the staged path, then timelines
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/g1_smi.txt
timeout 900 python -m pytest tests/test_gpu_fused_step.py -x -q > $O/g1_fused_tests.log 2>&1; echo "fused tests rc=$?" | tee -a $O/g1_summary.txt
tail -15 $O/g1_fused_tests.log
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fused_step.py > $O/g1_all_tests.log 2>&1; echo "all gpu tests rc=$?" | tee -a $O/g1_summary.txt
tail -8 $O/g1_all_tests.log
for f in 1 0; do
  AVI_FUSED_STEP=$f timeout 600 python bench.py --steps 200 --warmup 20 --no-extras --no-cpu-baseline > $O/g1_bench_fused$f.json 2> $O/g1_bench_fused$f.err; echo "bench fused=$f rc=$?" | tee -a $O/g1_summary.txt
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/g1_bench_fused$f.json") if l.startswith("{")][-1])
    print("fused=$f value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "e2e", round(d["e2e"]["value"]), d["e2e"].get("breakdown"), "launches/step", d["launches_per_step"], "roofline", d["roofline"]["kernel_ms"], round(d["roofline"]["frac"],3))
except Exception as e: print("parse failed", e)
PY
done | tee -a $O/g1_summary.txt
AVI_TIMELINE=1 timeout 300 python scripts/step_timeline.py 48 10000 1 > $O/g1_timeline_fused.txt 2>&1
AVI_TIMELINE=1 timeout 300 python scripts/step_timeline.py 48 1250 1 > $O/g1_timeline_fused_rows1250.txt 2>&1
AVI_TIMELINE=1 timeout 300 python scripts/step_timeline.py 48 10000 0 > $O/g1_timeline_staged.txt 2>&1
cat $O/g1_timeline_fused.txt $O/g1_timeline_fused_rows1250.txt $O/g1_timeline_staged.txt
