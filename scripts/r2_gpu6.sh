#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/g6_all_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -5 $O/g6_all_tests.log
timeout 900 python bench.py --steps 200 --warmup 20 > $O/g6_bench_full.json 2> $O/g6_bench_full.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/g6_bench_full.json") if l.startswith("{")][-1])
    print("c2 value", round(d["value"]), "us", round(1e3*d["ms_per_step"],2), "warm", round(d["value_l2_resident"]), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"]["frac"], d["roofline"]["kernel_ms"], "cpu", d.get("cpu_baseline",{}).get("value"), "parity", d.get("parity"))
    print("alt", json.dumps(d.get("alt_precision"))[:900])
    for k,v in d.get("configs",{}).items(): print(k, json.dumps(v)[:600])
    print("sample_large", d["roofline"].get("sample_kernel_hbm"))
except Exception as e: print("parse failed", e); print(open("gpurun_out/g6_bench_full.err").read()[-2500:])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/g6_ref.json 2> $O/g6_ref.err; echo "ref rc=$?"; cut -c1-700 $O/g6_ref.json
