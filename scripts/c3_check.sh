#!/bin/bash
# Full-rank iteration after a kernel change: the GPU test-suite, then the C3 bench line and warm launch list.
O=gpurun_out; mkdir -p $O; TAG=${1:-c3}
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/${TAG}_tests.log
bash scripts/c3_prof.sh $TAG
AVI_FR_TILED_UPDATE=0 timeout 300 python bench.py --config c3 --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('untiled update:', d['value'], d['value_l2_resident'], d['final_elbo'], d['launches_per_step'])"
AVI_GRAPH_UNROLL=1 timeout 300 python bench.py --config c3 --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph unroll 1:', d['value'], d['value_l2_resident'], d['final_elbo'], d['launches_per_step'])"
