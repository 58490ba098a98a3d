#!/bin/bash
# Full-rank iteration (C3): warm-cache ncu launch list (per-kernel durations as inside back-to-back replays) + the
# bench line of the configuration.  Outputs in gpurun_out/.
O=gpurun_out; mkdir -p $O; TAG=${1:-c3}
timeout 300 ncu --clock-control none --cache-control none --metrics gpu__time_duration.sum --launch-skip 40 -c 40 --csv \
    --log-file $O/${TAG}_launches_warm.csv python scripts/profile_steps.py 6 fullrank > $O/${TAG}_under_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --config c3 --steps 40 --warmup 5 --no-cpu-baseline --no-extras > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
cat $O/${TAG}_bench.json
