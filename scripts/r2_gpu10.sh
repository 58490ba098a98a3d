#!/bin/bash
O=gpurun_out; mkdir -p $O
for rows in 10000 1250; do
  timeout 90 python scripts/step_prof.py $rows 12 cold > $O/g10_prof_cold_$rows.txt 2>&1; grep -vE "^ ?(0|1|3|4|6|9|12|14|17|18) " $O/g10_prof_cold_$rows.txt
  timeout 60 python scripts/step_prof.py $rows > $O/g10_prof_warm_$rows.txt 2>&1; grep -vE "^ ?(0|1|3|4|6|9|12|14|17|18) " $O/g10_prof_warm_$rows.txt
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1"
timeout 300 python bench.py --config c5 --steps 100 --warmup 10 --no-extras 2>&1 | cut -c1-400
