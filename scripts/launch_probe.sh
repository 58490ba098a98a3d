#!/bin/bash
# Is the fused loop paced by the host's cudaGraphLaunch calls or by the device?  (AVI_DEBUG_LAUNCH prints both)
for u in 1 4 16; do
  echo "== AVI_GRAPH_UNROLL=$u"
  AVI_GRAPH_UNROLL=$u AVI_DEBUG_LAUNCH=1 timeout 200 python bench.py --steps 320 --warmup 32 --no-cpu-baseline 2> gpurun_out/probe_err.txt | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cold', round(d['value']), 'warm', round(d['value_l2_resident']), 'e2e', round(d['e2e']['value']), d['gpu_launches'])"
  grep avi_opt_steps gpurun_out/probe_err.txt | tail -4
done
