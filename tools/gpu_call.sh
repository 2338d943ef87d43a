#!/bin/bash
# GPU call r2b-12: A/B of the front-end shape under bench.py's own timed region (same box, alternating)
mkdir -p gpurun_out
for rep in 1 2 3; do
  for cfg in "" "AISGPU_ST_NB=5 AISGPU_ST_L=32" "AISGPU_ST_L=16" ; do
    env $cfg timeout 600 python bench.py --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$cfg'.ljust(32), round(d['ms_per_step'],4), d['spread']['min_ms_per_step'], d['spread']['max_ms_per_step'], 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4), 'iso', round(d['roofline']['isolated_ms_per_launch'],4))
"
  done
done
