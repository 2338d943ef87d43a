#!/bin/bash
# GPU call r2b-14: ModelDefault under bench.py's timed region: rows per CTA of the fused kernel / rows per warp of the decoder (same box, alternating)
mkdir -p gpurun_out
for rep in 1 2; do
  for cfg in "" "AISGPU_CF_ROWS=8" "AISGPU_DEC_RPW=3" "AISGPU_CF_ROWS=8 AISGPU_DEC_RPW=3" ; do
    env $cfg timeout 600 python bench.py --model 2 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m2', '$cfg'.ljust(36), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4))
"
  done
done
for cfg in "" "AISGPU_DEC_RPW=3" "AISGPU_DEC_RPW=1"; do
    env $cfg timeout 600 python bench.py --model 0 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m0', '$cfg'.ljust(36), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4))
"
done
