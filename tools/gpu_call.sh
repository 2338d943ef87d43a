#!/bin/bash
# GPU call r2b-15: phase search on its own most-urgent stream (parity of the pipelined paths, then live A/B), decoder rows per warp for the FM chain
mkdir -p gpurun_out
AISGPU_PS_STREAM=1 timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -x -q -k "default or challenger or config2 or bench_shape or back_to_back or poll_upto or fuzz_chunks" > gpurun_out/pytest_psstream.log 2>&1; tail -3 gpurun_out/pytest_psstream.log | cut -c1-600
for rep in 1 2; do
  for cfg in "" "AISGPU_PS_STREAM=1" ; do
    env $cfg timeout 600 python bench.py --model 2 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m2', '$cfg'.ljust(24), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4))
"
  done
  for cfg in "" "AISGPU_DEC_RPW=3" ; do
    env $cfg timeout 600 python bench.py --model 0 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m0', '$cfg'.ljust(24), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4))
"
  done
done
AISGPU_PS_STREAM=1 timeout 600 python tools/default_probe.py 4 - > gpurun_out/probe15_m4.jsonl 2>&1; cat gpurun_out/probe15_m4.jsonl | cut -c1-300
