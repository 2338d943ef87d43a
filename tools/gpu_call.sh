#!/bin/bash
# GPU call r2b-20: front end with the chunk loop rolled (1136 instead of 1552 instructions in the kernel, +21 % executed): parity, A/B against the previous build
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "1536k or lane_split" > gpurun_out/pytest_fe.log 2>&1; tail -2 gpurun_out/pytest_fe.log | cut -c1-300
for rep in 1 2 3; do
  for lib in "new:" "prev:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_prev.so"; do
    name=${lib%%:*}; envs=${lib#*:}
    env $envs timeout 600 python bench.py --model 0 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m0', '$name'.ljust(6), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4), 'iso', round(d['roofline']['isolated_ms_per_launch'],4))
"
  done
done
