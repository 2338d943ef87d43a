#!/bin/bash
# GPU call r2b-19: front-end register footprint live: prev (164 regs, droop filter branched) / a (160, selected) / b (152 via __maxnreg__, 3 spilled pairs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "1536k or lane_split or cu8" > gpurun_out/pytest_fe.log 2>&1; tail -2 gpurun_out/pytest_fe.log | cut -c1-300
for rep in 1 2 3; do
  for lib in "b:" "a:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_a.so" "prev:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_prev.so"; do
    name=${lib%%:*}; envs=${lib#*:}
    env $envs timeout 600 python bench.py --model 0 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m0', '$name'.ljust(6), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4), 'iso', round(d['roofline']['isolated_ms_per_launch'],4))
"
  done
done
for lib in "b:" "a:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_a.so" "prev:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_prev.so"; do
    name=${lib%%:*}; envs=${lib#*:}
    env $envs timeout 600 python bench.py --model 2 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m2', '$name'.ljust(6), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4))
"
done
