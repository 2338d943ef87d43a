#!/bin/bash
# phase search with a static trip structure for whole tiles: parity, instructions + time alone (ncu), live A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "default or challenger or config2 or bench_shape or back_to_back" > gpurun_out/pytest_ps.log 2>&1; tail -2 gpurun_out/pytest_ps.log | cut -c1-300
for lib in "new:" "prev:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_prev.so"; do
  name=${lib%%:*}; envs=${lib#*:}
  env $envs timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_phase -c 4 --csv --log-file gpurun_out/ps_$name.csv python tools/ncu_run.py 2 > /dev/null 2>&1; echo $name; grep "k_phase" gpurun_out/ps_$name.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -4
done
for rep in 1 2; do
  for lib in "new:" "prev:AISGPU_LIB=/root/repo/ais-catcher_b200/libaisgpu_prev.so"; do
    name=${lib%%:*}; envs=${lib#*:}
    env $envs timeout 600 python bench.py --model 2 --no-also --no-cpu --no-parity --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('m2', '$name'.ljust(6), round(d['ms_per_step'],4), round(d['spread']['min_ms_per_step'],4), round(d['spread']['max_ms_per_step'],4), 'fe_live', round(d['roofline']['frontend_ms_per_launch'],4))
"
  done
done
