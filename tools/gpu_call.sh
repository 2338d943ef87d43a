#!/bin/bash
# Final validation of a build: full GPU suite, smoke, bench.py (own arm + reference arm), ncu launch list of bench.py's own step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; tail -3 gpurun_out/pytest_all.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -1 gpurun_out/final_bench.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','ms_per_step','spread','clocks','gpu_launches']}); print(d['e2e']['value'], d['e2e_cu8']['value'], d['parity']['mismatches'], d['cpu_baseline']['value'])
print(d['roofline']['frac'], d['roofline']['isolated_frac'], d['roofline']['whole_chain_frac'])
for a in d.get('also',[]): print(a['workload'][:60], a['ms_per_step'], a.get('whole_chain_frac'), a['parity']['mismatches'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2>/dev/null; cut -c1-160 gpurun_out/final_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 200 --csv --log-file gpurun_out/final_launches_bench.csv python bench.py --steps 2 --warmup 3 --blocks 1 --no-also --no-cpu --no-parity --e2e-steps 1 > /dev/null 2>&1; grep -c k_ gpurun_out/final_launches_bench.csv
