#!/bin/bash
# GPU call r2b-16: final build -- full GPU suite, smoke, bench.py both arms, rate sweep, launch lists of the shipped shapes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; tail -3 gpurun_out/pytest_all.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -2 gpurun_out/r2j_bench.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_bench.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','ms_per_step','spread','e2e','e2e_cu8','clocks','gpu_launches']})
print(d['parity']); print(d['roofline']); print(d['cpu_baseline'])
for a in d.get('also',[]): print(a['workload'][:70], a['ms_per_step'], a.get('whole_chain_frac'), a['parity']['mismatches'], a.get('blocks_ms_per_step'))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r2j_bench_reference_arm.json
timeout 1200 python tools/rate_sweep.py 288000:1024:24576:0:1 384000:1024:32768:0:1 768000:1024:65536:0:1 1536000:1024:131072:0:1 3072000:1024:262144:0:1 6144000:1024:524288:0:1 12288000:512:1048576:0:1 1536000:8192:65536:0:1 1536000:8192:65536:2:1 6144000:1024:524288:2:1 6000000:4096:65536:0:1 > gpurun_out/r2j_rate_sweep.jsonl 2> gpurun_out/rate_sweep.err; cut -c1-260 gpurun_out/r2j_rate_sweep.jsonl
for m in 0 2 4; do timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 80 --csv --log-file gpurun_out/r2j_launches_m$m.csv python tools/ncu_run.py $m > /dev/null 2>&1; done
ls -la gpurun_out/r2j_launches_m*.csv
