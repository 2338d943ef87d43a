#!/bin/bash
# BASELINE configs[3] shape on one GPU: batch 8192 CF32 streams @1536k with injected bursts, NMEA + order checked against the reference on 32 sampled streams
mkdir -p gpurun_out
timeout 1200 python bench.py --batch 8192 --no-also --no-cpu --blocks 5 --steps 10 --e2e-steps 2 > gpurun_out/r2j_bench_b8192.json 2> gpurun_out/r2j_bench_b8192.err; tail -2 gpurun_out/r2j_bench_b8192.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_bench_b8192.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','ms_per_step','spread','msgs_per_s']}); print(d['parity']); print(d['roofline']['frac'], d['roofline']['whole_chain_frac'], d['e2e']['value'])
PY
timeout 1200 python bench.py --batch 8192 --model 2 --no-also --no-cpu --blocks 5 --steps 10 --e2e-steps 2 > gpurun_out/r2j_bench_b8192_m2.json 2> gpurun_out/r2j_bench_b8192_m2.err; tail -1 gpurun_out/r2j_bench_b8192_m2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_bench_b8192_m2.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','ms_per_step','spread','msgs_per_s']}); print(d['parity']); print(d['roofline']['frac'], d['roofline']['whole_chain_frac'])
PY
