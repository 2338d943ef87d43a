#!/bin/bash
# GPU call 18: lean atan2 path in the FM kernel + Challenger on two back-end streams: whole suite, probes, bench
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest18.log 2>&1; tail -5 gpurun_out/pytest18.log | cut -c1-500
echo "== probe"; timeout 900 python tools/default_probe.py 4 - - 2>&1 | tee gpurun_out/probe18.jsonl
timeout 900 python tools/default_probe.py 0 - - 2>&1 | tee -a gpurun_out/probe18.jsonl
timeout 900 python tools/default_probe.py 2 - 2>&1 | tee -a gpurun_out/probe18.jsonl
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench18.json 2> gpurun_out/bench18.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench18.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_cu8']['value'], d['parity']['mismatches'], d['roofline']['frac'], d['roofline']['isolated_frac'])
for a in d['also']: print(a['workload'][:70], round(a['ms_per_step'],4), a.get('blocks_ms_per_step'), a['parity']['mismatches'], a['parity']['msgs_checked'])
PY
tail -3 gpurun_out/bench18.err
