#!/bin/bash
# GPU call 14: whole GPU suite with the coherent chain on two back-end streams, stability of its step time, full bench
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest14.log 2>&1; tail -6 gpurun_out/pytest14.log | cut -c1-600
echo "== probe"; timeout 900 python tools/default_probe.py 2 - - - - 2>&1 | tee gpurun_out/probe14.jsonl
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench14.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_cu8']['value'], d['parity']['mismatches'])
for a in d['also']: print(a['workload'][:70], a['ms_per_step'], a.get('blocks_ms_per_step'), a['parity']['mismatches'])
PY
