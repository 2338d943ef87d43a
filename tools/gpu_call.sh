#!/bin/bash
# GPU call 13: feeder tests, Default-chain priorities at both quoted shapes, full bench
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_feeder.py -m gpu -q 2>&1 | tail -5 | cut -c1-600
echo "== probe"; timeout 900 python tools/default_probe.py 2 - - AISGPU_PRIO=1 AISGPU_PRIO=1,AISGPU_BE_PIPE=1 AISGPU_BE_PIPE=1 2>&1 | tee gpurun_out/probe13.jsonl
PROBE_SHAPE=4096,65536,6000000,3 timeout 600 python tools/default_probe.py 2 - AISGPU_PRIO=1 AISGPU_PRIO=1,AISGPU_BE_PIPE=1 2>&1 | tee -a gpurun_out/probe13.jsonl
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench13.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_cu8']['value'], d['parity']['mismatches'])
for a in d['also']: print(a['workload'][:70], a['ms_per_step'], a.get('blocks_ms_per_step'), a['parity']['mismatches'])
PY
