#!/bin/bash
# GPU call 19 (final single-GPU run): whole suite, smoke, bench + reference arm, rate sweep, sanitizer
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest19.log 2>&1; tail -4 gpurun_out/pytest19.log | cut -c1-500
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 1500 python bench.py > gpurun_out/bench19.json 2> gpurun_out/bench19.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench19.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e_cu8']['value'], d['parity']['mismatches'], d['roofline']['frac'], d['roofline']['isolated_frac'], d['clocks'])
for a in d['also']: print(a['workload'][:70], round(a['ms_per_step'],4), a.get('blocks_ms_per_step'), a['parity']['mismatches'], a['parity']['msgs_checked'])
PY
echo "== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench19_ref.json 2>/dev/null; cut -c1-300 gpurun_out/bench19_ref.json
echo "== rate sweep"; timeout 1500 python tools/rate_sweep.py 288000:1024:24576:0:1 384000:1024:32768:0:1 768000:1024:65536:0:1 1536000:1024:131072:0:1 3072000:1024:262144:0:1 6144000:1024:524288:0:1 12288000:512:1048576:0:1 1536000:8192:65536:0:1 1536000:8192:65536:2:1 6144000:1024:524288:2:1 6000000:4096:65536:0:1 2>&1 | grep "^{" | tee gpurun_out/r2h_rate_sweep.jsonl | cut -c1-200
echo "== sanitizer"; timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r2h_sanitizer_memcheck.log python tools/sanitizer_workload.py > gpurun_out/r2h_sanitizer_memcheck.out 2>&1; tail -3 gpurun_out/r2h_sanitizer_memcheck.out; tail -2 gpurun_out/r2h_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --log-file gpurun_out/r2h_sanitizer_racecheck.log python tools/sanitizer_workload.py > gpurun_out/r2h_sanitizer_racecheck.out 2>&1; tail -1 gpurun_out/r2h_sanitizer_racecheck.out; tail -2 gpurun_out/r2h_sanitizer_racecheck.log
