#!/bin/bash
# GPU call r2b-17: warm-up prefetch as a rolled loop (code size at K = 6, 7): parity, high-rate sweep, configs[2] launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_matrix.py -m gpu -x -q > gpurun_out/pytest_fe.log 2>&1; tail -3 gpurun_out/pytest_fe.log | cut -c1-600
timeout 900 python tools/rate_sweep.py 12288000:512:1048576:0:1 6144000:1024:524288:0:1 3072000:1024:262144:0:1 1536000:1024:131072:0:1 768000:1024:65536:0:1 > gpurun_out/r2j_rate_sweep_hi.jsonl 2>/dev/null; cut -c1-260 gpurun_out/r2j_rate_sweep_hi.jsonl
AISGPU_ST_NB=5 AISGPU_ST_L=32 timeout 900 python tools/rate_sweep.py 12288000:512:1048576:0:1 6144000:1024:524288:0:1 > gpurun_out/r2j_rate_sweep_hi_old.jsonl 2>/dev/null; cut -c1-260 gpurun_out/r2j_rate_sweep_hi_old.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 120 --csv --log-file gpurun_out/r2j_launches_c2.csv python tools/ncu_run.py 2 4096 65536 6000000 > /dev/null 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2j_launches_c2.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
for k,v in d.items(): print('  %-72s n=%2d last=%9.1f us' % (k,len(v),v[-1]/1000))
PY
