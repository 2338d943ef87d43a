#!/bin/bash
# GPU call 10: new tests (resampler taps, tiny inputs), phase-search timing
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "resampled or dsk or tiny or default_1536k or default_small or default_phasesearch" > gpurun_out/pytest10.log 2>&1; tail -14 gpurun_out/pytest10.log | cut -c1-900
cat > /tmp/one.py <<'PY'
import os, sys
sys.path.insert(0, "ais-catcher_b200"); sys.path.insert(0, "tests")
import numpy as np, torch, aisgpu, aissynth
B, N, FS = 1024, 131072, 1536000
model = int(sys.argv[1])
u = np.stack([aissynth.random_stream(FS, N * 2, 1000 + i)[0] for i in range(8)])
ud = torch.from_numpy(u.view(np.float32)).cuda().view(8, 2, N, 2)
x = torch.empty((2, B, N, 2), dtype=torch.float32, device="cuda")
for b0 in range(0, B, 8):
    x[:, b0:b0 + 8] = ud.permute(1, 0, 2, 3)
x += torch.randn_like(x) * 0.005
eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=B, max_chunk=N, max_frames=1 << 20, host_staging=False)
for i in range(4):
    eng.submit_device(x[i % 2].data_ptr(), N, N)
    eng.sync()
print(len(eng.poll()))
PY
for m in 2 4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_m$m.csv python /tmp/one.py $m > /dev/null 2>&1
python - $m <<'PY'
import csv, collections, sys
f = "gpurun_out/r2e_launches_m%s.csv" % sys.argv[1]
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ik][:60]].append(float(r[iv].replace(",", "")))
    except: pass
for k, v in d.items():
    if "aisgpu" in k: print("  %-60s n=%d last=%.1f us" % (k, len(v), v[-1] / 1000.0))
PY
done
SWEEP_MODEL=2 timeout 600 python tools/fe_sweep.py 4,0,4096,AISGPU_BE_PIPE=1,AISGPU_DEC_RPW=3 4,0,4096 2>&1 | grep -v "^$" | tee -a gpurun_out/sweep10.jsonl
SWEEP_MODEL=4 timeout 600 python tools/fe_sweep.py 4,0,4096 2>&1 | grep -v "^$" | tee -a gpurun_out/sweep10.jsonl
