#!/bin/bash
# GPU call 15: ticket-order fix (fuzz chunks), PhaseSearchEMA with 8 hypotheses per lane: parity + timing
mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests/test_gpu_scale.py -m gpu -q -k "fuzz_chunks or back_to_back or poll_upto or overflow" 2>&1 | tail -4 | cut -c1-400
echo "== ps8 parity"; AISGPU_PS_LANES=2 timeout 1800 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -k "default or challenger or bench_shape or fuzz_kernels" 2>&1 | tail -4 | cut -c1-400
echo "== probe"; timeout 900 python tools/default_probe.py 2 - AISGPU_PS_LANES=2 - AISGPU_PS_LANES=2 2>&1 | tee gpurun_out/probe15.jsonl
cat > /tmp/one.py <<'PY'
import os, sys
sys.path.insert(0, "ais-catcher_b200"); sys.path.insert(0, "tests")
import numpy as np, torch, aisgpu, aissynth
B, N, FS = 1024, 131072, 1536000
u = np.stack([aissynth.random_stream(FS, N * 2, 1000 + i)[0] for i in range(8)])
ud = torch.from_numpy(u.view(np.float32)).cuda().view(8, 2, N, 2)
x = torch.empty((2, B, N, 2), dtype=torch.float32, device="cuda")
for b0 in range(0, B, 8):
    x[:, b0:b0 + 8] = ud.permute(1, 0, 2, 3)
x += torch.randn_like(x) * 0.005
eng = aisgpu.Engine(model=2, sample_rate=FS, n_streams=B, max_chunk=N, max_frames=1 << 20, host_staging=False)
for i in range(4):
    eng.submit_device(x[i % 2].data_ptr(), N, N)
    eng.sync()
print(len(eng.poll()))
PY
for L in 4 2; do
AISGPU_PS_LANES=$L timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_ps$L.csv python /tmp/one.py > /dev/null 2>&1
python - $L <<'PY'
import csv, collections, sys
f = "gpurun_out/r2f_launches_ps%s.csv" % sys.argv[1]
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
d = collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ik][:60]].append(float(r[iv].replace(",", "")))
    except: pass
for k, v in d.items():
    if "phase_search" in k: print("  %-60s n=%d last=%.1f us" % (k, len(v), v[-1] / 1000.0))
PY
done
