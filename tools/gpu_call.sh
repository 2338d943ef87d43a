#!/bin/bash
# GPU call 17: ncu evidence for the final kernels (one --set full capture per kernel) + launch lists per model
mkdir -p gpurun_out
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
cap() { # name model kernel-regex
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$3" --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2g_$1 python /tmp/one.py $2 > /dev/null 2>&1
  ls -la gpurun_out/r2g_$1.ncu-rep 2>/dev/null | awk '{print $5, $9}'
}
cap frontend_st 0 k_frontend_st
cap fm_fir5 0 k_fm_fir5
cap decode3_fm 0 k_decode3
cap cgf_estimate 2 k_cgf_estimate
cap cgf_fused 2 k_cgf_fused
cap phase_search 2 k_phase_search_ema4
cap decode3_coh 2 k_decode3
cap decode10 4 k_decode10
for m in 0 2 4 11; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_m$m.csv python /tmp/one.py $m > /dev/null 2>&1
python - $m <<'PY'
import csv, collections, sys
f = "gpurun_out/r2g_launches_m%s.csv" % sys.argv[1]
rows = [r for r in csv.reader(open(f)) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
d = collections.OrderedDict()
for r in rows[1:]:
    try: d.setdefault(r[ik][:70], []).append(float(r[iv].replace(",", "")))
    except: pass
print("model", sys.argv[1])
for k, v in d.items():
    if "aisgpu" in k or "k_" in k: print("  %-70s n=%d last=%.1f us" % (k, len(v), v[-1] / 1000.0))
PY
done
