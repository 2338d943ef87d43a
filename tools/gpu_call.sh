#!/bin/bash
# GPU call 3: failing cases verbose, fuzz debug, launch list of the reworked ModelDefault back end, chunk-length sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== fuzz debug"; timeout 300 python tools/fuzz_debug.py 0 8192 4 2>&1 | tail -40
timeout 300 python tools/fuzz_debug.py 2 8192 4 2>&1 | tail -25
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scale.py::test_decoder_fuzz_chunks --deselect tests/test_gpu_scale.py::test_decoder_fuzz_kernels > gpurun_out/pytest3.log 2>&1; tail -15 gpurun_out/pytest3.log
echo "== sweeps"
SWEEP_MODEL=2 timeout 600 python tools/fe_sweep.py 4,0,4096,AISGPU_BE_V1=0 4,0,4096,AISGPU_BE_V1=0,AISGPU_BE_PIPE=1 4,0,4096,AISGPU_BE_V1=0,AISGPU_BE_PIPE=1,AISGPU_DEC_RPW=3 2>&1 | grep -v "^$" | tee -a gpurun_out/sweep3.jsonl
SWEEP_MODEL=0 timeout 600 python tools/fe_sweep.py 4,0,4096,AISGPU_ST_G=16 4,0,4096,AISGPU_ST_G=32 4,0,4096,AISGPU_ST_G=64 4,0,4096,AISGPU_ST_G=64,AISGPU_DEC_RPW=3 2>&1 | grep -v "^$" | tee -a gpurun_out/sweep3.jsonl
echo "== ncu launch list (ModelDefault)"
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_default.csv python /tmp/one.py 2 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/r2b_launches_default.csv",):
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try: d[r[ik][:60]].append(float(r[iv].replace(",", "")))
        except: pass
    for k, v in d.items():
        if "aisgpu" in k: print("  %-60s n=%d last=%.1f us" % (k, len(v), v[-1] / 1000.0))
PY
