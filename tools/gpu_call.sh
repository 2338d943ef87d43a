#!/bin/bash
# GPU call 2: full parity suite (new back-end kernels), A/B against the round-1 kernels, bench line, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== pytest (new kernels)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== pytest subset with round-1 back end"; AISGPU_BE_V1=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "default" 2>&1 | tail -3
for m in 2; do
SWEEP_MODEL=$m timeout 600 python tools/fe_sweep.py \
  4,0,4096,AISGPU_BE_V1=1 \
  4,0,4096,AISGPU_BE_V1=0 \
  4,0,4096,AISGPU_BE_V1=0,AISGPU_BE_PIPE=1 \
  4,0,4096,AISGPU_BE_V1=0,AISGPU_BE_PIPE=1,AISGPU_DEC_RPW=3 \
  2>&1 | grep -v "^$" | tee -a gpurun_out/sweep2.jsonl
done
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; tail -5 gpurun_out/bench2.err; cat gpurun_out/bench2.json
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
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_default.csv python /tmp/one.py 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2a_launches_standard.csv python /tmp/one.py 0 > /dev/null 2>&1
python - <<'PY'
import csv, collections
for f in ("gpurun_out/r2a_launches_default.csv", "gpurun_out/r2a_launches_standard.csv"):
    try:
        rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    except Exception as e:
        print(f, e); continue
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
    d = collections.defaultdict(list)
    for r in rows[1:]:
        try: d[r[ik][:60]].append(float(r[iv].replace(",", "")))
        except: pass
    print(f)
    for k, v in d.items(): print("  %-60s n=%d last=%.1f us" % (k, len(v), v[-1] / 1000.0 if v[-1] > 5000 else v[-1]))
PY
