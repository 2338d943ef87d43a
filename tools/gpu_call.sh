#!/bin/bash
mkdir -p gpurun_out
echo "== failing fuzz test"; timeout 600 python -m pytest "tests/test_gpu_scale.py::test_decoder_fuzz_kernels" -m gpu -q -x 2>&1 | grep -E "AssertionError|passed|failed" | head -5 | cut -c1-700
echo "== seeds 32..159"
FUZZ_SEED0=32 FUZZ_MODES=back_to_back,sync_every_submit timeout 600 python tools/fuzz_debug2.py 0 8192 128 2>&1 | tail -30 | cut -c1-400
echo "== ncu full captures"
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
for i in range(3):
    eng.submit_device(x[i % 2].data_ptr(), N, N)
    eng.sync()
print(len(eng.poll()))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_phase_search_ema4 -s 2 -c 1 -o gpurun_out/r2c_phase_search python /tmp/one.py 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cgf_fused -s 2 -c 1 -o gpurun_out/r2c_cgf_fused python /tmp/one.py 2 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_v2_engine -s 2 -c 1 -o gpurun_out/r2c_v2_engine python /tmp/one.py 11 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_frontend_st -s 2 -c 1 -o gpurun_out/r2c_frontend_st python /tmp/one.py 0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
