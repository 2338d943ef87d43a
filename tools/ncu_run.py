"""Workload for ncu captures: MODEL at the bench shape (1024 x 131072 CF32 @1536k resident in HBM), 3 warm-up submits, then 2 submits
with a sync after each (so a launch list shows every kernel alone).  python tools/ncu_run.py MODEL [B N FS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ais-catcher_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, aisgpu, aissynth
model = int(sys.argv[1])
B, N, FS = (int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1024, 131072, 1536000)
R = 2
u = np.stack([aissynth.random_stream(FS, N * R, 1000 + i)[0] for i in range(8)])
ud = torch.from_numpy(u.view(np.float32)).cuda().view(8, R, N, 2)
x = torch.empty((R, B, N, 2), dtype=torch.float32, device="cuda")
for b0 in range(0, B, 8):
    x[:, b0:b0 + 8] = ud.permute(1, 0, 2, 3)
x += torch.randn_like(x) * 0.005
torch.cuda.synchronize()
eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=B, max_chunk=N, max_frames=1 << 20)
for i in range(5):
    eng.submit_device(x[i % R].data_ptr(), N, N)
    eng.sync()
print("frames", eng.poll_upto_count(-1)[0])
