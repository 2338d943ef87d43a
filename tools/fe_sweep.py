#!/usr/bin/env python
"""GPU tuning helper: front-end launch-shape sweep (AISGPU_FE_WARPS / _TILE / _CTAS) on the bench workload.
Prints isolated front-end time (sync after every submit) and pipelined step time per setting."""
import os
import sys
import json

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ais-catcher_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import aisgpu
import aissynth

FS, N, B, R = 1536000, 131072, int(os.environ.get("SWEEP_B", "1024")), 4
model = int(os.environ.get("SWEEP_MODEL", "0"))
dev = torch.device("cuda", 0)
uniq = np.stack([aissynth.random_stream(FS, N * R, 1000 + u)[0] for u in range(16)])
ud = torch.from_numpy(uniq.view(np.float32)).to(dev).view(16, N * R, 2)
x = torch.empty((R, B, N, 2), dtype=torch.float32, device=dev)
for b0 in range(0, B, 16):
    nb = min(16, B - b0)
    x[:, b0:b0 + nb] = ud[:nb].view(nb, R, N, 2).permute(1, 0, 2, 3)
torch.manual_seed(7)
x += torch.randn_like(x) * 0.005
torch.cuda.synchronize()

torch.manual_seed(7)
settings = [s.split(",") for s in (sys.argv[1:] or ["4,0,4096"])]
for st in settings:
    w, tile, ctas = st[:3]
    os.environ["AISGPU_FE_WARPS"], os.environ["AISGPU_FE_TILE"], os.environ["AISGPU_FE_CTAS"] = w, tile, ctas
    extra = {}
    for kv in st[3:]:  # further KEY=VALUE environment settings, e.g. AISGPU_DEC_RPW=3
        k_, v_ = kv.split("=")
        os.environ[k_] = v_
        extra[k_] = v_
    eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=B, max_chunk=N, max_frames=1 << 20)
    for i in range(3):
        eng.submit_device(x[i % R].data_ptr(), N, N)
        eng.sync()
    iso = []
    for i in range(6):
        eng.submit_device(x[i % R].data_ptr(), N, N)
        eng.sync()
        iso.append(eng.last_frontend_ms())
    eng.poll()
    est = torch.cuda.ExternalStream(eng.cuda_stream(), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 20
    e0.record(est)
    for i in range(K):
        eng.submit_device(x[i % R].data_ptr(), N, N)
    eng.join()
    e1.record(est)
    e1.synchronize()
    step = e0.elapsed_time(e1) / K
    nm = len(eng.poll())
    print(json.dumps({"extra": extra, "warps": int(w), "tile": int(tile), "ctas": int(ctas), "model": model, "B": B, "fe_iso_ms": round(min(iso), 4),
                      "fe_iso_GBs": round(B * N * 8 / min(iso) / 1e6, 1), "step_ms": round(step, 4), "step_GBs": round(B * N * 8 / step / 1e6, 1), "msgs": nm}), flush=True)
    eng.close()
