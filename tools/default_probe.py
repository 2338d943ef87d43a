"""Per-block step times of one model under the bench's conditions (1024 x 131072 CF32 @1536k resident in HBM, 20-step blocks,
no polling in between), for a list of env settings: python tools/default_probe.py MODEL "K=V,K=V" "K=V" ...  ("-" = defaults)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 2 and sys.argv[1] != "--child":
    model = sys.argv[1]
    for spec in sys.argv[2:]:
        env = dict(os.environ)
        if spec != "-":
            for kv in spec.split(","):
                k, v = kv.split("=")
                env[k] = v
        out = subprocess.run([sys.executable, __file__, "--child", model], env=env, capture_output=True, text=True)
        print(json.dumps({"env": spec, "model": int(model), **json.loads(out.stdout.strip().splitlines()[-1])}) if out.returncode == 0 else out.stderr[-600:], flush=True)
    sys.exit(0)
sys.path.insert(0, os.path.join(ROOT, "ais-catcher_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, aisgpu, aissynth
model = int(sys.argv[2])
B, N, FS, R = [int(v) for v in os.environ.get("PROBE_SHAPE", "1024,131072,1536000,4").split(",")]
u = np.stack([aissynth.random_stream(FS, N * R, 1000 + i)[0] for i in range(8)])
ud = torch.from_numpy(u.view(np.float32)).cuda().view(8, R, N, 2)
x = torch.empty((R, B, N, 2), dtype=torch.float32, device="cuda")
for b0 in range(0, B, 8):
    x[:, b0:b0 + 8] = ud.permute(1, 0, 2, 3)
x += torch.randn_like(x) * 0.005
eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=B, max_chunk=N, max_frames=1 << 20)
est = torch.cuda.ExternalStream(eng.cuda_stream())
steps, blocks = 20, 6
i = 0
for _ in range(3):
    eng.submit_device(x[i % R].data_ptr(), N, N); i += 1
eng.sync()
evs = [torch.cuda.Event(enable_timing=True) for _ in range(blocks + 1)]
eng.join(); evs[0].record(est)
for b in range(blocks):
    for _ in range(steps):
        eng.submit_device(x[i % R].data_ptr(), N, N); i += 1
    eng.join(); evs[b + 1].record(est)
evs[-1].synchronize()
ms = [round(evs[b].elapsed_time(evs[b + 1]) / steps, 4) for b in range(blocks)]
fe = eng.frontend_times(64)
print(json.dumps({"blocks_ms_per_step": ms, "fe_ms": round(sum(fe) / len(fe), 4), "frames": eng.poll_upto_count(-1)[0]}))
