import os, sys, json
os.environ["AISGPU_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ais-catcher_b200"))
import numpy as np, torch, aisgpu, bench
FS, N, B, R = 1536000, 131072, 1024, 4
uniq = bench.make_unique_streams(32, N * R)
dev = torch.device("cuda", 0)
ud = torch.view_as_complex(torch.from_numpy(uniq.view(np.float32)).to(dev).view(32, N * R, 2))
x = torch.empty((R, B, N), dtype=torch.complex64, device=dev)
for b0 in range(0, B, 32): x[:, b0:b0+32, :] = ud.view(32, R, N).permute(1, 0, 2)
noise = torch.empty((B, N), dtype=torch.complex64, device=dev)
for r in range(R):
    torch.view_as_real(noise).normal_(0.0, 0.005); x[r] += noise
for model in (0, 2):
    eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=B, max_chunk=N)
    for i in range(R): eng.submit_device(x[i].data_ptr(), N, N)
    eng.sync()
    d = eng.tap(6, 0, 0, dtype=np.int64).reshape(-1, 4)
    cyc, slow, ncrc, bits = d[:,0], d[:,1], d[:,2], d[:,3]
    print("model", model, "rows", len(d), "cycles mean %.0f p50 %.0f p90 %.0f max %.0f" % (cyc.mean(), np.median(cyc), np.percentile(cyc, 90), cyc.max()))
    print("  slow steps mean %.0f max %d; crc runs mean %.1f max %d; crc bits mean %.0f max %d" % (slow.mean(), slow.max(), ncrc.mean(), ncrc.max(), bits.mean(), bits.max()))
    A = np.stack([np.ones(len(d)), slow, bits], 1)
    coef, *_ = np.linalg.lstsq(A, cyc.astype(float), rcond=None)
    print("  fit: cycles ~ %.0f + %.1f*slow_steps + %.1f*crc_bits ; fast step ~ %.1f cycles" % (coef[0], coef[1], coef[2], coef[0]/820))
    eng.poll(); eng.close()
