"""Small run of every model and entry point for compute-sanitizer (memcheck / racecheck):
compute-sanitizer --tool memcheck python tools/sanitizer_workload.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "ais-catcher_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import aisgpu, aissynth
import __graft_entry__ as g

g.smoke()
fs, N, B = 1536000, 16384, 3
xs = np.stack([aissynth.random_stream(fs, N * 4, 300 + s)[0] for s in range(B)])
for model in (aisgpu.MODEL_STANDARD, aisgpu.MODEL_BASE, aisgpu.MODEL_DEFAULT, aisgpu.MODEL_CHALLENGER, aisgpu.MODEL_V2):
    for taps in (False, True):  # without taps the back end runs over two streams
        eng = aisgpu.Engine(model=model, sample_rate=fs, n_streams=B, max_chunk=N, taps=taps)
        n = 0
        for c in range(4):
            eng.submit(np.ascontiguousarray(xs[:, c * N:(c + 1) * N]), N)
        n += len(eng.poll())
        eng.close()
        print("model", model, "taps", taps, "messages", n)
for fmt, nodd in ((aisgpu.FMT_CF32, 64 * 257), (aisgpu.FMT_CU8, 64 * 131)):  # block lengths with no power-of-two lane split: uneven sub-segments, spare lanes
    xo = np.stack([aissynth.random_stream(fs, nodd * 3, 350 + s)[0] for s in range(5)])
    eng = aisgpu.Engine(model=aisgpu.MODEL_DEFAULT, sample_rate=fs, fmt=fmt, n_streams=5, max_chunk=nodd)
    for c in range(3):
        blk = np.ascontiguousarray(xo[:, c * nodd:(c + 1) * nodd])
        eng.submit(np.stack([aissynth.to_cu8(r) for r in blk]) if fmt == aisgpu.FMT_CU8 else blk, nodd)
    print("odd block", nodd, "fmt", fmt, "messages", len(eng.poll()))
    eng.close()
eng = aisgpu.Engine(model=aisgpu.MODEL_DEFAULT, sample_rate=6000000, n_streams=2, max_chunk=32768)  # resampler pre-stage
x6 = np.stack([aissynth.random_stream(6000000, 32768 * 3, 400 + s)[0] for s in range(2)])
for c in range(3):
    eng.submit(np.ascontiguousarray(x6[:, c * 32768:(c + 1) * 32768]), 32768)
print("6 MSPS messages", len(eng.poll()))
eng.close()
with tempfile.TemporaryDirectory() as d:  # file feeder, CU8, ragged lengths, FP_DS integer front end
    paths = []
    for s in range(2):
        p = os.path.join(d, "r%d.cu8" % s)
        aissynth.to_cu8(xs[s][:N * 3 - 100 * s]).tofile(p)
        paths.append(p)
    eng = aisgpu.Engine(model=aisgpu.MODEL_STANDARD, sample_rate=fs, fmt=aisgpu.FMT_CU8, n_streams=2, max_chunk=N, fp_ds=True)
    msgs, nb = eng.feed_files(paths, N)
    print("feeder blocks", nb, "messages", len(msgs))
    eng.close()
print("sanitizer workload done")
