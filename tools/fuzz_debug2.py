#!/usr/bin/env python
"""GPU debugging helper: the decoder-fuzz configuration WITHOUT taps (submits overlap), bisecting over synchronisation
modes; prints which streams differ and the first differing messages."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("ais-catcher_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import aisgpu
import aissynth as S
import oracle as O

model = int(sys.argv[1]) if len(sys.argv) > 1 else 0
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
B = int(sys.argv[3]) if len(sys.argv) > 3 else 32
FS, N = 96000, 262144
SEED0 = int(os.environ.get("FUZZ_SEED0", "0"))
xs = np.stack([S.fuzz_stream(FS, N, SEED0 + s)[0] for s in range(B)])
want = {}
for s in range(B):
    r = O.RefModel(model=model, sample_rate=FS)
    r.run(xs[s], chunk)
    want[s] = [(m.channel, m.nbits, m.start_idx, m.end_idx, tuple(m.nmea)) for m in r.messages()]
for mode in os.environ.get("FUZZ_MODES", "sync_every_submit,poll_every_submit,back_to_back,back_to_back_pipe0,back_to_back_B1").split(","):
    if mode == "back_to_back_pipe0":
        os.environ["AISGPU_BE_PIPE"] = "0"
    else:
        os.environ.pop("AISGPU_BE_PIPE", None)
    nb = 1 if mode == "back_to_back_B1" else B
    bad_total = 0
    for s0 in range(0, B if nb == 1 else 1):
        sel = [s0] if nb == 1 else list(range(B))
        eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=len(sel), max_chunk=chunk, max_frames=1 << 16)
        got = {s: [] for s in sel}
        def take(msgs):
            for m in msgs:
                got[sel[m.stream]].append((m.channel, m.nbits, m.start_idx, m.end_idx, tuple(m.nmea)))
        for c in range(N // chunk):
            eng.submit(np.ascontiguousarray(xs[sel][:, c * chunk:(c + 1) * chunk]), chunk)
            if mode == "sync_every_submit":
                eng.sync()
            elif mode == "poll_every_submit":
                take(eng.poll())
        take(eng.poll())
        bad = [s for s in sel if got[s] != want[s]]
        bad_total += len(bad)
        if bad and bad_total <= 3:
            s = bad[0]
            gs, ws = set(got[s]), set(want[s])
            print("  mode", mode, "stream", s, "got", len(got[s]), "want", len(want[s]))
            for a in got[s]:
                if a not in ws:
                    print("     only GPU:", a[:4], a[4][0][:40])
            for a in want[s]:
                if a not in gs:
                    print("     only REF:", a[:4], a[4][0][:40])
            if gs == ws:
                k = next(i for i, (a, b) in enumerate(zip(got[s], want[s])) if a != b)
                print("     same set, order differs from index", k, ":", [a[:4] for a in got[s][k:k + 4]], "|", [a[:4] for a in want[s][k:k + 4]])
        eng.close()
    print("mode %-22s model %d chunk %d: %d of %d streams differ" % (mode, model, chunk, bad_total, B))
