#!/usr/bin/env python
"""GPU debugging helper: one decoder-fuzz configuration (tests/test_gpu_scale.py) with every tap compared, and the first
differing messages of the first differing stream printed side by side."""
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
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
FS, N = 96000, 262144
xs = np.stack([S.fuzz_stream(FS, N, s)[0] for s in range(B)])
eng = aisgpu.Engine(model=model, sample_rate=FS, n_streams=B, max_chunk=chunk, taps=True, max_frames=1 << 16)
refs = [O.RefModel(model=model, sample_rate=FS, taps=True) for _ in range(B)]
got = [[] for _ in range(B)]
want = [[] for _ in range(B)]
tap_bad = 0
for c in range(N // chunk):
    eng.submit(np.ascontiguousarray(xs[:, c * chunk:(c + 1) * chunk]), chunk)
    for s in range(B):
        refs[s].push(xs[s, c * chunk:(c + 1) * chunk])
        for ch in range(2):
            w = refs[s].tap_c(O.TAP_CA + ch)
            g = eng.tap(aisgpu.TAP_C, s, ch)
            if len(w) != len(g) or not np.array_equal(w.view(np.uint32), g.view(np.uint32)):
                tap_bad += 1
                if tap_bad < 4:
                    print("C tap differs: chunk", c, "stream", s, "ch", ch, len(g), len(w))
            nph = 5
            for ph in range(nph):
                w = refs[s].tap_f(ch * 5 + ph)
                g = eng.tap(aisgpu.TAP_DEC, s, ch + 2 * ph, dtype=np.float32)
                if len(w) != len(g) or not np.array_equal((w > 0), (g > 0)):
                    tap_bad += 1
                    if tap_bad < 8:
                        print("DEC tap differs: chunk", c, "stream", s, "ch", ch, "phase", ph, len(g), len(w))
    for m in eng.poll():
        got[m.stream].append((m.channel, m.nbits, m.start_idx, m.end_idx, m.chunk, m.nmea))
    for s in range(B):
        for m in refs[s].messages():
            want[s].append((m.channel, m.nbits, m.start_idx, m.end_idx, c, m.nmea))
print("tap mismatches:", tap_bad, "counters", eng.counters())
nbad = 0
for s in range(B):
    g = [(a[0], a[1], a[2], a[3], tuple(a[5])) for a in got[s]]
    w = [(a[0], a[1], a[2], a[3], tuple(a[5])) for a in want[s]]
    if g != w:
        nbad += 1
        print("stream", s, "got", len(g), "want", len(w))
        gs, ws = set(g), set(w)
        for a in got[s]:
            if (a[0], a[1], a[2], a[3], tuple(a[5])) not in ws:
                print("   only GPU :", a)
        for a in want[s]:
            if (a[0], a[1], a[2], a[3], tuple(a[5])) not in gs:
                print("   only REF :", a)
        if gs == ws:
            print("   same set, different order; first 12 of each:")
            for a, b in list(zip(got[s], want[s]))[:12]:
                print("     ", a[:5], "|", b[:5])
print("streams differing:", nbad, "of", B)
