#!/usr/bin/env python
"""Generates tests/golden/golden.json + tests/golden/burst_96k.cu8 from the UNMODIFIED reference
(oracle/_ref/libaisref.so, compiled from /root/reference by oracle/Makefile with strict IEEE flags).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there (and the CPU suite when _ref is absent) check against these files.
Each case records: sha256 of the input bytes (guards the seeded generator), every message the reference emitted
(NMEA, payload, 48 kHz start/end sample counters, level/ppm as float bit patterns) per chunk, and sha256 of each
intermediate tap accumulated over all chunks.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "ais-catcher_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import aissynth as S  # noqa: E402
import oracle as O  # noqa: E402

CASES = [
    # name, model, fs, N, nchunks, fmt, flags, seed, multi
    ("default_1536k", O.MODEL_DEFAULT, 1536000, 65536, 4, O.FMT_CF32, O.DEFAULT_FLAGS, 0, False),
    ("standard_1536k", O.MODEL_STANDARD, 1536000, 65536, 4, O.FMT_CF32, O.DEFAULT_FLAGS, 0, False),
    ("base_1536k", O.MODEL_BASE, 1536000, 65536, 4, O.FMT_CF32, O.DEFAULT_FLAGS, 0, False),
    ("default_phasesearch", O.MODEL_DEFAULT, 1536000, 32768, 8, O.FMT_CF32, O.FLAG_AFC_WIDE | O.FLAG_DROOP, 11, False),
    ("default_narrow_nodroop", O.MODEL_DEFAULT, 1536000, 32768, 8, O.FMT_CF32, O.FLAG_PS_EMA, 13, False),
    ("default_cu8", O.MODEL_DEFAULT, 1536000, 65536, 4, O.FMT_CU8, O.DEFAULT_FLAGS, 17, False),
    ("default_multi", O.MODEL_DEFAULT, 1536000, 65536, 6, O.FMT_CF32, O.DEFAULT_FLAGS, 31, True),
    ("default_96k", O.MODEL_DEFAULT, 96000, 4096, 4, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
    ("default_288k", O.MODEL_DEFAULT, 288000, 12288, 4, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
    ("default_768k", O.MODEL_DEFAULT, 768000, 32768, 4, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
    ("default_6000k", O.MODEL_DEFAULT, 6000000, 262144, 3, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
    ("default_6144k", O.MODEL_DEFAULT, 6144000, 262144, 3, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
    ("default_12288k", O.MODEL_DEFAULT, 12288000, 524288, 3, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
    ("standard_6000k", O.MODEL_STANDARD, 6000000, 262144, 3, O.FMT_CF32, O.DEFAULT_FLAGS, 23, False),
]

REF_SAMPLE_B = "177KQJ5000G?tO`K>RA1wUbN0TKH"  # reference python/tests/test_decode.py:13
REF_TYPE5 = ["55O0W7`00001L@gCWGA2uItLth@DqtL5@F22220j1h742t0Ht0000000", "000000000000000"]  # :14-17, fill 2
REF_TYPE26_MAX = ["J1mg=5AEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEEE", "E" * 56, "E" * 56, "EEEEE@4SA@"]  # :18-23, fill 4

CTAPS = {"C_a": O.TAP_CA, "C_b": O.TAP_CB, "CGF_a": O.TAP_CGF_A, "CGF_b": O.TAP_CGF_B, "FC_a": O.TAP_FC_A, "FC_b": O.TAP_FC_B}
FTAPS = {"FM_a": O.TAP_FM_A, "FM_b": O.TAP_FM_B, "FR_a": O.TAP_FR_A, "FR_b": O.TAP_FR_B}
FTAPS.update({"DEC_a%d" % i: O.TAP_DEC_A0 + i for i in range(5)})
FTAPS.update({"DEC_b%d" % i: O.TAP_DEC_B0 + i for i in range(5)})


def case_input(fs, N, nchunks, fmt, seed, multi):
    x = S.random_stream(fs, N * nchunks, seed, multi_sentence=multi)[0]
    if fmt == O.FMT_CU8:
        return S.to_cu8(x), 2
    return x, 1


def fbits(v):
    return int(np.float32(v).view(np.uint32))


def run_model(Model, model, fs, N, nchunks, fmt, flags, raw, per):
    m = Model(model=model, sample_rate=fs, fmt=fmt, flags=flags, taps=True)
    hs = {k: hashlib.sha256() for k in list(CTAPS) + list(FTAPS)}
    cnt = {k: 0 for k in hs}
    chunks = []
    for c in range(nchunks):
        m.push(raw[c * N * per:(c + 1) * N * per])
        for k, t in CTAPS.items():
            a = m.tap_c(t)
            hs[k].update(a.tobytes())
            cnt[k] += len(a)
        for k, t in FTAPS.items():
            a = m.tap_f(t)
            hs[k].update(a.tobytes())
            cnt[k] += len(a)
        chunks.append([{"ch": q.channel, "nbits": q.nbits, "payload": q.payload.hex(), "nmea": q.nmea, "start": q.start_idx,
                        "end": q.end_idx, "level": fbits(q.level), "ppm": fbits(q.ppm)} for q in m.messages()])
    return {"messages": chunks, "taps": {k: [cnt[k], hs[k].hexdigest()] for k in hs}}


def file_case(Model):
    """The committed raw file: 0.5 s of CU8 @96 kSPS with six bursts carrying the reference's own known-answer payloads."""
    path = os.path.join(HERE, "burst_96k.cu8")
    raw = np.fromfile(path, dtype=np.uint8)
    N = 4096
    nchunks = len(raw) // 2 // N
    return run_model(Model, O.MODEL_DEFAULT, 96000, N, nchunks, O.FMT_CU8, O.DEFAULT_FLAGS, raw, 2), nchunks


def main():
    if not O.have_ref():
        sys.exit("oracle/_ref/libaisref.so missing: run `make -C oracle ref` where /root/reference exists")
    path = os.path.join(HERE, "burst_96k.cu8")
    if not os.path.exists(path):
        rng = np.random.default_rng(96)
        fs, n = 96000, 49152
        # payloads = the reference's own known-answer sentences (python/tests/test_decode.py:12-23) + its CI sample
        t5 = np.concatenate([S.payload_to_bits(REF_TYPE5[0]), S.payload_to_bits(REF_TYPE5[1], fill=2)])
        t26 = np.concatenate([S.payload_to_bits(q) for q in REF_TYPE26_MAX[:3]] + [S.payload_to_bits(REF_TYPE26_MAX[3], fill=4)])
        assert len(t5) == 424 and len(t26) == 1064
        bursts = [S.Burst(1500, "A", S.payload_to_bits(S.SAMPLE_A), amp=0.4, foffs=300.0, timing=0.3),
                  S.Burst(5000, "B", S.payload_to_bits(REF_SAMPLE_B), amp=0.25, foffs=-450.0, timing=0.7),
                  S.Burst(9000, "A", t5, amp=0.3, foffs=150.0, timing=0.9),
                  S.Burst(16000, "B", t26, amp=0.35, foffs=-200.0, timing=0.2),
                  S.Burst(30000, "A", S.payload_to_bits(S.CI_LINE), amp=0.15, foffs=-120.0, timing=0.1),
                  S.Burst(40000, "B", S.random_message_bits(rng), amp=0.5, foffs=700.0, timing=0.5)]
        S.to_cu8(S.render_stream(fs, n, bursts, noise_sigma=0.02, seed=5)).tofile(path)
    out = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libaisref.so (unmodified reference, strict IEEE flags)", "cases": {}}
    for name, model, fs, N, nchunks, fmt, flags, seed, multi in CASES:
        raw, per = case_input(fs, N, nchunks, fmt, seed, multi)
        r = run_model(O.RefModel, model, fs, N, nchunks, fmt, flags, raw, per)
        r.update({"model": model, "fs": fs, "N": N, "nchunks": nchunks, "fmt": fmt, "flags": flags, "seed": seed, "multi": multi,
                  "input_sha256": hashlib.sha256(np.ascontiguousarray(raw).tobytes()).hexdigest()})
        out["cases"][name] = r
        print(name, sum(len(c) for c in r["messages"]), "messages")
    r, nchunks = file_case(O.RefModel)
    r.update({"model": O.MODEL_DEFAULT, "fs": 96000, "N": 4096, "nchunks": nchunks, "fmt": O.FMT_CU8, "flags": O.DEFAULT_FLAGS,
              "file": "burst_96k.cu8"})
    out["cases"]["file_burst_96k_cu8"] = r
    print("file_burst_96k_cu8", sum(len(c) for c in r["messages"]), "messages")
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
