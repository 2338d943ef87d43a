"""Shared helpers for the golden-vector tests (tests/golden/golden.json, generated from the unmodified reference by
tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np

import aissynth as S

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "golden.json")

FMT_CF32, FMT_CU8 = 0, 1


def load():
    with open(GOLDEN) as f:
        return json.load(f)["cases"]


def case_input(case):
    """(raw numpy array, elements per complex sample).  Seeded cases verify the generator against input_sha256."""
    if "file" in case:
        return np.fromfile(os.path.join(HERE, "golden", case["file"]), dtype=np.uint8), 2
    x = S.random_stream(case["fs"], case["N"] * case["nchunks"], case["seed"], multi_sentence=case["multi"])[0]
    raw, per = (S.to_cu8(x), 2) if case["fmt"] == FMT_CU8 else (x, 1)
    got = hashlib.sha256(np.ascontiguousarray(raw).tobytes()).hexdigest()
    assert got == case["input_sha256"], "seeded generator no longer reproduces the golden input (numpy RNG change?)"
    return raw, per


def fbits(v):
    return int(np.float32(v).view(np.uint32))


def msg_record(ch, nbits, payload, nmea, start, end, level, ppm):
    return {"ch": ch, "nbits": nbits, "payload": payload.hex(), "nmea": list(nmea), "start": start, "end": end,
            "level": fbits(level), "ppm": fbits(ppm)}
