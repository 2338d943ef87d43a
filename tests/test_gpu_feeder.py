"""aisgpu_feed_files -- one recording per stream, the batch form of the reference's Device::RAWFile
(FileRAW.cpp:36-165) -- against the oracle fed with the same bytes, zero-padded to whole blocks as RAWFile pads its tail."""
import numpy as np
import pytest

import aisgpu
import aissynth as S
import oracle as O

pytestmark = pytest.mark.gpu


def oracle_model(**kw):
    return (O.RefModel if O.have_ref() else O.PortModel)(**kw)


@pytest.mark.parametrize("fmt", [aisgpu.FMT_CU8, aisgpu.FMT_CF32])
def test_feed_files_matches_oracle(built, tmp_path, fmt):
    fs, N, B = 1536000, 32768, 6
    lengths = [N * 7 + 1000, N * 9, N * 3 + 17, 5 * N - 2, 1, N * 9 - 1]  # ragged: whole blocks, partial tails, a one-sample file
    paths, raws = [], []
    for s in range(B):
        x = S.random_stream(fs, N * 9, 7700 + s)[0][:lengths[s]]
        raw = S.to_cu8(x) if fmt == aisgpu.FMT_CU8 else np.ascontiguousarray(x).view(np.float32)
        p = tmp_path / ("rec%d.raw" % s)
        raw.tofile(p)
        paths.append(str(p))
        raws.append(raw)
    eng = aisgpu.Engine(model=aisgpu.MODEL_DEFAULT, sample_rate=fs, fmt=fmt, n_streams=B, max_chunk=N)
    got, nblocks = eng.feed_files(paths, N)
    assert nblocks == 9  # the longest recording: 9 blocks; the run ends with it
    total = 0
    for s in range(B):
        padded = np.zeros(nblocks * N * 2, dtype=raws[s].dtype)  # RAWFile zero-fills the tail of its last block (FileRAW.cpp:91-94)
        padded[:len(raws[s])] = raws[s]
        ref = oracle_model(model=O.MODEL_DEFAULT, sample_rate=fs, fmt=fmt)
        for c in range(nblocks):
            ref.push(padded[c * N * 2:(c + 1) * N * 2])
        want = [(m.key(), m.start_idx, m.end_idx) for m in ref.messages()]
        have = [(m.key(), m.start_idx, m.end_idx) for m in got if m.stream == s]
        assert have == want, "stream %d" % s
        total += len(want)
    assert total >= 5  # 0.19 s per stream: a handful of bursts
    c = eng.counters()
    assert c[3] == nblocks and c[2] == nblocks * N
    eng.close()


def test_feed_files_errors(built, tmp_path):
    fs, N = 1536000, 4096
    eng = aisgpu.Engine(model=aisgpu.MODEL_STANDARD, sample_rate=fs, fmt=aisgpu.FMT_CU8, n_streams=2, max_chunk=N)
    good = tmp_path / "a.raw"
    np.zeros(100, dtype=np.uint8).tofile(good)
    with pytest.raises(aisgpu.AisGpuError, match="cannot open input file"):
        eng.feed_files([str(good), str(tmp_path / "missing.raw")], N)
    with pytest.raises(aisgpu.AisGpuError):
        eng.feed_files([str(good), str(good)], N + 2)  # longer than max_chunk_samples
    empty = tmp_path / "empty.raw"
    empty.write_bytes(b"")
    got, nb = eng.feed_files([str(empty), str(empty)], N)  # nothing to read: no block is submitted
    assert got == [] and nb == 0
    eng.close()
