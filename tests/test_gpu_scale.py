"""-m gpu parity tests at the scale the bench numbers are quoted on, plus the concurrency and decoder stress cases.

* BASELINE.json configs[1] launch shape (batch 1024 x 131072 samples @1536 kS/s: one lane sub-segment of 4096 samples,
  256 four-warp CTAs, the decoder kernel over 2048 rows, back-end stages pipelined over two streams) for ModelStandard and
  ModelDefault, and configs[2] (batch 4096 @6 MSPS through Upsample, PhaseSearchEMA and PhaseSearch): sampled streams are
  compared with the unmodified reference (oracle/_ref/libaisref.so) -- NMEA sentences, their order, start/end counters.
* K submits back to back without any poll or synchronisation in between (the three-deep 48 kHz ring, the two back-end
  streams with per-stage events, the double-buffered input staging and the speculative Rotate table all run concurrently),
  for aisgpu_submit, aisgpu_submit_v, aisgpu_submit_async and aisgpu_submit_device.
* Decoder fuzz: >= 10 000 bursts of 40..1064 bits with stuffing-heavy payloads, invalid types, CRC failures, collisions and
  false start flags, chunk lengths from 256 samples to 131072 (frames straddle every kind of chunk boundary), for the
  word-parallel decoder with 1/3/6 rows per warp and the bit-serial cross-check kernel.
"""
import os
import threading

import numpy as np
import pytest

import aisgpu
import aissynth as S
import oracle as O

pytestmark = pytest.mark.gpu


def oracle_model(**kw):
    return (O.RefModel if O.have_ref() else O.PortModel)(**kw)


def run_oracle(streams, model, fs, chunk, fmt=O.FMT_CF32, ps_ema=True):
    """streams: {id: array}.  One oracle instance per stream, host threads in parallel (ctypes releases the GIL)."""
    flags = (O.FLAG_PS_EMA if ps_ema else 0) | O.FLAG_AFC_WIDE | O.FLAG_DROOP
    out = {}
    ids = list(streams)
    nthr = min(len(ids), max(1, len(os.sched_getaffinity(0))))

    def work(sub):
        for s in sub:
            m = oracle_model(model=model, sample_rate=fs, fmt=fmt, flags=flags)
            m.run(streams[s], chunk)
            out[s] = [(q.key(), q.start_idx, q.end_idx) for q in m.messages()]

    ths = [threading.Thread(target=work, args=(ids[t::nthr],)) for t in range(nthr)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return out


def collect(msgs, ids):
    got = {s: [] for s in ids}
    for m in msgs:
        if m.stream in got:
            got[m.stream].append((m.key(), m.start_idx, m.end_idx))
    return got


def compare(got, want):
    bad = [(s, len(got[s]), len(want[s])) for s in want if got[s] != want[s]]
    assert not bad, "streams whose message list (NMEA, order, start/end idx) differs from the reference: %r" % (bad[:8],)
    return sum(len(v) for v in want.values())


def device_batch(fs, B, N, nchunks, n_unique, seed0):
    """[nchunks][B][N] complex64 on the GPU: stream b = unique[b % U] + its own noise realisation (as bench.py builds it)."""
    import torch
    dev = torch.device("cuda", 0)
    uniq = np.stack([S.random_stream(fs, N * nchunks, seed0 + u)[0] for u in range(n_unique)])
    ud = torch.view_as_complex(torch.from_numpy(uniq.view(np.float32)).to(dev).view(n_unique, N * nchunks, 2))
    x = torch.empty((nchunks, B, N), dtype=torch.complex64, device=dev)
    for b0 in range(0, B, n_unique):
        nb = min(n_unique, B - b0)
        x[:, b0:b0 + nb, :] = ud[:nb].view(nb, nchunks, N).permute(1, 0, 2)
    g = torch.Generator(device=dev)
    g.manual_seed(seed0)
    noise = torch.empty((B, N), dtype=torch.complex64, device=dev)
    for c in range(nchunks):
        torch.view_as_real(noise).normal_(0.0, 0.005, generator=g)
        x[c] += noise
    torch.cuda.synchronize()
    return x


def run_scale_case(model, fs, B, N, nchunks, n_sample, ps_ema=True, seed0=500):
    import torch
    x = device_batch(fs, B, N, nchunks, 16, seed0)
    eng = aisgpu.Engine(model=model, sample_rate=fs, n_streams=B, max_chunk=N, ps_ema=ps_ema, max_frames=1 << 19, host_staging=False)
    for c in range(nchunks):  # back to back, like the bench's timed region
        eng.submit_device(x[c].data_ptr(), N, N)
    msgs = eng.poll()
    assert eng.overflows == 0
    rng = np.random.default_rng(seed0)
    ids = sorted(int(s) for s in rng.choice(B, size=n_sample, replace=False))
    idx = torch.tensor(ids, device=x.device)
    host = x.index_select(1, idx).cpu().numpy()  # [nchunks][n_sample][N]
    streams = {s: np.ascontiguousarray(host[:, j, :]).reshape(-1) for j, s in enumerate(ids)}
    want = run_oracle(streams, model, fs, N, ps_ema=ps_ema)
    n = compare(collect(msgs, ids), want)
    # every stream decoded something and the batch total is consistent with the sample
    per_stream = len(msgs) / B
    assert n >= n_sample and per_stream > 0.5 * n / n_sample
    eng.close()
    del x
    torch.cuda.empty_cache()
    return n


@pytest.mark.parametrize("model", [aisgpu.MODEL_STANDARD, aisgpu.MODEL_DEFAULT])
def test_bench_shape_batch1024(built, model):
    # BASELINE.json configs[1]: the exact launch shape of bench.py, three chunks, 64 sampled streams
    run_scale_case(model, 1536000, 1024, 131072, 3, 64)


@pytest.mark.parametrize("ps_ema", [True, False])
def test_config2_batch4096_6msps(built, ps_ema):
    # BASELINE.json configs[2]: AirSpy shape, coherent chain, PhaseSearchEMA and Demod::PhaseSearch
    run_scale_case(aisgpu.MODEL_DEFAULT, 6000000, 4096, 65536, 3, 32, ps_ema=ps_ema, seed0=700)


@pytest.mark.parametrize("model", [aisgpu.MODEL_STANDARD, aisgpu.MODEL_DEFAULT])
@pytest.mark.parametrize("how", ["submit", "submit_v", "submit_async", "submit_device"])
@pytest.mark.parametrize("pipe", ["0", "1"])
def test_back_to_back_submits(built, model, how, pipe, monkeypatch):
    import torch
    monkeypatch.setenv("AISGPU_BE_PIPE", pipe)
    fs, N, K, B = 1536000, 32768, 10, 24
    xs = np.stack([S.random_stream(fs, N * K, 900 + s)[0] for s in range(B)])
    eng = aisgpu.Engine(model=model, sample_rate=fs, n_streams=B, max_chunk=N)
    keep = []
    if how == "submit_device":
        xd = torch.from_numpy(xs.view(np.float32)).cuda().view(B, K, N, 2).permute(1, 0, 2, 3).contiguous()
        torch.cuda.synchronize()
    last = -1
    for c in range(K):
        blk = np.ascontiguousarray(xs[:, c * N:(c + 1) * N])
        if how == "submit":
            eng.submit(blk, N)
        elif how == "submit_v":
            eng.submit_v([xs[s, c * N:(c + 1) * N] for s in range(B)], N)
        elif how == "submit_async":
            t = torch.from_numpy(blk.view(np.float32)).pin_memory()
            keep.append(t)  # the buffer stays alive and untouched until the poll below
            last = eng.submit_async_ptr(t.data_ptr(), N)
        else:
            eng.submit_device(xd[c].data_ptr(), N, N)
    assert how != "submit_async" or last == K - 1
    msgs = eng.poll()  # the one and only synchronisation
    want = run_oracle({s: xs[s] for s in range(B)}, model, fs, N)
    n = compare(collect(msgs, range(B)), want)
    assert n >= B
    assert all(m.chunk < K for m in msgs)
    eng.close()


def test_poll_upto_returns_only_finished_submits(built):
    fs, N, K, B = 1536000, 32768, 6, 8
    xs = np.stack([S.random_stream(fs, N * K, 1300 + s)[0] for s in range(B)])
    import torch
    eng = aisgpu.Engine(model=aisgpu.MODEL_DEFAULT, sample_rate=fs, n_streams=B, max_chunk=N)
    bufs = [torch.from_numpy(np.ascontiguousarray(xs[:, c * N:(c + 1) * N]).view(np.float32)).pin_memory() for c in range(K)]
    got = []
    prev = None
    for c in range(K):
        t = eng.submit_async_ptr(bufs[c].data_ptr(), N)
        assert t == c
        if prev is not None:
            part = eng.poll_upto(prev)
            assert all(m.chunk <= prev for m in part)
            got += part
        prev = t
    got += eng.poll_upto(prev)
    want = run_oracle({s: xs[s] for s in range(B)}, aisgpu.MODEL_DEFAULT, fs, N)
    compare(collect(got, range(B)), want)
    eng.close()


def test_ring_overflow_is_reported(built):
    fs, N, B = 1536000, 65536, 16
    xs = np.stack([S.random_stream(fs, N * 2, 1500 + s, bursts_per_sec=(12, 16))[0] for s in range(B)])
    eng = aisgpu.Engine(model=aisgpu.MODEL_STANDARD, sample_rate=fs, n_streams=B, max_chunk=N, max_frames=4)
    for c in range(2):
        eng.submit(np.ascontiguousarray(xs[:, c * N:(c + 1) * N]), N)
    msgs = eng.poll()
    c = eng.counters()
    assert eng.overflows == 1 and c[4] > 0 and len(msgs) <= 4
    # after the loss the engine keeps working and reports nothing further
    eng.submit(np.ascontiguousarray(xs[:, 0:N]), N)
    eng.poll()
    eng.close()


FUZZ_FS = 96000
FUZZ_N = 262144
_fuzz_oracle = {}


@pytest.fixture(scope="module")
def fuzz_streams():
    xs, nb = [], 0
    for s in range(160):
        x, n = S.fuzz_stream(FUZZ_FS, FUZZ_N, s)
        xs.append(x)
        nb += n
    assert nb >= 10000
    return np.stack(xs), nb


@pytest.mark.parametrize("model", [aisgpu.MODEL_STANDARD, aisgpu.MODEL_DEFAULT])
@pytest.mark.parametrize("chunk", [256, 1000, 4096, 32768, 131072])
def test_decoder_fuzz_chunks(built, fuzz_streams, model, chunk):
    xs, _ = fuzz_streams
    sel = list(range(0, 160, 5)) if chunk < 1000 else list(range(160))  # tiny chunks: 1024 submits, fewer streams
    x = xs[sel]
    B = len(sel)
    eng = aisgpu.Engine(model=model, sample_rate=FUZZ_FS, n_streams=B, max_chunk=chunk, max_frames=1 << 16)
    nfull = FUZZ_N // chunk
    msgs = []
    for c in range(nfull):
        eng.submit(np.ascontiguousarray(x[:, c * chunk:(c + 1) * chunk]), chunk)
        if c % 64 == 63:
            msgs += eng.poll()
    msgs += eng.poll()
    assert eng.overflows == 0
    want = run_oracle({j: x[j][:nfull * chunk] for j in range(B)}, model, FUZZ_FS, chunk)
    n = compare(collect(msgs, range(B)), want)
    assert n >= 20 * B
    eng.close()


@pytest.mark.parametrize("model", [aisgpu.MODEL_STANDARD, aisgpu.MODEL_DEFAULT])
@pytest.mark.parametrize("decoder,rpw", [("3", "1"), ("3", "3"), ("3", "6"), ("1", "1")])
def test_decoder_fuzz_kernels(built, fuzz_streams, model, decoder, rpw, monkeypatch):
    monkeypatch.setenv("AISGPU_DECODER", decoder)
    monkeypatch.setenv("AISGPU_DEC_RPW", rpw)
    xs, _ = fuzz_streams
    chunk = 8192
    B = xs.shape[0]
    eng = aisgpu.Engine(model=model, sample_rate=FUZZ_FS, n_streams=B, max_chunk=chunk, max_frames=1 << 16)
    nfull = FUZZ_N // chunk
    for c in range(nfull):
        eng.submit(np.ascontiguousarray(xs[:, c * chunk:(c + 1) * chunk]), chunk)
    msgs = eng.poll()
    key = (model, chunk)
    if key not in _fuzz_oracle:  # the four kernel variants share one oracle run
        _fuzz_oracle[key] = run_oracle({j: xs[j] for j in range(B)}, model, FUZZ_FS, chunk)
    compare(collect(msgs, range(B)), _fuzz_oracle[key])
    eng.close()
