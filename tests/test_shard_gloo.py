"""CPU (-m "not gpu"): the N>1 plumbing of bench.py / shard.py over gloo with world_size 2.  The per-rank "engine" is
the C restatement (the checker standing in for the GPU in a CPU-only test): the sharded job must report exactly the
totals and the per-stream message order of the unsharded one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stream_slices_cover_batch():
    for B in (0, 1, 7, 1024, 65536, 1000):
        for G in (1, 2, 3, 4, 8):
            sl = [shard.stream_slice(B, g, G) for g in range(G)]
            assert sl[0][0] == 0 and sl[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= 1
    with pytest.raises(ValueError):
        shard.stream_slice(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _decode_slice(lo, hi, fs, N, nchunks):
    import aissynth as S
    import oracle as O
    msgs, counts = [], [0] * 8
    for s in range(lo, hi):
        x = S.random_stream(fs, N * nchunks, 500 + s)[0]
        m = O.PortModel(model=O.MODEL_DEFAULT, sample_rate=fs)
        m.run(x, N)
        got = m.messages()
        msgs += [(s - lo, tuple(q.nmea)) for q in got]
        counts[1] += len(got)
        counts[5] += sum(q.channel == "A" for q in got)
        counts[6] += sum(q.channel == "B" for q in got)
    counts[2] = N * nchunks
    counts[3] = nchunks
    return msgs, counts


def _worker(rank, world, port, B, fs, N, nchunks, q):
    for p in (ROOT, os.path.join(ROOT, "ais-catcher_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import shard as sh
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = sh.stream_slice(B, rank, world)
    msgs, counts = _decode_slice(lo, hi, fs, N, nchunks)
    total = sh.gather_counts(counts)
    tmax = sh.max_over_ranks(10.0 + rank)
    allm = sh.gather_messages(msgs, lo)
    if rank == 0:
        q.put((total, tmax, allm))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(built):
    B, fs, N, nchunks = 6, 96000, 4096, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, fs, N, nchunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, tmax, allm = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    msgs1, counts1 = _decode_slice(0, B, fs, N, nchunks)
    assert counts1[1] > 0
    assert total[1] == counts1[1] and total[5] == counts1[5] and total[6] == counts1[6]
    assert total[2] == 2 * counts1[2] and total[3] == 2 * counts1[3]  # per-rank quantities add up over ranks
    assert tmax == 11.0
    assert [(s, n) for s, n in allm] == sorted(msgs1, key=lambda t: t[0])
