"""ais-catcher_b200/host/ModelGPU.h inside the reference's own block graph (tests/host/adapter_main.cpp linked with the
UNMODIFIED reference objects by `make -C oracle adapter`): MemDevice --Connection<RAW>--> AIS::ModelGPU --> message sink.
CPU: configuration errors surface as std::runtime_error with the reference's wording, a missing GPU as
Error()+StopRequest(); the same binary with the reference's CPU model reproduces the golden file.
GPU (-m gpu): the adapter prints exactly the golden messages."""
import os
import subprocess

import pytest

import golden_util as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "adapter_test")
FILE = os.path.join(ROOT, "tests", "golden", "burst_96k.cu8")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="oracle/_ref/adapter_test not built (needs /root/reference at build time)")


def run(*args, env=None):
    p = subprocess.run([EXE] + list(args), capture_output=True, env=env)
    return p.returncode, p.stdout.decode("latin-1"), p.stderr.decode("latin-1")


def golden_lines():
    out = []
    for chunk in G.load()["file_burst_96k_cu8"]["messages"]:
        for m in chunk:
            out.append("%s|%d|%d|%d|%d|%d|%s" % (m["ch"], m["nbits"], m["start"], m["end"], m["level"], m["ppm"], " ".join(m["nmea"])))
    return out


@needs_exe
def test_config_errors_use_reference_wording(built):
    rc, _, err = run(FILE, "CU8", "48000", "4096", "2")
    assert rc == 4 and "sample rate must be between 96K and 12288K" in err
    rc, _, err = run(FILE, "CU8", "96000", "4096", "2", "gpu", "SOXR", "on")
    assert rc == 4 and "soxr" in err.lower()


@needs_exe
def test_no_gpu_stops_the_application(built):
    rc, out, err = run(FILE, "CU8", "96000", "4096", "2", env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert rc == 3 and out == "" and "1 stop requests" in err


@needs_exe
def test_same_binary_with_reference_model_matches_golden(built):
    rc, out, _ = run(FILE, "CU8", "96000", "4096", "2", "cpu")
    assert rc == 0 and out.splitlines() == golden_lines()


@needs_exe
@pytest.mark.gpu
@pytest.mark.parametrize("block,model", [(4096, 2), (12288, 2), (2048, 2), (4096, 0), (16384, 1)])
def test_adapter_on_gpu_equals_reference_model_in_same_binary(built, block, model):
    """A/B inside one binary: AIS::ModelGPU vs the reference's own ModelDefault/Standard/Base behind the same device,
    same block length (message order and the float tags depend on it) -- identical output lines."""
    rc, out, err = run(FILE, "CU8", "96000", str(block), str(model))
    assert rc == 0, err
    rc2, want, _ = run(FILE, "CU8", "96000", str(block), str(model), "cpu")
    assert rc2 == 0
    assert out.splitlines() == want.splitlines()
    if block == 4096 and model == 2:
        assert out.splitlines() == golden_lines()


@needs_exe
@pytest.mark.gpu
@pytest.mark.parametrize("rate,block", [(96000, 4096), (6000000, 65536), (2000000, 16384)])
def test_adapter_reblocks_odd_and_varying_buffers(built, rate, block, tmp_path):
    """Device buffers of odd and varying length (after a first one of the nominal length) give exactly what the reference
    model gives for fixed buffers of that length -- also at interpolated rates, where the engine insists on one length."""
    import numpy as np
    import aissynth as S
    f = str(tmp_path / "x.cu8")
    S.to_cu8(S.random_stream(rate, block * 12, 77, bursts_per_sec=(20, 30))[0]).tofile(f)
    rc, out, err = run(f, "CU8", str(rate), str(block), "2", env=dict(os.environ, ADAPTER_VARY="1"))
    assert rc == 0, err
    rc2, want, _ = run(f, "CU8", str(rate), str(block), "2", "cpu")
    assert rc2 == 0 and out == want and len(want.splitlines()) >= 1
