"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm prints exactly ONE JSON line on
stdout with the agreed keys (everything else goes to stderr), and the rate -> chain table of the C ABI."""
import json
import os
import subprocess
import sys

import pytest

import aisgpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(built):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-500:]
    lines = [l for l in p.stdout.decode().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MSamples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]


@pytest.mark.parametrize("fs,granule", [(96000, 4), (192000, 8), (288000, 64), (250000, 64), (300000, 16), (384000, 16), (768000, 32),
                                        (1000000, 64), (1536000, 64), (2000000, 128), (3072000, 128), (6000000, 256), (6144000, 256),
                                        (12288000, 512)])
def test_rate_table_granules(built, fs, granule):
    # every CIC stage needs an even block (reference DSP.cpp:94,135): granule = 2^(stages + 2); the /3 path re-blocks itself
    assert aisgpu.chunk_granule(fs) == granule


@pytest.mark.parametrize("fs", [95999, 12288001, 0])
def test_rate_table_rejects_out_of_range(built, fs):
    with pytest.raises(aisgpu.AisGpuError, match="between 96K and 12288K"):
        aisgpu.chunk_granule(fs)
