"""Host check of the exact-arithmetic restatements the kernels rely on: tests/host/atan2_check.c holds the common-case path of
fd_atan2f_common (csrc/exact.cuh) in C and compares it with the C library's atan2f -- the function the reference calls
(Demod.cpp:27-37) -- on 4e7 arguments, bit for bit.  The CUDA transcription itself is checked by the FM taps of the GPU tests."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_atan2_common_path_matches_libm(tmp_path):
    exe = str(tmp_path / "atan2_check")
    subprocess.check_call(["gcc", "-O2", "-fno-fast-math", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "host", "atan2_check.c"), "-lm"])
    out = subprocess.run([exe, "40000000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-800:]
    assert "mismatches 0" in out.stdout
    checked = int(out.stdout.split("checked")[1].split()[0])
    assert checked > 20000000  # most of the arguments must actually have taken the path under test
