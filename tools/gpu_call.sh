#!/bin/bash
# GPU call 8: Challenger + V2 CU8 parity; sanitizer runs for profiles/
mkdir -p gpurun_out
echo "== challenger tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "challenger" > gpurun_out/pytest9.log 2>&1; tail -12 gpurun_out/pytest9.log | cut -c1-900
echo "== compute-sanitizer memcheck (smoke + one parity case per model)"
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "ais-catcher_b200"); sys.path.insert(0, "oracle")
import __graft_entry__ as g
g.smoke()
import test_gpu_parity as T, aisgpu
T.run_case(g, aisgpu.MODEL_STANDARD, 1536000, 32768, 3, 2, check_taps=False, seed0=3)
T.run_case(g, aisgpu.MODEL_DEFAULT, 1536000, 32768, 3, 2, check_taps=False, seed0=3)
T.run_case(g, aisgpu.MODEL_DEFAULT, 6000000, 65536, 3, 2, check_taps=False, seed0=3)
T.run_v2_case(1536000, 32768, 6, 2, False, 5)
T.run_msg_case(aisgpu.MODEL_CHALLENGER, 1536000, 32768, 6, 2, 7)
print("sanitizer workload done")
PY
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/r2_sanitizer_memcheck.log | cut -c1-300
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/r2_sanitizer_racecheck.log | cut -c1-300
