#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
AISGPU_DEC_RPW=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "standard or base" > gpurun_out/pytest_gpu_rpw0.log 2>&1; echo "rpw0 pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_rpw0.log
timeout 600 python tools/fe_sweep.py $SWEEP_ARGS 2>&1 | tee gpurun_out/be_sweep.log
