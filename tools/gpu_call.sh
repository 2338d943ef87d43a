#!/bin/bash
# GPU call 22: randomised configuration matrix
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_matrix.py -m gpu -q > gpurun_out/pytest22.log 2>&1; tail -30 gpurun_out/pytest22.log | cut -c1-900
