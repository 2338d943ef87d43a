#!/bin/bash
# Final validation of a build: full GPU suite + smoke (bench.py: see tools/gpu_call2.sh / gpu_call3.sh for the multi-GPU and batch-8192 runs)
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; tail -3 gpurun_out/pytest_all.log | cut -c1-600
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
