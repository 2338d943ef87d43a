#!/bin/bash
# GPU call r2b-21: compute-sanitizer memcheck + racecheck over the round-2b kernels (lane-split front end, phase search 4b, double-buffered Ec)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python tools/sanitizer_workload.py > gpurun_out/r2j_sanitizer_memcheck.log 2>&1; tail -4 gpurun_out/r2j_sanitizer_memcheck.log | cut -c1-200
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitizer_workload.py > gpurun_out/r2j_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r2j_sanitizer_racecheck.log | cut -c1-200
