#!/bin/bash
# GPU call r2b-7: per-chain front-end shape (coherent: ring 3, one balanced wave), V2 engine with one converged decoder pass per group
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_all.log 2>&1; tail -5 gpurun_out/pytest_all.log | cut -c1-800
timeout 600 python tools/default_probe.py 11 - > gpurun_out/probe7_m11.jsonl 2>&1; cat gpurun_out/probe7_m11.jsonl | cut -c1-300
timeout 600 python tools/default_probe.py 2 - AISGPU_ST_NB=5 > gpurun_out/probe7_m2.jsonl 2>&1; cat gpurun_out/probe7_m2.jsonl | cut -c1-300
timeout 600 python tools/default_probe.py 4 - > gpurun_out/probe7_m4.jsonl 2>&1; cat gpurun_out/probe7_m4.jsonl | cut -c1-300
timeout 600 python tools/default_probe.py 0 - > gpurun_out/probe7_m0.jsonl 2>&1; cat gpurun_out/probe7_m0.jsonl | cut -c1-300
