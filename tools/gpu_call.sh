#!/bin/bash
# GPU call r2b-3: front-end ring depth (shared-memory footprint) x lane split, live, per model
mkdir -p gpurun_out
timeout 900 python tools/default_probe.py 2 - AISGPU_ST_NB=3 AISGPU_ST_NB=4 AISGPU_ST_NB=3,AISGPU_ST_L=32 AISGPU_ST_NB=3,AISGPU_ST_L=16 > gpurun_out/probe3_m2.jsonl 2>&1; cat gpurun_out/probe3_m2.jsonl | cut -c1-300
timeout 900 python tools/default_probe.py 0 - AISGPU_ST_L=32 AISGPU_ST_NB=3 AISGPU_ST_NB=3,AISGPU_ST_L=32 AISGPU_ST_NB=4,AISGPU_ST_L=32 AISGPU_ST_NB=3,AISGPU_ST_L=16 > gpurun_out/probe3_m0.jsonl 2>&1; cat gpurun_out/probe3_m0.jsonl | cut -c1-300
timeout 600 python tools/default_probe.py 4 - AISGPU_ST_NB=3 > gpurun_out/probe3_m4.jsonl 2>&1; cat gpurun_out/probe3_m4.jsonl | cut -c1-300
timeout 600 python tools/default_probe.py 11 - AISGPU_ST_NB=3 AISGPU_ST_L=32 > gpurun_out/probe3_m11.jsonl 2>&1; cat gpurun_out/probe3_m11.jsonl | cut -c1-300
