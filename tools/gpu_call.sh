#!/bin/bash
# one GPU call: parity tests, then launch-shape sweeps of the streaming front end and the decoder on the bench workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for m in 0 2; do
SWEEP_MODEL=$m timeout 600 python tools/fe_sweep.py \
  4,0,4096,AISGPU_ST_WPC=4,AISGPU_ST_NB=8,AISGPU_ST_S=4096 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=4,AISGPU_ST_S=4096 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=2048 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096,AISGPU_DEC_RPW=1 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096,AISGPU_DEC_RPW=3 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096,AISGPU_DEC_RPW=6 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096,AISGPU_DEC_RPW=3,AISGPU_BE_PIPE=0 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096,AISGPU_DEC_RPW=3,AISGPU_BE_PIPE=1 \
  2>&1 | grep -v "^$" | tee -a gpurun_out/sweep1.jsonl
done
SWEEP_B=4096 SWEEP_MODEL=0 timeout 300 python tools/fe_sweep.py \
  4,0,4096,AISGPU_ST_WPC=4,AISGPU_ST_NB=8,AISGPU_ST_S=0 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=4096 \
  4,0,4096,AISGPU_ST_WPC=1,AISGPU_ST_NB=6,AISGPU_ST_S=8192 \
  2>&1 | grep -v "^$" | tee -a gpurun_out/sweep1.jsonl
