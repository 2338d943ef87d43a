#!/bin/bash
# GPU call 20: FM chain with the front end privileged over its back end
mkdir -p gpurun_out
timeout 900 python tools/default_probe.py 0 - AISGPU_PRIO=2 - AISGPU_PRIO=2 AISGPU_PRIO=2,AISGPU_DEC_RPW=3 AISGPU_PRIO=2,AISGPU_DEC_RPW=1 AISGPU_DEC_RPW=3 2>&1 | tee gpurun_out/probe20.jsonl
PROBE_SHAPE=8192,65536,1536000,3 timeout 600 python tools/default_probe.py 0 - AISGPU_PRIO=2 2>&1 | tee -a gpurun_out/probe20.jsonl
