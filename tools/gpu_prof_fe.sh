#!/bin/bash
# ncu --set full capture of one front-end launch (4th launch: after warm-up) on the bench workload
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_frontend -s 4 -c 1 -f -o gpurun_out/frontend python tools/fe_sweep.py ${SWEEP_ARGS:-4,0,4096} > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
