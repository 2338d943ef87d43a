#!/bin/bash
mkdir -p gpurun_out
for m in 0 2; do
SWEEP_MODEL=$m timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_m$m.csv python tools/fe_sweep.py 4,0,4096 > gpurun_out/ncu_list_m$m.log 2>&1
done
tail -2 gpurun_out/ncu_list_m2.log
