#!/bin/bash
mkdir -p gpurun_out
SWEEP_MODEL=${SWEEP_MODEL:-0} timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_fm_fir5|k_decode3}" -s 8 -c 2 -f -o gpurun_out/be_prof python tools/fe_sweep.py 4,0,4096 > gpurun_out/ncu_be.log 2>&1
tail -3 gpurun_out/ncu_be.log
