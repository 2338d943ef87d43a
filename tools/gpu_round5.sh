#!/bin/bash
mkdir -p gpurun_out
for rpw in 6 3 1; do
  AISGPU_DEC_RPW=$rpw timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_d3_rpw$rpw.log 2>&1; echo "decoder3 rpw=$rpw pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_d3_rpw$rpw.log | cut -c1-600
done
timeout 600 python tools/fe_sweep.py 4,0,4096,AISGPU_DECODER=2 4,0,4096,AISGPU_DECODER=3,AISGPU_DEC_RPW=6 4,0,4096,AISGPU_DEC_RPW=3 4,0,4096,AISGPU_DEC_RPW=1 2>&1 | tee gpurun_out/be_sweep.log
SWEEP_MODEL=2 timeout 600 python tools/fe_sweep.py 4,0,4096,AISGPU_DECODER=2,AISGPU_DEC_RPW=6 4,0,4096,AISGPU_DECODER=3,AISGPU_DEC_RPW=6 4,0,4096,AISGPU_DEC_RPW=3 2>&1 | tee -a gpurun_out/be_sweep.log
