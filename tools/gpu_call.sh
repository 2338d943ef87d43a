#!/bin/bash
# GPU call 16: front-end ring depth / CTA shape (co-resident CTAs) sweep
mkdir -p gpurun_out
for m in 0 2; do
SWEEP_MODEL=$m timeout 900 python tools/fe_sweep.py 4,0,4096 4,0,4096,AISGPU_ST_SHAPE=24 4,0,4096,AISGPU_ST_SHAPE=34 4,0,4096,AISGPU_ST_SHAPE=32 4,0,4096,AISGPU_ST_SHAPE=42 4,0,4096,AISGPU_ST_SHAPE=0 2>&1 | grep -v "^$" | tee -a gpurun_out/sweep16.jsonl
done
echo "== parity with the shapes"; for sh in 24 32; do AISGPU_ST_SHAPE=$sh timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "standard_1536k or default_1536k or bench_shape" 2>&1 | tail -2 | cut -c1-300; done
