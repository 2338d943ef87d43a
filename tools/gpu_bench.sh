#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --also-default > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
kill $SMI
python -c "
import torch,time
x=torch.empty(1<<30,dtype=torch.uint8).pin_memory(); d=torch.empty(1<<30,dtype=torch.uint8,device='cuda')
for i in range(2): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); t=time.perf_counter()
for i in range(5): d.copy_(x,non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/5
print('pinned H2D 1 GiB: %.1f GB/s'%((1<<30)/dt/1e9))
" 2>&1 | tee gpurun_out/h2d.txt
cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
