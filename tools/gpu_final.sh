#!/bin/bash
# Round-end evidence: bench (both arms), ncu launch lists, ncu full of the front-end kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --also-default > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
kill $SMI
bash tools/gpu_list.sh > /dev/null 2>&1
bash tools/gpu_prof_fe.sh > /dev/null 2>&1
cat gpurun_out/bench.json
