#!/bin/bash
# 2-GPU run: bench.py under torchrun (own arm + reference arm), NCCL counters through the C library
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2j_bench_2gpu.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ['value','n_gpus','ms_per_step','per_rank_ms','counters_allreduce_c','msgs_per_s']})
print(d['e2e'], d.get('parity'))
PY
tail -3 gpurun_out/r2j_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2j_bench_2gpu_reference_arm.json 2>/dev/null; cut -c1-250 gpurun_out/r2j_bench_2gpu_reference_arm.json
