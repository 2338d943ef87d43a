#!/bin/bash
# GPU call 11: whole GPU suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest11.log 2>&1; tail -6 gpurun_out/pytest11.log | cut -c1-600
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 6000 gpurun_out/bench11.json; tail -3 gpurun_out/bench11.err
echo "== ref arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench11_ref.json 2>&1; tail -c 1200 gpurun_out/bench11_ref.json
