#!/bin/bash
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do
  RIB_PDL=0 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_nopdl_$rep.txt
  timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_base_$rep.txt
done
RIB_PDL=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-aten --no-c4 2>/dev/null | cut -c1-140
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-aten --no-c4 2>/dev/null | cut -c1-140
