#!/bin/bash
# Round 2, call e: whole GPU suite (folder entry, resize kernel included), accumulator-prefetch depth A/B
# (in-tree = 1 chunk ahead, build/ld2.so, build/ld4.so), B=1 / B=4 wall-clock against kernel sums.
tag=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for v in base ld2 ld4 base ld2 ld4; do
  if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
  timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_$v.txt
done
unset RIB_LIB
for b in 1 4; do timeout 300 python tools/conv_bench.py --batch $b --size 512 --iters 5 --out gpurun_out/conv_events_${tag}_B${b}_512.txt; tail -2 gpurun_out/conv_events_${tag}_B${b}_512.txt; done
