#!/bin/bash
# Same-box A/B of kernel-variant builds against the in-tree library with the shipped tuning table:
#   tools/gpu_ab_libs.sh <tag> <variant> [<variant> ...]     (variants = render-in-between_b200/build/<variant>.so)
# Every build is timed twice, interleaved (base v1 v2 ... base v1 v2 ...): per-layer CUDA-event times + whole-forward wall.
tag=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for v in base "$@"; do
    if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
    timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_${v}_$rep.txt
  done
done
