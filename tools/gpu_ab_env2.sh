#!/bin/bash
# Same-box A/B of an environment switch with the shipped tuning table (new launch shapes are tuned on the fly):
#   tools/gpu_ab_env2.sh <tag> <VAR> <value>
tag=$1; var=$2; val=$3
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_base_$rep.txt
  env $var=$val timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_${var}_$rep.txt
done
