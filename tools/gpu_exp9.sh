#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x9_base.txt
for v in 1 2 3; do
RIB_LIB=$PWD/render-in-between_b200/build/exp$v.so timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x9_exp$v.txt
done
