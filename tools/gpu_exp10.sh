#!/bin/bash
mkdir -p gpurun_out
export RIB_NO_TUNE_TABLE=1
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x10_base.txt
RIB_LIB=$PWD/render-in-between_b200/build/exp4.so timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x10_exp4.txt
