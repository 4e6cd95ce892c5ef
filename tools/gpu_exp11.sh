#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
RIB_LIB=$PWD/render-in-between_b200/build/base.so timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x11_base.txt
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x11_new.txt
RIB_LIB=$PWD/render-in-between_b200/build/base.so timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x11_base2.txt
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x11_new2.txt
