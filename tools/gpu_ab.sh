#!/bin/bash
# same-box A/B of two builds: render-in-between_b200/build/base.so vs the in-tree library
mkdir -p gpurun_out
RIB_LIB=$PWD/render-in-between_b200/build/base.so timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_ab_base.txt
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_ab_new.txt
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --no-cpu-baseline 2> gpurun_out/bench_ab.err | cut -c1-200
