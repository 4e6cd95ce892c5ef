#!/bin/bash
mkdir -p gpurun_out
export RIB_NO_TUNE_TABLE=1
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200.txt
rm -f $RIB_TUNE_FILE
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x12.txt
timeout 200 python bench.py --steps 20 --no-cpu-baseline 2> gpurun_out/bench_x12.err | cut -c1-250
grep -E "mask.down|emb_" gpurun_out/conv_events_x12.txt.tune | cut -c1-250
