#!/bin/bash
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200_$tag.txt
rm -f $RIB_TUNE_FILE
RIB_NO_TUNE_TABLE=1 timeout 600 python tools/conv_bench.py --iters 2 --out gpurun_out/conv_events_tuning_$tag.txt
RIB_NO_TUNE_TABLE=1 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_1.txt
RIB_NO_TUNE_TABLE=1 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_2.txt
grep -E "p5" gpurun_out/conv_events_tuning_$tag.txt.tune | cut -c1-420
