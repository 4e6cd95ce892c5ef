#!/bin/bash
# Round 2, call h: tests of the current tree, fresh tuning, and the XF kernels with four instead of eight transform warps
# (build/xf128.so: 320 threads, no register cap at 128).
tag=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_generator.py tests/test_gpu_clip.py -m gpu -x -q 2>&1 | tail -4
for v in base xf128; do
  if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
  export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200_${tag}_$v.txt
  rm -f $RIB_TUNE_FILE
  RIB_NO_TUNE_TABLE=1 timeout 600 python tools/conv_bench.py --iters 2 --out gpurun_out/conv_events_tuning_${tag}_$v.txt
  RIB_NO_TUNE_TABLE=1 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_$v.txt
  RIB_NO_TUNE_TABLE=1 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_${v}_2.txt
done
