#!/bin/bash
# Round 2, call f: bias through the tensor core (in-tree) against the previous build (build/prev.so).
tag=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_generator.py tests/test_gpu_clip.py -m gpu -x -q 2>&1 | tail -4
for v in prev base prev base; do
  if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
  timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_$v.txt
done
