#!/bin/bash
# Round 2, call g: fresh tuning of the in-tree build (bias MMA, fast epilogues, pairs), then the same for the build whose
# narrow-tile kernels are compiled for two CTAs per SM (build/minb2.so), per-layer times of both.
tag=${1:-r2g}
mkdir -p gpurun_out
for v in base minb2; do
  if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
  export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200_${tag}_$v.txt
  rm -f $RIB_TUNE_FILE
  RIB_NO_TUNE_TABLE=1 timeout 600 python tools/conv_bench.py --iters 2 --out gpurun_out/conv_events_tuning_${tag}_$v.txt
  RIB_NO_TUNE_TABLE=1 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_$v.txt
  RIB_NO_TUNE_TABLE=1 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_${v}_2.txt
done
