#!/bin/bash
tag=${1:-r2k}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -4
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200_$tag.txt
cp render-in-between_b200/rib/tune_b200.txt $RIB_TUNE_FILE
for rep in 1 2; do
  RIB_GATHER=0 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_nogather_$rep.txt
  timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_base_$rep.txt
done
grep emb_1 gpurun_out/conv_events_${tag}_*.txt gpurun_out/conv_events_${tag}_base_1.txt.tune | cut -c1-300
