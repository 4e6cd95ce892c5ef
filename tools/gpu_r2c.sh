#!/bin/bash
# Round 2, third GPU call: kernel tests of the pair / sub-pixel forms, fresh tuning table, and a same-box A/B of the
# two-epilogue-group builds (build/epi2a.so, build/epi2b.so) and of RIB_SUBPIX_PPC=4.
tag=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -3
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200_$tag.txt
rm -f $RIB_TUNE_FILE
RIB_NO_TUNE_TABLE=1 timeout 600 python tools/conv_bench.py --iters 2 --out gpurun_out/conv_events_tuning_$tag.txt
export RIB_NO_TUNE_TABLE=1
timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_base.txt
for v in epi2a epi2b; do
  RIB_LIB=$PWD/render-in-between_b200/build/$v.so timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_$v.txt
done
RIB_SUBPIX_PPC=4 timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_ppc4.txt
RIB_SUBPIX_PPC=4 RIB_LIB=$PWD/render-in-between_b200/build/epi2b.so timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_epi2b_ppc4.txt
timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_base2.txt
timeout 400 tools/ncu_src.sh $tag 63 > /dev/null 2>&1
python tools/src_roles.py gpurun_out/src_${tag}_63.source.csv | cut -c1-300
