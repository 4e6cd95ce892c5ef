#!/bin/bash
# Round 2, call d: specialised EPI_STORE epilogues against the generic code (build/nofast.so), same box, same tuning table.
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_generator.py -m gpu -x -q 2>&1 | tail -3
RIB_LIB=$PWD/render-in-between_b200/build/nofast.so timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_nofast.txt
timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_fast.txt
RIB_LIB=$PWD/render-in-between_b200/build/nofast.so timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_nofast2.txt
timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_fast2.txt
timeout 400 tools/ncu_src.sh $tag 0 5 > /dev/null 2>&1
for s in 0 5; do python tools/src_roles.py gpurun_out/src_${tag}_$s.source.csv | cut -c1-300; done
