#!/bin/bash
# experiment round: sub-pixel mask.up convs + MT knobs + source-level captures of the raster kernels and two small convs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_x2.log
python tools/conv_bench.py --out gpurun_out/conv_events_x2_base.txt
RIB_S2_MT2=1 python -m pytest tests/test_gpu_generator.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
RIB_S2_MT2=1 python tools/conv_bench.py --out gpurun_out/conv_events_x2_s2mt2.txt
RIB_SMALL_MT=2 python -m pytest tests/test_gpu_generator.py tests/test_gpu_kernels.py -x -q 2>&1 | tail -2
RIB_SMALL_MT=2 python tools/conv_bench.py --out gpurun_out/conv_events_x2_smallmt2.txt
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err; cat gpurun_out/bench_x2.json
tools/ncu_kernel_src.sh mark raster_mark 0
tools/ncu_kernel_src.sh paint raster_paint 0
tools/ncu_src.sh x2 5 75 > /dev/null 2>&1
ls gpurun_out | head -50
