#!/bin/bash
# Kernel tests, ncu launch list of the bench command and source-level captures of the bandwidth kernels.  tools/gpu_nonconv.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread \
    --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aten --no-c4 > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu bench rc=$?"
for k in raster_mark raster_paint warp_px avgpool3s2; do
  timeout 300 bash tools/ncu_kernel_src.sh ${tag}_$k $k 0; echo "$k rc=$?"
done
