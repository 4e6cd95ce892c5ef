#!/bin/bash
# GPU tests + ncu launch list of the bench command (no bench line).  tools/gpu_quick2.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
cp gpurun_out/parity_report.txt gpurun_out/parity_report_$tag.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread \
    --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aten --no-c4 > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu bench rc=$?"
