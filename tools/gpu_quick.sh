#!/bin/bash
# Quick round: GPU tests, bench line, ncu launch list of the bench command (no per-layer sections).  tools/gpu_quick.sh <tag>
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
cp gpurun_out/parity_report.txt gpurun_out/parity_report_$tag.txt 2>/dev/null
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread \
    --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aten --no-c4 > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu bench rc=$?"
