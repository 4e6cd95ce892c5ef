#!/bin/bash
# Reduced form of gpu_round.sh (no per-layer ncu sections, no --set full): tests, smoke, fresh tuning table, bench line,
# per-layer CUDA-event times and the ncu launch list of the bench command.   tools/gpu_round_lite.sh <tag>
tag=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200.txt
rm -f $RIB_TUNE_FILE
export RIB_NO_TUNE_TABLE=1
python tools/conv_bench.py --iters 2 --out gpurun_out/conv_events_tuning_$tag.txt
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python tools/conv_bench.py --out gpurun_out/conv_events_$tag.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread \
    --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aten --no-c4 > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu bench rc=$?"
