#!/bin/bash
# One gpurun call that refreshes every judged artefact: GPU parity tests, smoke, the bench line, the tuning table, the
# ncu launch list of the bench command, per-layer ncu sections of the conv launches and one --set full capture.
# Usage (under gpurun): tools/gpu_round.sh <tag> [full-capture skip index among conv_gemm launches]
tag=$1; skip=${2:-63}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
# fresh tuning table from this box (the shipped one is ignored), written at exit and re-used by everything below so
# that no candidate launches show up under ncu
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200.txt
rm -f $RIB_TUNE_FILE
RIB_NO_TUNE_TABLE=1 python tools/conv_bench.py --iters 2 --out gpurun_out/conv_events_tuning_$tag.txt
export RIB_NO_TUNE_TABLE=1
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cat gpurun_out/bench_$tag.json
python tools/conv_bench.py --out gpurun_out/conv_events_$tag.txt
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__registers_per_thread \
    --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aten --no-c4 > gpurun_out/ncu_bench_$tag.log 2>&1
echo "ncu bench rc=$?"
tools/ncu_conv.sh $tag > /dev/null 2>&1
python tools/conv_report.py gpurun_out/plan_B32_512.txt gpurun_out/conv_$tag.csv > gpurun_out/conv_layers_$tag.txt 2>&1
tail -1 gpurun_out/conv_layers_$tag.txt
rm -f gpurun_out/conv_$tag.ncu-rep
tools/ncu_src.sh $tag $skip > /dev/null 2>&1
ls -la gpurun_out | tail -25
