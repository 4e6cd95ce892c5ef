#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x3.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_x3.log
RIB_AUTOTUNE=0 python tools/conv_bench.py --out gpurun_out/conv_events_x3_notune.txt
python tools/conv_bench.py --out gpurun_out/conv_events_x3_tune.txt
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_x3.json 2> gpurun_out/bench_x3.err; cat gpurun_out/bench_x3.json
