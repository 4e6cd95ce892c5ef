#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x6.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_x6.log
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x6.txt
timeout 200 python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_x6.json 2> gpurun_out/bench_x6.err; cut -c1-420 gpurun_out/bench_x6.json; tail -3 gpurun_out/bench_x6.err
