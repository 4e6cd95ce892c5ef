#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
RIB_TUNE_FILE=gpurun_out/tune_b200.txt timeout 200 python bench.py --steps 20 --no-cpu-baseline 2> gpurun_out/bench_x8.err | cut -c1-250
wc -l gpurun_out/tune_b200.txt
RIB_TUNE_FILE=gpurun_out/tune_b200.txt timeout 200 python bench.py --steps 20 --no-cpu-baseline 2>> gpurun_out/bench_x8.err | cut -c1-250
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:raster -c 3 --csv --log-file gpurun_out/raster_x8.csv python tools/profile_forward.py --clip --iters 1 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/raster_x8.csv
timeout 120 tools/probe/tma_probe > gpurun_out/tma_probe.txt 2>&1; cat gpurun_out/tma_probe.txt
