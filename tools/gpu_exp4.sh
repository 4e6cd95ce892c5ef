#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x4.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_x4.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_x4.json 2> gpurun_out/bench_x4.err; cat gpurun_out/bench_x4.json
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref_x4.json 2> gpurun_out/bench_ref_x4.err; cat gpurun_out/bench_ref_x4.json; tail -4 gpurun_out/bench_ref_x4.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:raster -c 6 --csv --log-file gpurun_out/raster_x4.csv python tools/profile_forward.py --clip --iters 2 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/raster_x4.csv
