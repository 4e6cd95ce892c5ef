#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_x5.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_x5.log
python bench.py --steps 10 --no-cpu-baseline > gpurun_out/bench_x5.json 2> gpurun_out/bench_x5.err; cut -c1-420 gpurun_out/bench_x5.json; tail -3 gpurun_out/bench_x5.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"raster|pack|avgpool|warp|composite" -c 12 --csv --log-file gpurun_out/misc_x5.csv python tools/profile_forward.py --clip --iters 1 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/misc_x5.csv
