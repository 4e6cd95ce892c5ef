#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --no-cpu-baseline 2> gpurun_out/bench_x13.err | cut -c1-250
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"warp|avgpool" -c 5 --csv --log-file gpurun_out/misc_x13.csv python tools/profile_forward.py --clip --iters 1 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/misc_x13.csv
