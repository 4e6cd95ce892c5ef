#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x14_base.txt
for v in 4 5 6; do
RIB_LIB=$PWD/render-in-between_b200/build/exp$v.so timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_x14_exp$v.txt
done
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:raster -c 3 --csv --log-file gpurun_out/raster_x14.csv python tools/profile_forward.py --clip --iters 1 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/raster_x14.csv
