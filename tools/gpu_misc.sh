#!/bin/bash
# parity tests + bench + device times of the bandwidth-bound (non-conv) kernels of one clip
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python bench.py --steps 20 --no-cpu-baseline 2> gpurun_out/bench_misc.err | cut -c1-250
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"raster|warp|avgpool|composite|pack_images|in_apply" -c 40 --csv --log-file gpurun_out/misc.csv python tools/profile_forward.py --clip --iters 1 > /dev/null 2>&1
python tools/launch_table.py gpurun_out/misc.csv
