#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 1 0 2; do
  RIB_XF=$m timeout 200 python bench.py --steps 20 --no-cpu-baseline 2> gpurun_out/bench_x7_$m.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RIB_XF=$m', d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'])"
done
RIB_XF=0 timeout 300 python -m pytest tests/test_gpu_generator.py tests/test_gpu_clip.py -x -q 2>&1 | tail -2
