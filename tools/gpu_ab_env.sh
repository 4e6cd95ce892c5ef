#!/bin/bash
# same-box A/B of an environment switch: tools/gpu_ab_env.sh VAR valueA valueB
mkdir -p gpurun_out
export RIB_NO_TUNE_TABLE=1
for v in $2 $3 $2 $3; do
  env $1=$v timeout 200 python bench.py --steps 20 --no-cpu-baseline 2>> gpurun_out/bench_abenv.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1=$v', round(d['value'],1), round(d['ms_per_step'],3), round(d['roofline']['kernel_ms_per_step'],3), round(d['e2e']['value'],1))"
done
env $1=$3 timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -2
env $1=$3 timeout 200 python tools/conv_bench.py --out gpurun_out/conv_events_abenv.txt
