#!/bin/bash
# BASELINE configs[4]: resolution / batch sweep of the generator conv stack (256-1024 px, B in {1, 4, 16}) against the
# tensor / HBM roofline.  Per shape: per-layer CUDA-event times (tools/conv_bench.py), the plan, and a one-line summary
# (conv ms per forward, reference conv TFLOP/s, fraction of the sustained tensor peak of MEASURED_PEAKS.json).
# Usage (under gpurun): tools/gpu_sweep.sh <tag>
tag=$1
mkdir -p gpurun_out
: > gpurun_out/sweep_$tag.txt
for size in 256 512 1024; do
  for b in 1 4 16; do
    python tools/conv_bench.py --batch $b --size $size --iters 3 --out gpurun_out/sweep_${tag}_B${b}_${size}.txt > /dev/null 2>&1 || continue
    python - "$b" "$size" "gpurun_out/sweep_${tag}_B${b}_${size}.txt" >> gpurun_out/sweep_$tag.txt <<'PY'
import json, os, sys
b, size, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
us = float([l for l in open(path) if l.startswith('total')][0].split()[1])
wall = [l for l in open(path) if l.startswith('forward_wall')]
wall = float(wall[0].split()[1]) if wall else float('nan')
flop = 883.5e3 * size * size * b
peak = 1390.7
try:
    peak = json.load(open('MEASURED_PEAKS.json')).get('bf16_tflops_sustained', peak)
except OSError:
    pass
tf = flop / (us * 1e-6) / 1e12
print('B=%-3d %4dx%-4d conv kernels %9.1f us/forward  %7.1f TFLOP/s  %.3f of the sustained tensor peak | whole forward '
      '(wall, incl. launch gaps and non-conv kernels) %9.1f us  %7.1f frames/s' % (
          b, size, size, us, tf, tf / peak, wall, b / (wall * 1e-6)))
PY
  done
done
cat gpurun_out/sweep_$tag.txt
