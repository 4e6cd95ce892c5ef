#!/bin/bash
# Plan-to-plan noise of render_clips (tests/test_gpu_clip.py) for the in-tree library and A/B builds, with the auto-tuner
# on and off.  tools/gpu_clip_ab.sh <tag> <variant> ...   (variants = render-in-between_b200/build/<variant>.so)
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/clip_ab_$tag.txt; : > $out
for v in base "$@"; do
  for tune in 1 0; do
    if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
    echo "== lib=$v RIB_AUTOTUNE=$tune" >> $out
    RIB_AUTOTUNE=$tune timeout 300 python -m pytest tests/test_gpu_clip.py -m gpu -q -s -k "render_clips_matches" 2>&1 | grep -E "render_clips vs|passed|failed" >> $out
  done
done
unset RIB_LIB
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "avgpool" 2>&1 | tail -2 >> $out
cat $out
