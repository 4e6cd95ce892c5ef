#!/bin/bash
# Source-level ncu capture of one launch of an arbitrary kernel of the clip renderer.
# Usage (under gpurun): tools/ncu_kernel_src.sh <tag> <kernel regex> <skip>
# Leaves gpurun_out/ksrc_<tag>.{raw,source,cuda}.csv (headline metrics, SASS view, CUDA-C line view).
tag=$1; k=$2; s=${3:-0}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$k -s $s -c 1 -f -o gpurun_out/ksrc_$tag \
    python tools/profile_forward.py --clip --iters 1 > gpurun_out/ncu_ksrc_$tag.log 2>&1
ncu -i gpurun_out/ksrc_$tag.ncu-rep --page raw --csv > gpurun_out/ksrc_$tag.raw.csv 2>/dev/null
ncu -i gpurun_out/ksrc_$tag.ncu-rep --page source --csv > gpurun_out/ksrc_$tag.source.csv 2>/dev/null
ncu -i gpurun_out/ksrc_$tag.ncu-rep --page source --print-source cuda --csv > gpurun_out/ksrc_$tag.cuda.csv 2>/dev/null
rm -f gpurun_out/ksrc_$tag.ncu-rep
