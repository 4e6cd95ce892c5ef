#!/bin/bash
# Source-level ncu capture (--set full --import-source on) of selected conv_gemm launches of one forward.
# Usage (under gpurun): tools/ncu_src.sh <tag> <skip1> [<skip2> ...]   (skip = 0-based index among conv_gemm launches)
# Leaves gpurun_out/src_<tag>_<skip>.{raw,source}.csv; the .ncu-rep files are removed (size cap).
tag=$1; shift
mkdir -p gpurun_out
for s in "$@"; do
  ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_gemm -s $s -c 1 -f -o gpurun_out/src_${tag}_$s \
      python tools/profile_forward.py --batch 32 --size 512 --iters 1 > gpurun_out/ncu_src_${tag}_$s.log 2>&1
  ncu -i gpurun_out/src_${tag}_$s.ncu-rep --page raw --csv > gpurun_out/src_${tag}_$s.raw.csv 2>/dev/null
  ncu -i gpurun_out/src_${tag}_$s.ncu-rep --page source --csv > gpurun_out/src_${tag}_$s.source.csv 2>/dev/null
  rm -f gpurun_out/src_${tag}_$s.ncu-rep
done
ls -la gpurun_out | tail -20
