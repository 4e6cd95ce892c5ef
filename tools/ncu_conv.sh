#!/bin/bash
# Per-launch ncu metrics for every implicit-GEMM launch of one forward (B200_PROFILING.md recipe).
# Usage (under gpurun): tools/ncu_conv.sh <tag> [batch] [size]
# Writes gpurun_out/conv_<tag>.csv (raw page) and gpurun_out/plan_B<batch>_<size>.txt; the .ncu-rep is
# deleted when it is larger than 40 MiB (gpurun_out/ is capped at 64 MiB).
tag=$1; b=${2:-32}; s=${3:-512}
mkdir -p gpurun_out
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section LaunchStats \
    --section Occupancy --section SchedulerStats --clock-control none --profile-from-start off -k regex:conv_gemm -c 100 \
    -o gpurun_out/conv_$tag -f python tools/profile_forward.py --batch $b --size $s --iters 1 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/conv_$tag.ncu-rep --page raw --csv > gpurun_out/conv_$tag.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/conv_$tag.ncu-rep)
if [ "$sz" -gt 41943040 ]; then rm -f gpurun_out/conv_$tag.ncu-rep; fi
ls -la gpurun_out
