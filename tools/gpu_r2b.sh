#!/bin/bash
# Round 2, second GPU call: the full evidence round (tools/gpu_round.sh: tests, smoke, fresh tuning table with the CTA-pair
# candidates, bench, per-layer times, ncu launch list, per-layer ncu, full capture of mask.res.0.conv1) plus source-level
# captures of three epilogue-bound launches (emb_0, label3x3, mask.down_img.0).
tag=${1:-r2m}
tools/gpu_round.sh $tag 63
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200.txt RIB_NO_TUNE_TABLE=1
timeout 900 tools/ncu_src.sh $tag 0 5 58 > /dev/null 2>&1
for s in 0 5 58 63; do python tools/src_roles.py gpurun_out/src_${tag}_$s.source.csv > gpurun_out/src_roles_${tag}_$s.txt 2>&1; done
cat gpurun_out/src_roles_${tag}_*.txt | cut -c1-400
ls gpurun_out | grep $tag | wc -l
