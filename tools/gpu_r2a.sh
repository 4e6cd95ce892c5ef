#!/bin/bash
# Round 2, first GPU call: parity on the new tests, bench line (aten baseline, c4, lean upload), same-box A/B of the
# prepared kernel variants, the configs[4] sweep, and source-level captures of three epilogue-bound launches.
tag=${1:-r2a}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$tag.log
for v in base acc16 acc32 bkc16; do
  if [ $v = base ]; then unset RIB_LIB; else export RIB_LIB=$PWD/render-in-between_b200/build/$v.so; fi
  timeout 300 python tools/conv_bench.py --out gpurun_out/conv_events_${tag}_$v.txt
done
unset RIB_LIB
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
# tuning entries for the batch-16 (4x clips) and sweep shapes, written to a fresh table on this box
export RIB_TUNE_FILE=$PWD/gpurun_out/tune_b200_$tag.txt
cp render-in-between_b200/rib/tune_b200.txt $RIB_TUNE_FILE
timeout 900 tools/gpu_sweep.sh $tag > gpurun_out/sweep_$tag.log 2>&1; tail -10 gpurun_out/sweep_$tag.log
unset RIB_TUNE_FILE
timeout 600 tools/ncu_src.sh $tag 0 5 58 > /dev/null 2>&1
ls gpurun_out | grep $tag | head -50
