#!/bin/bash
# Final check of a tree: all GPU tests, smoke(), the bench line and the reference arm.  tools/gpu_final.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
cp gpurun_out/parity_report.txt gpurun_out/parity_report_$tag.txt 2>/dev/null
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$tag.json
