#!/bin/bash
# N-GPU bench exactly as the driver launches it.  Usage (under gpurun --gpus N): tools/gpu_multi.sh N tag
n=$1; tag=$2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_${tag}_n$n.json 2> gpurun_out/bench_${tag}_n$n.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_${tag}_n$n.json; tail -5 gpurun_out/bench_${tag}_n$n.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $n --steps 1 --warmup 1 > gpurun_out/bench_ref_${tag}_n$n.json 2> gpurun_out/bench_ref_${tag}_n$n.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref_${tag}_n$n.json
