#!/bin/bash
# 8 GPUs: weak-scaling bench line (8 x 128^3) and BASELINE config 5 itself (512^3 cells = 8 z-slabs of 512x512x64)
mkdir -p gpurun_out
nvidia-smi -L | head -8
free -g | head -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
tail -c 1800 gpurun_out/bench_n8.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --cells 512,512,64 > gpurun_out/bench_config5_n8.json 2> gpurun_out/bench_config5_n8.err; echo "config5 rc=$?"
tail -c 1800 gpurun_out/bench_config5_n8.json
tail -3 gpurun_out/bench_config5_n8.err
