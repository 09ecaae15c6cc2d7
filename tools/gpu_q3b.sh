#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_highorder.py --n 64 > gpurun_out/q3_n64.json 2> gpurun_out/q3_n64.err; echo "q3 n64 rc=$?"; tail -c 1700 gpurun_out/q3_n64.json; tail -3 gpurun_out/q3_n64.err
