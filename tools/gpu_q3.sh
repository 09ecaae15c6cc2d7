#!/bin/bash
# config 3: DMMA-path tests + Q3 64^3 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dmma.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_dmma.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_dmma.log
timeout 900 python tools/bench_highorder.py --n 64 --no-check > gpurun_out/q3_n64.json 2> gpurun_out/q3_n64.err; echo "q3 n64 rc=$?"; tail -c 1500 gpurun_out/q3_n64.json; tail -3 gpurun_out/q3_n64.err
