#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dmma.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_dmma.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_dmma.log
timeout 900 python tools/bench_highorder.py --n 64 > gpurun_out/q3_n64.json 2> gpurun_out/q3_n64.err; echo "q3 n64 rc=$?"; tail -c 1500 gpurun_out/q3_n64.json; tail -3 gpurun_out/q3_n64.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elem_laplace_dmma -s 2 -c 1 -o gpurun_out/prof_dmma -f python tools/bench_highorder.py --n 32 --steps 2 --no-check > gpurun_out/prof_dmma.log 2>&1; echo "ncu rc=$?"
