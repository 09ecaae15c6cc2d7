#!/bin/bash
# usage: gpu_prof.sh <kernel regex> <out name> [env assignments...]
K=$1; OUT=$2; shift 2
mkdir -p gpurun_out
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/$OUT -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-high-order > gpurun_out/${OUT}.log 2>&1; echo "ncu-full rc=$?"
