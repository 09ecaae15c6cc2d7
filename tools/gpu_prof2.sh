#!/bin/bash
# round-2 profiles: ncu --set full of the headline kernel and of K4, launch list of a short bench run
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_affine_w -s 4 -c 1 -o gpurun_out/r02_affine -f python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_affine.log 2>&1; echo "ncu affine rc=$?"
cd tools; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_elasticity_w -s 1 -c 1 -o ../gpurun_out/r02_k4 -f python bench_elasticity.py --n 32 --steps 1 > ../gpurun_out/r02_k4.log 2>&1; echo "ncu k4 rc=$?"; cd ..
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/r02_launches.log 2>&1; echo "ncu launches rc=$?"
ls -la gpurun_out/*.ncu-rep
