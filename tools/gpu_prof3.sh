#!/bin/bash
# round-2 final profiles: ncu --set full of the headline kernel (final build), of the general sweep and of the block kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_affine_w -s 4 -c 1 -o gpurun_out/r02_affine_final -f python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_affine_final.log 2>&1; echo "ncu affine rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_sweep -s 2 -c 1 -o gpurun_out/r02_sweep_general -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-high-order --no-config5 --no-elasticity --no-multifield > gpurun_out/r02_sweep_general.log 2>&1; echo "ncu sweep rc=$?"
cd tools; timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_blocks -s 1 -c 1 -o ../gpurun_out/r02_blocks -f python bench_multifield.py --n 12 --m 8 --steps 1 > ../gpurun_out/r02_blocks.log 2>&1; echo "ncu blocks rc=$?"; cd ..
ls -la gpurun_out/*.ncu-rep
