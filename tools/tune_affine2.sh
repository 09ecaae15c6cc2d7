#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python tools/bench_slab.py --cells ${CELLS:-128,128,128} --steps 300 2>&1 | grep -o '"ms_per_step": [0-9.]*'; }
for CELLS in 128,128,128 512,512,64; do
  export CELLS
  echo "#### cells $CELLS"
  for V in 11 12 13 14 15 16 17 6; do run GTK_AFFINE_VARIANT=$V GTK_AFFINE_GF=2.0; done
  for GF in 1.5 3.0 4.0; do run GTK_AFFINE_VARIANT=11 GTK_AFFINE_GF=$GF; done
  run GTK_AFFINE_VARIANT=11 GTK_AFFINE_GF=2.0 GTK_AFFINE_SMAX=16
  run GTK_AFFINE_VARIANT=11 GTK_AFFINE_GF=2.0 GTK_AFFINE_SMAX=32
  run GTK_AFFINE_VARIANT=11 GTK_AFFINE_GF=2.0 GTK_AFFINE_SMIN=2
  run GTK_AFFINE_VARIANT=11 GTK_AFFINE_GF=3.0 GTK_AFFINE_SMIN=2 GTK_AFFINE_SMAX=16
done 2>&1 | tee gpurun_out/tune_affine2.txt
