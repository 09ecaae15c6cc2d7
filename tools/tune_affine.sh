#!/bin/bash
# headline kernel A/B over the scheduler knobs (single GPU, 128^3 and 512x512x64), device time per step
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" python tools/bench_slab.py --cells ${CELLS:-128,128,128} --steps 300 2>&1 | grep -o '"ms_per_step": [0-9.]*'; }
for CELLS in 128,128,128 512,512,64; do
  export CELLS
  echo "#### cells $CELLS"
  run GTK_AFFINE_SEG=6
  run GTK_AFFINE_GF=1.0
  run GTK_AFFINE_GF=1.5
  run GTK_AFFINE_GF=2.0
  run GTK_AFFINE_GF=3.0
  run GTK_AFFINE_GF=1.5 GTK_AFFINE_SMIN=2
  run GTK_AFFINE_GF=1.5 GTK_AFFINE_SMIN=4
  run GTK_AFFINE_GF=1.5 GTK_AFFINE_SMAX=12
  run GTK_AFFINE_GF=1.5 GTK_AFFINE_SMAX=40
  run GTK_AFFINE_GF=2.0 GTK_AFFINE_SMIN=2 GTK_AFFINE_SMAX=16
  run GTK_AFFINE_VARIANT=9 GTK_AFFINE_GF=1.5
  run GTK_AFFINE_VARIANT=10 GTK_AFFINE_GF=1.5
  run GTK_AFFINE_VARIANT=11 GTK_AFFINE_GF=1.5
  run GTK_AFFINE_VARIANT=6 GTK_AFFINE_GF=1.5
done 2>&1 | tee gpurun_out/tune_affine.txt
