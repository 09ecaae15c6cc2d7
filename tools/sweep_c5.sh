#!/bin/bash
# 2 GPUs: knobs of the fused sweep + exchange at the config-5 slab size (512 x 512 x 64 cells per GPU); one line per setting
mkdir -p gpurun_out
out=gpurun_out/sweep_c5.txt
: > $out
run() {  # label, env...
  local label="$1"; shift
  local r
  r=$(env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/bench_slab.py --cells 512,512,64 --steps 30 2>/dev/null | grep "max ms_per_step")
  echo "$label | $r" | tee -a $out
}
run1() {
  local label="$1"; shift
  local r
  r=$(env "$@" timeout 300 python tools/bench_slab.py --cells 512,512,64 --steps 30 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.read().strip().splitlines()[-1])['ms_per_step'])")
  echo "$label | single GPU $r" | tee -a $out
}
run1 "1gpu default" X=1
run1 "1gpu smax32" GTK_AFFINE_SMAX=32
run1 "1gpu smax16" GTK_AFFINE_SMAX=16
run "default" X=1
run "edge6" GTK_FUSED_EDGE_SEG=6
run "edge12" GTK_FUSED_EDGE_SEG=12
run "edge22" GTK_FUSED_EDGE_SEG=22
run "edge12 il16" GTK_FUSED_EDGE_SEG=12 GTK_FUSED_INTERLEAVE=16
run "edge12 bf0.7" GTK_FUSED_EDGE_SEG=12 GTK_FUSED_BOTTOM_FROM=0.7
run "il16" GTK_FUSED_INTERLEAVE=16
run "skipall(timing only)" GTK_FUSED_DBG_SKIP=7
