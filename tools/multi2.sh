#!/bin/bash
# 2-GPU: ghost-row tests + weak-scaling bench with and without the overlapped exchange
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_multi.log
for mode in overlap serial; do
  if [ $mode = serial ]; then export GTK_DISABLE_OVERLAP=1; else unset GTK_DISABLE_OVERLAP; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n2_$mode.json") if l.startswith("{")][-1])
print("$mode", d["ms_per_step"], d["value"], d["roofline"]["kernels_ms"], d["e2e"]["ms_per_step"])
PY
done
tail -3 gpurun_out/bench_n2_overlap.err
