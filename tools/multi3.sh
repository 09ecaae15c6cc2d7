#!/bin/bash
mkdir -p gpurun_out
for mode in overlap serial; do
  if [ $mode = serial ]; then export GTK_DISABLE_OVERLAP=1; else unset GTK_DISABLE_OVERLAP; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench $mode rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --cells 512,512,64 > gpurun_out/bench_c5slab_n2_$mode.json 2> gpurun_out/bench_c5slab_n2_$mode.err; echo "rc=$?"
  python - <<PY
import json
for f in ("gpurun_out/bench_n2_$mode.json", "gpurun_out/bench_c5slab_n2_$mode.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print("$mode", f, d["ms_per_step"], d["value"], d["roofline"]["kernels_ms"])
PY
done
