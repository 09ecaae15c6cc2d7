#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_multi.log
for mode in p2p; do
  if [ $mode = nccl ]; then export GTK_DISABLE_P2P=1; else unset GTK_DISABLE_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench $mode rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --cells 512,512,64 > gpurun_out/bench_c5slab_n2_$mode.json 2> gpurun_out/bench_c5slab_n2_$mode.err; echo "rc=$?"
  python - <<PY
import json
for f in ("gpurun_out/bench_n2_$mode.json", "gpurun_out/bench_c5slab_n2_$mode.json"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("$mode", f, d["ms_per_step"], d["value"], d["roofline"]["kernels_ms"])
    except Exception as e:
        print("$mode", f, "failed", e)
PY
done
tail -5 gpurun_out/bench_n2_p2p.err
