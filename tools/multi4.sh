#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fastpath.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
GTK_COMM_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --cells 512,512,64 > gpurun_out/t_c5.json 2> gpurun_out/t_c5.err; echo "rc=$?"
grep "overlap timeline" gpurun_out/t_c5.err | tail -4
GTK_COMM_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/t_n2.json 2> gpurun_out/t_n2.err; echo "rc=$?"
grep "overlap timeline" gpurun_out/t_n2.err | tail -4
for mode in overlap; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2_$mode.json 2> gpurun_out/bench_n2_$mode.err; echo "bench $mode rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --cells 512,512,64 > gpurun_out/bench_c5slab_n2_$mode.json 2> gpurun_out/bench_c5slab_n2_$mode.err; echo "rc=$?"
  python - <<PY
import json
for f in ("gpurun_out/bench_n2_$mode.json", "gpurun_out/bench_c5slab_n2_$mode.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print("$mode", f, d["ms_per_step"], d["value"], d["roofline"]["kernels_ms"])
PY
done
