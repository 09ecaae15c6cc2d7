#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n8.json") if l.startswith("{")][-1])
print("n8", d["ms_per_step"], d["value"], d["roofline"]["kernels_ms"])
PY
tail -2 gpurun_out/bench_n8.err
