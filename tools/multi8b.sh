#!/bin/bash
mkdir -p gpurun_out
export GTK_DISABLE_OVERLAP=1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 3 --cells 512,512,64 > gpurun_out/bench_config5_n8_serial.json 2> gpurun_out/bench_config5_n8_serial.err; echo "config5 serial rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_config5_n8_serial.json") if l.startswith("{")][-1])
print("serial", d["ms_per_step"], d["value"], d["roofline"]["kernels_ms"])
PY
unset GTK_DISABLE_OVERLAP
export GTK_COMM_TIMING=1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 4 --warmup 3 --cells 512,512,64 > gpurun_out/t8.json 2> gpurun_out/t8.err; echo "timing rc=$?"
grep "overlap timeline" gpurun_out/t8.err | sort | awk '{c[$3]++; if (c[$3]>3 && c[$3]<=5) print}' | head -20
