#!/bin/bash
# usage: gpu_multi.sh N [extra bench args]: bench.py under torchrun on N GPUs of one box
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 1200 gpurun_out/bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    keep={k:d.get(k) for k in ("value","ms_per_step","symbolic_ms","symbolic_first_ms","partition","cpu_affinity","reassembly","e2e","config5")}
    print(json.dumps(keep)[:3000]); print(json.dumps(d["roofline"]["kernels_ms"]))
except Exception as e: print("ERR", e)
PY
