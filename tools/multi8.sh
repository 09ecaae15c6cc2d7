#!/bin/bash
# 8 GPUs: the multi-rank parity tests, then the bench line the driver's scaling run produces at N = 8 (weak 8 x 128^3 headline
# + the config5 object: 512^3 cells as 8 z-slabs with T_1 measured in the same job)
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print({k:d.get(k) for k in ("value","ms_per_step","symbolic_ms","e2e","partition")})
print(json.dumps(d.get("config5"))[:1500])
PY
tail -3 gpurun_out/bench_n$N.err
