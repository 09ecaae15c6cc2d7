#!/bin/bash
# one GPU call: full -m gpu suite, then bench N=1 (both arms)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_n1.json","gpurun_out/bench_ref.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        keep={k:d.get(k) for k in ("value","ms_per_step","symbolic_ms","reassembly","first_assembly","e2e","roofline","general_path","unstructured_path","config5")}
        print(f, json.dumps(keep)[:3000])
        if "high_order" in d: print("high_order", json.dumps(d["high_order"])[:2500])
    except Exception as e: print(f, "ERR", e)
PY
