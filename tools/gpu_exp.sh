#!/bin/bash
mkdir -p gpurun_out
for seg in 4 5 6 7 8 11; do
  echo -n "GTK_AFFINE_SEG=$seg: "
  GTK_AFFINE_SEG=$seg timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-high-order 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['frac'])"
done
free -g | head -2
( time timeout 1500 python bench.py --steps 10 --warmup 3 --cells 512,512,512 > gpurun_out/bench_512cube_n1.json 2> gpurun_out/bench_512cube_n1.err ) 2>&1 | tail -3; echo "rc=$?"
tail -c 1200 gpurun_out/bench_512cube_n1.json; tail -3 gpurun_out/bench_512cube_n1.err
