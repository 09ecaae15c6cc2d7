mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for v in 0 1 2 3 4; do
  echo "affine variant $v"; GTK_AFFINE_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['kernels_ms'], d['e2e']['ms_per_step'])"
done
for w in 2 8; do
  echo "affine variant 0 waves $w"; GTK_AFFINE_WAVES=$w timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
