mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 3 1 4 0; do
for w in 8 16 32; do
  echo "sweep variant $v seg $w"; GTK_SWEEP_VARIANT=$v GTK_SWEEP_SEG=$w timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['general_path']['ms_per_step'], d['general_path']['roofline_frac'])"
done
done
