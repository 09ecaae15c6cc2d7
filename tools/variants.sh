mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 3 4 0 1; do
  echo "variant $v"; GTK_SWEEP_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks'])"
done
