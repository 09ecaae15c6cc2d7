mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 6 9 10 11; do
for w in 0 32; do
  echo "affine variant $v nseg $w"; GTK_AFFINE_VARIANT=$v GTK_AFFINE_NSEG=$w timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
done
bash tools/gpu_prof.sh k_q1hex_affine_w prof_affine_w5 A=1
