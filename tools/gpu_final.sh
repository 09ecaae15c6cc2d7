#!/bin/bash
# end-of-round validation on one GPU: smoke, GPU tests, bench (both arms), ncu launch list, ncu full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
./tools/dmma_peak > gpurun_out/dmma_peak.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 4500 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_affine_w -s 3 -c 1 -o gpurun_out/prof_affine -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-high-order > gpurun_out/ncu_full.log 2>&1; echo "ncu-full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_laplace_dmma -s 2 -c 1 -o gpurun_out/prof_dmma -f python tools/bench_highorder.py --n 32 --steps 2 --no-check > gpurun_out/prof_dmma.log 2>&1; echo "ncu-dmma rc=$?"
ls -la gpurun_out | head -30
