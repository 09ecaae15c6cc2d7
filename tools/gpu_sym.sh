#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
cat > /tmp/sym.py <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np, gtk_b200
H, E = gtk_b200.hostprep, gtk_b200.engine
n = 128
mesh = H.cartesian_mesh((0,1,0,1,0,1),(n,n,n)); V = H.lagrange_space(mesh,1,"boundary"); tab = H.measure_tabulation(V,2)
eng = E.Engine(0)
eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes); eng.set_space(V.cell_dofs, V.n_free, V.n_dirichlet); eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
for i in range(4):
    t=time.perf_counter(); eng.matrix_symbolic(); eng.vector_symbolic(); print("symbolic", i, 1e3*(time.perf_counter()-t), "ms", flush=True)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_sym.csv python /tmp/sym.py > gpurun_out/sym.log 2>&1; echo "ncu rc=$?"
python /tmp/sym.py; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-high-order 2>&1 | grep -o "\"symbolic[a-z_]*\": [0-9.]*"
