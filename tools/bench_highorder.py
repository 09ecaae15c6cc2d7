"""BASELINE config 3 (3D Poisson, Q3 hexahedra, 64^3 cells): numeric re-assembly through the DMMA element-GEMM path.
Prints one JSON object (also imported by bench.py for its `high_order` entry).

    python tools/bench_highorder.py [--n 64] [--order 3] [--steps 5]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DMMA_PEAK_TFLOPS = 37.0   # measured: 16 cycles per DMMA.8x8x4 per SM sub-partition (profiles/r01_dmma_peak.txt)


def run(n=64, order=3, steps=5, warmup=2, device=0, check=True):
    import numpy as np
    import gtk_b200
    H, E = gtk_b200.hostprep, gtk_b200.engine
    t0 = time.perf_counter()
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
    V = H.lagrange_space(mesh, order, "boundary")
    tab = H.measure_tabulation(V, 2 * order)
    host_s = time.perf_counter() - t0
    eng = E.Engine(device)
    eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
    eng.set_space(V.cell_dofs, V.n_free, V.n_dirichlet)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    t0 = time.perf_counter()
    nnz = eng.matrix_symbolic()
    symbolic_ms = 1e3 * (time.perf_counter() - t0)
    eng.set_profiling(True)
    times = []
    kern = {}
    for it in range(warmup + steps):
        eng.matrix_numeric_device(E.FORM_LAPLACE, alpha=1.0)
        recs = eng.profile()
        if it >= warmup:
            times.append(sum(ms for _, ms in recs))
            for name, ms in recs:
                kern[name] = kern.get(name, 0.0) + ms / steps
    fast = eng.info(5)
    ms = sum(times) / len(times)
    nld, nq = V.cell_dofs.shape[1], tab.w.size
    f_alg = mesh.n_cells * 2.0 * nld * nld * 3 * nq
    gemm_ms = kern.get("k_elem_laplace_dmma", float("nan"))
    out = {
        "workload": f"BASELINE config 3: 3D Poisson Q{order} hex {n}^3 cells, full Dirichlet boundary, numeric re-assembly on a cached pattern",
        "cells": int(mesh.n_cells), "n_ldofs": int(nld), "n_q": int(nq), "free_dofs": int(V.n_free), "nnz": int(nnz),
        "n_coo": int(eng.info(4)), "fast_path": int(fast),
        "ms_per_step": ms, "nnz_per_s": nnz / (ms * 1e-3), "dofs_per_s": V.n_free / (ms * 1e-3),
        "kernels_ms": kern, "symbolic_ms": symbolic_ms, "host_prep_s": host_s,
        "roofline": {"bound": "tensor", "unit": "TFLOP/s", "algorithmic_flops": f_alg,
                     "achieved": f_alg / (gemm_ms * 1e-3) / 1e12, "peak": DMMA_PEAK_TFLOPS,
                     "frac": f_alg / (gemm_ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
                     "kernel": "k_elem_laplace_dmma", "kernel_ms": gemm_ms,
                     "whole_step_frac": f_alg / (ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
                     "peak_source": "measured FP64 DMMA.8x8x4 issue rate on B200 (tools/dmma_peak.cu)"},
        "device_bytes": int(eng.info(2)),
    }
    if check:
        # size-independent properties: nnz formula for Q_k with full Dirichlet BC, zero row sums in the interior
        # (constants are in the kernel of the Laplacian), symmetry of the diagonal-block sums
        nz = eng.copy_nzval()
        cp, rv = eng.matrix_pattern()
        m = order * n - 1
        out["checks"] = {"n_free_ok": bool(V.n_free == m ** 3), "finite": bool(np.isfinite(nz).all()),
                         "sum_nz": float(nz.sum()), "diag_positive": None}
        col = np.repeat(np.arange(V.n_free, dtype=np.int64), np.diff(cp.astype(np.int64)))
        diag = nz[(rv.astype(np.int64) - 1) == col]
        out["checks"]["diag_positive"] = bool(diag.size == V.n_free and (diag > 0).all())
        colsum = np.bincount(col, weights=nz, minlength=V.n_free)
        # a column whose dof shares no cell with a Dirichlet dof sums to zero (constants are in the kernel)
        nzero = int((np.abs(colsum) <= 1e-10 * np.abs(nz).max()).sum())
        out["checks"]["zero_colsum_columns"] = nzero
        out["checks"]["zero_colsum_ok"] = bool(nzero >= (order * (n - 2) - 1) ** 3)
    eng.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.n, a.order, a.steps, check=not a.no_check)))
