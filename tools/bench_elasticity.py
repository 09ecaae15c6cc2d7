"""BASELINE config 4 (3D vector linear elasticity, P2 x 3 on the simplexified Cartesian mesh, Strang degree-4 rule):
numeric re-assembly on a cached pattern.  Prints one JSON object (also imported by bench.py for its `elasticity` entry).

    python tools/bench_elasticity.py [--n 64] [--steps 3]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run(n=64, steps=3, warmup=1, device=0, check=True):
    import numpy as np
    import gtk_b200
    H, E = gtk_b200.hostprep, gtk_b200.engine
    t0 = time.perf_counter()
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n), simplexify=True)
    V = H.lagrange_space(mesh, 2, "boundary", 3)
    tab = H.measure_tabulation(V, 4)
    host_s = time.perf_counter() - t0
    eng = E.Engine(device)
    eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
    eng.set_space(V.cell_dofs, V.n_free, V.n_dirichlet, 3)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    t0 = time.perf_counter()
    nnz = eng.matrix_symbolic()
    symbolic_ms = 1e3 * (time.perf_counter() - t0)
    eng.set_profiling(True)
    times, kern = [], {}
    params = dict(alpha=1.0, lam=1.0, mu=1.0)
    for it in range(warmup + steps):
        eng.matrix_numeric_device(E.FORM_ELASTICITY_ISO, **params)
        recs = eng.profile()
        if it >= warmup:
            times.append(sum(ms for _, ms in recs))
            for name, ms in recs:
                kern[name] = kern.get(name, 0.0) + ms / steps
    ms = sum(times) / len(times)
    peak, peak_src = measured_peak_gbs()
    nld, nln = V.cell_dofs.shape[1], mesh.cell_nodes.shape[1]
    alg = 8 * 3 * mesh.n_nodes + 4 * nln * mesh.n_cells + 4 * nld * mesh.n_cells + 8 * nnz
    n_coo = eng.info(4)
    out = {
        "workload": f"BASELINE config 4: 3D linear elasticity, P2 x 3 on the simplexified {n}^3 Cartesian mesh ({mesh.n_cells} tetrahedra), "
                    f"Strang degree-4 rule (11 points), full Dirichlet boundary, numeric re-assembly on a cached pattern",
        "cells": int(mesh.n_cells), "n_ldofs": int(nld), "n_q": int(tab.w.size), "free_dofs": int(V.n_free), "nnz": int(nnz),
        "n_coo": int(n_coo), "fast_path": int(eng.info(5)), "ms_per_step": ms, "nnz_per_s": nnz / (ms * 1e-3),
        "dofs_per_s": V.n_free / (ms * 1e-3), "kernels_ms": kern, "symbolic_ms": symbolic_ms, "host_prep_s": host_s,
        "roofline": {"bound": "hbm", "unit": "GB/s", "algorithmic_bytes_per_step": int(alg), "achieved": alg / (ms * 1e-3) / 1e9,
                     "peak": peak, "frac": alg / (ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                     "bytes_moved_by_construction": int(alg + 2 * 8 * n_coo + 4 * n_coo + 4 * nnz),
                     "note": "element matrices are staged once (write + read) and summed through a 4-byte slot index: 20 B per COO "
                             "entry on top of the compulsory bytes"},
        "device_bytes": int(eng.info(2)),
    }
    if check:
        # size-independent properties, on the device: rigid translations are in the kernel of the operator => column sums
        # vanish away from the Dirichlet boundary; symmetric pattern
        import torch
        from bench_highorder import _DevArr
        pn, nn = eng.device_pointer(0)
        pc, nc = eng.device_pointer(2)
        nz = torch.as_tensor(_DevArr(pn, nn, "<f8"), device="cuda")
        cp = torch.as_tensor(_DevArr(pc, nc, "<i8"), device="cuda")
        lens = cp[1:] - cp[:-1]
        colsum = torch.segment_reduce(nz, "sum", lengths=lens)
        scale = float(nz.abs().max().item())
        out["checks"] = {"finite": bool(torch.isfinite(nz).all().item()),
                         "zero_colsum_columns": int((colsum.abs() <= 1e-10 * scale).sum().item()),
                         "columns": int(V.n_free), "sum_nz": float(nz.sum().item())}
    eng.close()
    return out


def cpu_baseline(n=10):
    """C port of the reference loop on the host cores for the same element (bounded sample)."""
    import c_oracle
    import gtk_b200
    H = gtk_b200.hostprep
    threads = os.cpu_count() or 1
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n), simplexify=True)
    V = H.lagrange_space(mesh, 2, "boundary", 3)
    tab = H.measure_tabulation(V, 4)
    tabd = dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)
    c_oracle.set_vector_space(3, 1.0, 1.0)
    try:
        R = c_oracle.Reassembly(3, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tabd, nthreads=threads)
        t = time.perf_counter()
        R.step()
        dt = time.perf_counter() - t
    finally:
        c_oracle.set_vector_space(1)
    return {"value": R.nzval.size / dt, "unit": "nnz/s", "cores": threads, "kind": "port",
            "sample": f"one re-assembly of P2 x 3 elasticity on the simplexified {n}^3 mesh ({mesh.n_cells} tetrahedra, {R.nzval.size} nnz), "
                      f"{dt:.2f} s; phases(s) loop/compress!/vector={[round(float(x), 3) for x in R.phases]}; cell loop on {threads} pthreads"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    r = run(a.n, a.steps)
    r["cpu_baseline"] = cpu_baseline()
    print(json.dumps(r))
