"""SURVEY.md §8 f4 on the device: Stokes (Q2 x 3 velocity x Q1 pressure, docs/src/src_jl/example_stokes.jl) on an n^3
hexahedral mesh and the skeleton term ∫_Λ jump(u) jump(v) (test/assembly_tests.jl:397-401) on m^3 Q1 cells — numeric
re-assembly on a cached pattern through the block kernels.  Prints one JSON object (imported by bench.py: `multifield`).

    python tools/bench_multifield.py [--n 24] [--m 48] [--steps 3]
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT,):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def _time(eng, blocks, steps, warmup):
    eng.set_profiling(True)
    times, kern = [], {}
    for it in range(warmup + steps):
        eng.matrix_numeric_blocks_device(blocks)
        recs = eng.profile()
        if it >= warmup:
            times.append(sum(ms for _, ms in recs))
            for name, ms in recs:
                kern[name] = kern.get(name, 0.0) + ms / steps
    eng.set_profiling(False)
    return sum(times) / len(times), kern


def _engine(E, bp, device):
    eng = E.Engine(device)
    eng.set_mesh(bp.node_coordinates, bp.face_nodes)
    if bp.manifold_dim != bp.node_coordinates.shape[1]:
        eng.set_manifold_dim(bp.manifold_dim)
    eng.set_space(bp.super_dofs, bp.n_free, bp.n_dirichlet, 1)
    eng.set_parts(bp.w, bp.M, bp.dM, bp.parts, bp.n_sides, bp.face_var)
    return eng


def run(n=24, m=48, steps=3, warmup=1, device=0):
    import numpy as np
    import gtk_b200
    H, E = gtk_b200.hostprep, gtk_b200.engine
    MF = importlib.import_module("galerkintoolkit_jl_b200.multifield")
    out = {}
    # ---- Stokes ----
    t0 = time.perf_counter()
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
    V = H.lagrange_space(mesh, 2, "boundary", 3)
    Q = H.lagrange_space(mesh, 1, None, 1)
    bp = MF.volume_problem([V, Q], 4)
    host_s = time.perf_counter() - t0
    eng = _engine(E, bp, device)
    t0 = time.perf_counter()
    nnz = eng.matrix_symbolic()
    symbolic_ms = 1e3 * (time.perf_counter() - t0)
    blocks = [(0, 0, E.BLOCK_LAPLACE, 1.0), (1, 0, E.BLOCK_VALU_DIVV, -1.0), (0, 1, E.BLOCK_DIVU_VALV, 1.0)]
    ms, kern = _time(eng, blocks, steps, warmup)
    nz = eng.copy_nzval()
    cp, rv = eng.matrix_pattern()
    col = np.repeat(np.arange(bp.n_free), np.diff(cp.astype(np.int64)))
    pp = (col >= V.n_free) & (rv - 1 >= V.n_free)          # the (p, q) block: stored, identically zero
    out["stokes"] = {
        "workload": f"Stokes a((u,p),(v,q)) = ∫ ∇v⋅∇u - div(v) p + q div(u), Q2 x 3 velocity x Q1 pressure on {n}^3 hexahedra, 3^3 Gauss points, "
                    f"velocity Dirichlet on the whole boundary, monolithic matrix (all four field blocks stored), numeric re-assembly",
        "cells": int(mesh.n_cells), "super_element_dofs": int(bp.super_dofs.shape[1]), "free_dofs": int(bp.n_free), "nnz": int(nnz),
        "n_coo": int(eng.info(4)), "fast_path": int(eng.info(5)), "ms_per_step": ms, "nnz_per_s": nnz / (ms * 1e-3), "kernels_ms": kern,
        "symbolic_ms": symbolic_ms, "host_prep_s": host_s,
        "checks": {"finite": bool(np.isfinite(nz).all()), "pressure_block_entries_stored": int(pp.sum()),
                   "pressure_block_zero": bool(np.all(nz[pp] == 0.0))},
    }
    eng.close()
    # ---- skeleton ----
    t0 = time.perf_counter()
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (m, m, m))
    W = H.lagrange_space(mesh, 1, "boundary", 1)
    bs = MF.skeleton_problem([W], 2)
    host_s = time.perf_counter() - t0
    eng = _engine(E, bs, device)
    t0 = time.perf_counter()
    nnz = eng.matrix_symbolic()
    symbolic_ms = 1e3 * (time.perf_counter() - t0)
    blocks = [(0, 0, E.BLOCK_MASS, 1.0), (0, 1, E.BLOCK_MASS, -1.0), (1, 0, E.BLOCK_MASS, -1.0), (1, 1, E.BLOCK_MASS, 1.0)]
    ms, kern = _time(eng, blocks, steps, warmup)
    nz = eng.copy_nzval()
    out["skeleton"] = {
        "workload": f"∫_Λ jump(u) jump(v) dΛ over the {bs.face_nodes.shape[0]} interior faces of {m}^3 Q1 hexahedra (two cells around each face, "
                    f"2 x 2 Gauss points per face), numeric re-assembly",
        "faces": int(bs.face_nodes.shape[0]), "super_element_dofs": int(bs.super_dofs.shape[1]), "free_dofs": int(bs.n_free), "nnz": int(nnz),
        "n_coo": int(eng.info(4)), "ms_per_step": ms, "nnz_per_s": nnz / (ms * 1e-3), "kernels_ms": kern, "symbolic_ms": symbolic_ms,
        "host_prep_s": host_s, "checks": {"max_abs_value": float(np.abs(nz).max()), "continuous_space_has_no_jump": bool(np.abs(nz).max() < 1e-12)},
    }
    eng.close()
    try:
        out["interior_penalty"] = run_dg(24, steps, device)
    except Exception as exc:   # reported, never hidden
        out["interior_penalty"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def run_dg(n=24, steps=3, device=0):
    """docs/src/src_jl/example_hello_world_dg.jl at n^3 cells through the GT mirror: discontinuous Q1 space, Laplace operator on Ω +
    interior penalty on Λ + Nitsche terms on Γ merged into ONE matrix (gtk_matrix_sum_*); timed: update_matrix! (the three
    integrals re-assembled on the device, merged, values copied to the host)."""
    import numpy as np
    import gtk_b200
    GT = gtk_b200.gt
    GT.set_device(device)
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
    nrm = GT.unit_normal(mesh, 2)
    Om, Gd, Lam = GT.interior(mesh), GT.boundary(mesh), GT.skeleton(mesh)
    h_L, h_G = GT.face_diameter_field(Lam), GT.face_diameter_field(Gd)
    mean = lambda fn, u, x: 0.5 * (fn(u[1], x) + fn(u[2], x))
    jump = lambda u, n_, x: u[2](x) * n_[2](x) + u[1](x) * n_[1](x)
    gamma = 0.2
    t0 = time.perf_counter()
    V = GT.lagrange_space(Om, 1, continuous=False)
    dO, dL, dG = GT.measure(Om, 2), GT.measure(Lam, 2), GT.measure(Gd, 2)
    grad, dot = GT.grad, GT.dot
    a = lambda u, v: (GT.integrate(lambda x: dot(grad(u, x), grad(v, x)), dO)
                      + GT.integrate(lambda x: dot((gamma / h_L(x)) * jump(v, nrm, x), jump(u, nrm, x)) - dot(jump(v, nrm, x), mean(grad, u, x))
                                     - dot(mean(grad, v, x), jump(u, nrm, x)), dL)
                      + GT.integrate(lambda x: (gamma / h_G(x)) * v(x) * u(x) - dot(v(x) * nrm(x), grad(u, x)) - dot(nrm(x), grad(v, x)) * u(x), dG))
    A, cache = GT.assemble_matrix(a, float, V, V, reuse=True)
    first_s = time.perf_counter() - t0
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        GT.update_matrix(A, cache)
        times.append(time.perf_counter() - t0)
    S = A.to_scipy()
    sym = float(abs(S - S.T).max() / abs(S).max())
    for e, _ in cache.params["parts"]:
        e.close()
    cache.engine.close()
    ms = 1e3 * min(times)
    return {"workload": f"symmetric interior penalty (docs/src/src_jl/example_hello_world_dg.jl) on {n}^3 hexahedra, discontinuous Q1 space: "
                        f"∫_Ω ∇u⋅∇v + interior-penalty terms on the skeleton + Nitsche terms on the boundary, merged into one matrix on the device",
            "free_dofs": int(A.m), "nnz": int(A.nzval.size), "update_matrix_ms_incl_copy_out": ms, "nnz_per_s": A.nzval.size / (ms * 1e-3),
            "first_assembly_s_incl_host_prep": first_s, "checks": {"symmetric_relerr": sym, "finite": bool(np.isfinite(A.nzval).all())}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--m", type=int, default=48)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    print(json.dumps(run(a.n, a.m, a.steps)))
