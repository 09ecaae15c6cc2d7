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


class _DevArr:
    """device pointer -> torch tensor (zero copy) through __cuda_array_interface__"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_invariants(eng, order, n, n_free):
    """Size-independent properties of the assembled Q_k Laplacian with full Dirichlet boundary, checked ON THE DEVICE at full
    size (the matrix of config 3 is 6.8 GB of values): nnz of the tensor-product pattern, finite values, positive diagonal,
    zero column sums away from the boundary (constants are in the kernel of the Laplacian), structural symmetry of the
    pattern and BITWISE symmetry of the values (the DMMA path mirrors the upper triangle of every element matrix and adds
    the contributions of a pair (i,j) and (j,i) in the same cell order)."""
    import torch
    m = order * n - 1
    pn, nn = eng.device_pointer(0)
    pc, nc = eng.device_pointer(2)
    pr, nr = eng.device_pointer(3)
    nz = torch.as_tensor(_DevArr(pn, nn, "<f8"), device="cuda")
    cp = torch.as_tensor(_DevArr(pc, nc, "<i8"), device="cuda")
    rv = torch.as_tensor(_DevArr(pr, nr, "<i4"), device="cuda")
    # nnz of the Q_k pattern with full Dirichlet boundary: per direction sum over free nodes of the 1D neighbour count
    per_dir = 0
    for i in range(1, order * n):              # free 1D nodes 1 .. kn-1
        if i % order == 0:                     # vertex node: two cells
            lo, hi = i - order, i + order
        else:                                  # cell-interior node: one cell
            lo, hi = (i // order) * order, (i // order) * order + order
        per_dir += min(hi, order * n - 1) - max(lo, 1) + 1
    res = {"n_free_ok": bool(n_free == m ** 3), "nnz": int(nn), "nnz_formula": int(per_dir ** 3), "nnz_ok": bool(nn == per_dir ** 3)}
    res["finite"] = bool(torch.isfinite(nz).all().item())
    lens = (cp[1:] - cp[:-1])
    col = torch.repeat_interleave(torch.arange(n_free, device="cuda", dtype=torch.int32), lens)
    diag = nz[(rv - 1) == col]
    res["diag_positive"] = bool(diag.numel() == n_free and (diag > 0).all().item())
    colsum = torch.segment_reduce(nz, "sum", lengths=lens)
    scale = float(nz.abs().max().item())
    nzero = int((colsum.abs() <= 1e-10 * scale).sum().item())
    res["zero_colsum_columns"] = nzero
    res["zero_colsum_ok"] = bool(nzero >= (order * (n - 2) - 1) ** 3)
    del diag, colsum
    # transpose by sorting the entries by (row, col): position t of the sorted order holds entry (col_t, row_t) of A^T
    key = (rv.to(torch.int64) - 1) * n_free + col.to(torch.int64)
    del col
    perm = torch.argsort(key)
    del key
    key_t = cp.new_empty(0)
    nzt = nz[perm]
    res["values_bitwise_symmetric"] = bool(torch.equal(nzt.view(torch.int64), nz.view(torch.int64)))
    res["values_symmetric_relerr"] = float(((nzt - nz).abs().max() / scale).item())
    del nzt
    rows_sorted = rv[perm]
    # structural symmetry: in transposed order the "column" ids (original rows) must reproduce colptr's run lengths
    cnt = torch.bincount((rows_sorted - 1).to(torch.int64), minlength=n_free)
    res["pattern_symmetric"] = bool(torch.equal(cnt, lens.to(cnt.dtype)))
    res["sum_nz"] = float(nz.sum().item())
    del perm, rows_sorted, cnt, key_t
    torch.cuda.empty_cache()
    return res


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
        # the kernel computes the upper triangle of 8x8 tiles only (Ke is symmetric): executed flops = exec_share of F_alg.  `frac` is on the
        # EXECUTED flops (what the tensor pipe actually does); the figure on F_alg is kept beside it
        "roofline": (lambda exec_share: {
            "bound": "tensor", "unit": "TFLOP/s", "algorithmic_flops": f_alg, "executed_flops": f_alg * exec_share,
            "executed_share": exec_share,
            "achieved": f_alg * exec_share / (gemm_ms * 1e-3) / 1e12, "peak": DMMA_PEAK_TFLOPS,
            "frac": f_alg * exec_share / (gemm_ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
            "frac_on_algorithmic_flops": f_alg / (gemm_ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
            "kernel": "k_elem_laplace_dmma", "kernel_ms": gemm_ms,
            "whole_step_frac": f_alg * exec_share / (ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
            "whole_step_frac_on_algorithmic_flops": f_alg / (ms * 1e-3) / 1e12 / DMMA_PEAK_TFLOPS,
            "peak_source": "measured FP64 DMMA.8x8x4 issue rate on B200 (tools/dmma_peak.cu)"})(
                (lambda t: t * (t + 1) / 2 / (t * t))(max(1, V.cell_dofs.shape[1] // 8))),
        "device_bytes": int(eng.info(2)),
    }
    if check:
        out["checks"] = device_invariants(eng, order, n, V.n_free)
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
