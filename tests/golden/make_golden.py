#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from oracle/gt_oracle.py (run from the repo root: python tests/golden/make_golden.py).

These fixtures freeze the ORACLE's output (inputs + colptr/rowval/nzval/b) on small cases; they are NOT outputs of the
reference itself (pure Julia, not runnable in this image) — see DESIGN.md §6 "parity unpinned".  If a Julia-equipped box
appears, the same cases can be dumped from GalerkinToolkit (face_dofs(V).data, A.colptr/rowval/nzval, b) and diffed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import gt_oracle as O  # noqa: E402
from util import oracle_matrix, oracle_vector, problem  # noqa: E402

CASES = {
    # name: (cells, order, simplexify, n_comp, bc, warp, form, params)
    "q1_2d_8x8_laplace": ((8, 8), 1, False, 1, "boundary", 0.0, O.LAPLACE, {}),            # config 1 element
    "q1_3d_4x3x3_warped_laplace": ((4, 3, 3), 1, False, 1, "boundary", 0.2, O.LAPLACE, {}),  # config 2 element, non-affine
    "q1_3d_5x4x3_nobc_laplace": ((5, 4, 3), 1, False, 1, None, 0.0, O.LAPLACE, {}),          # corner-first numbering
    "q2_2d_3x2_mass": ((3, 2), 2, False, 1, [1, 4], 0.1, O.MASS, {}),
    "p2x3_tets_2x2x1_elasticity": ((2, 2, 1), 2, True, 3, [1], 0.1, O.ELASTICITY, dict(lam=1.0, mu=1.0)),  # config 4 element
}


def build(name):
    cells, order, simp, n_comp, bc, warp, form, params = CASES[name]
    mesh, V, tab = problem(cells, order=order, bc=bc, n_comp=n_comp, simplexify=simp, warp=warp)
    colptr, rowval, nzval = oracle_matrix(form, mesh, V, tab, alpha=1.0, **params)
    f = [1.0, -2.0, 0.5][:n_comp]
    b = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=f)
    return dict(xyz=mesh.node_coordinates, cell_nodes=mesh.cell_nodes, cell_dofs=V.cell_dofs, n_free=V.n_free,
                n_dirichlet=V.n_dirichlet, n_comp=n_comp, w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM, form=form,
                lam=params.get("lam", 0.0), mu=params.get("mu", 0.0), f_const=np.array(f), colptr=colptr, rowval=rowval,
                nzval=nzval, b=b)


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
