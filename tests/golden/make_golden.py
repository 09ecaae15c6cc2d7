#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from oracle/gt_oracle.py (run from the repo root: python tests/golden/make_golden.py).

These fixtures freeze the ORACLE's output (inputs + colptr/rowval/nzval/b) on small cases; they are NOT outputs of the
reference itself (pure Julia, not runnable in this image) — see DESIGN.md §6 (the oracle itself is pinned by the reference's known answers, not by these files).  If a Julia-equipped box
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
    "p2x3_tets_2x2x2_elasticity": ((2, 2, 2), 2, True, 3, [1], 0.1, O.ELASTICITY, dict(lam=1.0, mu=1.0)),  # config 4 element
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


def build_linear_problem():
    """A whole linear problem (problems.jl:439-453): variable-coefficient Laplacian, volume source + Neumann flux on two
    sides, Dirichlet data on one side: A, Ad, b = l - Ad*xd.  Lives in golden/extra/ (its own schema)."""
    import gtk_b200
    H = gtk_b200.hostprep
    mesh, V, tab = problem((4, 3, 2), order=1, bc=[1], warp=0.15)
    rng = np.random.default_rng(42)
    kn = 1.0 + rng.random(mesh.n_nodes)
    xd = rng.standard_normal(V.n_dirichlet)
    fp = H.face_problem(V, [2, 6], 2)
    gq = rng.standard_normal((fp.face_nodes.shape[0], fp.tab.w.size, 1))
    tabd = dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)
    tabf = dict(w=fp.tab.w, N=fp.tab.N, dN=fp.tab.dN, M=fp.tab.M, dM=fp.tab.dM)
    base = (mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd)
    A = O.assemble_matrix(O.LAPLACE, *base, coef_nodal=kn)
    Ad = O.assemble_matrix(O.LAPLACE, *base, coef_nodal=kn, free_or_dirichlet=(O.FREE, O.DIRICHLET))
    b_vol = O.assemble_vector(O.SOURCE_CONST, *base, f_const=[1.0])
    b_l = O.assemble_vector(O.SOURCE_QP, mesh.node_coordinates, fp.face_nodes, fp.face_dofs, V.n_free, V.n_dirichlet, tabf,
                            f_qp=gq, b0=b_vol)
    b = O.spmatmul_add(Ad[0], Ad[1], Ad[2], xd, -1.0, 1.0, b_l)
    out = dict(xyz=mesh.node_coordinates, cell_nodes=mesh.cell_nodes, cell_dofs=V.cell_dofs, n_free=V.n_free,
               n_dirichlet=V.n_dirichlet, coef_nodal=kn, xd=xd, face_nodes=fp.face_nodes, face_dofs=fp.face_dofs, g_qp=gq,
               A_colptr=A[0], A_rowval=A[1], A_nzval=A[2], Ad_colptr=Ad[0], Ad_rowval=Ad[1], Ad_nzval=Ad[2],
               b_vol=b_vol, b_l=b_l, b=b)
    for k, v in tabd.items():
        out["cell_" + k] = v
    for k, v in tabf.items():
        out["face_" + k] = v
    return out


if __name__ == "__main__":
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
    os.makedirs(os.path.join(HERE, "extra"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "extra", "linear_problem_q1_3d_4x3x2.npz"), **build_linear_problem())
    print("wrote extra/linear_problem_q1_3d_4x3x2")
