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


def build_f4():
    """SURVEY §8 f4 fixtures (golden/extra/, own schema): Stokes Q2 x Q1 on a warped 3 x 2 quad mesh (monolithic matrix with all
    four field blocks) and the interior-penalty operator of docs/src/src_jl/example_hello_world_dg.jl on a warped 3 x 3 mesh
    (discontinuous Q1: Laplace + skeleton + Nitsche triplets compressed once, Nitsche right-hand side for g = x + 2y)."""
    import importlib
    import gtk_b200
    import test_multifield as T
    H = gtk_b200.hostprep
    MF = importlib.import_module("galerkintoolkit_jl_b200.multifield")
    out = {}
    mesh, spaces = T._stokes_spaces((3, 2))
    bp = MF.volume_problem(spaces, 4)
    cp, rv, nz = T._oracle_matrix(bp, spaces, mesh, "stokes")
    out["f4_stokes_q2q1_2d_3x2"] = dict(xyz=mesh.node_coordinates, cell_nodes=mesh.cell_nodes, super_dofs=bp.super_dofs, n_free=bp.n_free,
                                        n_dirichlet=bp.n_dirichlet, colptr=cp, rowval=rv, nzval=nz)
    mesh = H.cartesian_mesh((0, 1, 0, 1), (3, 3))
    T._warp(mesh)
    V = H.discontinuous_lagrange_space(mesh, 1)
    gamma = 0.2
    coo = [T._volume_coo(O.LAPLACE, mesh, V, 2)]
    bs = MF.skeleton_problem([V], 2, gradients=True)
    sides = [[(int(bs.side_cells[i, a]), int(bs.face_var[i, a])) for a in range(2)] for i in range(bs.face_nodes.shape[0])]
    full = lambda p: sum(T._ip(p, su, sv, (gamma, -0.5, -0.5)) for su in (1, 2) for sv in (1, 2))
    coo.append(O.assemble_matrix_multifield(2, mesh.node_coordinates, bs.face_nodes, dict(w=bs.w, dM=bs.dM), sides, T._oracle_fields(bs, [V], True), full,
                                            skeleton_geometry=(bs.cell_nodes, bs.dM_cell, bs.ref_normals), return_coo=True))
    bb = MF.boundary_problem([V], None, 2)
    bsides = T._boundary_oracle_sides(bb)
    geo = (bb.cell_nodes, bb.dM_cell, bb.ref_normals)
    nit = lambda p: (gamma / p.h) * p.v(0) * p.u(0) - O.frobenius(p.v(0) * p.n(1), p.grad_u(0)) - O.frobenius(p.n(1), p.grad_v(0)) * p.u(0)
    coo.append(O.assemble_matrix_multifield(2, mesh.node_coordinates, bb.face_nodes, dict(w=bb.w, dM=bb.dM), bsides, T._oracle_fields(bb, [V], True), nit,
                                            skeleton_geometry=geo, return_coo=True))
    cp, rv, nz = O.assemble_matrix_sum(coo, V.n_free, V.n_free)
    xq = MF.face_point_coordinates(bb)
    g = xq[..., 0] + 2.0 * xq[..., 1]
    rhs = lambda p: (gamma / p.h) * p.v(0) * p.g - O.frobenius(p.n(1), p.grad_v(0)) * p.g
    b = O.assemble_vector_multifield(2, mesh.node_coordinates, bb.face_nodes, dict(w=bb.w, dM=bb.dM), bsides, T._oracle_fields(bb, [V], True), rhs,
                                     skeleton_geometry=geo, point_data=g)
    out["f4_interior_penalty_dg_2d_3x3"] = dict(xyz=mesh.node_coordinates, cell_nodes=mesh.cell_nodes, cell_dofs=V.cell_dofs, gamma=gamma,
                                                colptr=cp, rowval=rv, nzval=nz, b=b)
    return out


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "extra"), exist_ok=True)
    for name, data in build_f4().items():
        np.savez_compressed(os.path.join(HERE, "extra", name + ".npz"), **data)
        print("wrote extra/" + name)
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **build(name))
        print("wrote", name)
    os.makedirs(os.path.join(HERE, "extra"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "extra", "linear_problem_q1_3d_4x3x2.npz"), **build_linear_problem())
    print("wrote extra/linear_problem_q1_3d_4x3x2")
