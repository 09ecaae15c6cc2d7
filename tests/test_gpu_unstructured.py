"""Q1 hexahedra WITHOUT lattice structure (cells in a random order, nodes renumbered): the fused cell kernel
k_q1hex_cells + fixed-order reductions (fast-path id 5) against the oracle on the very same permuted inputs."""
import os

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, oracle_matrix, oracle_vector, problem, tab_dict

pytestmark = pytest.mark.gpu
E = gtk_b200.engine


@pytest.mark.parametrize("cells,bc,warp", [((6, 5, 4), "boundary", 0.2), ((4, 4, 4), None, 0.0), ((7, 3, 5), [1, 4], 0.15)])
def test_permuted_q1_hex_mesh(cells, bc, warp):
    mesh, V, tab = problem(cells, bc=bc, warp=warp)
    rng = np.random.default_rng(7)
    cperm = rng.permutation(mesh.n_cells)
    nperm = rng.permutation(mesh.n_nodes)                 # new id of old node k is nperm[k] + 1
    xyz = np.empty_like(mesh.node_coordinates)
    xyz[nperm] = mesh.node_coordinates
    cn = (nperm[mesh.cell_nodes.astype(np.int64) - 1] + 1).astype(np.int32)[cperm]
    cd = V.cell_dofs[cperm]
    tabd = tab_dict(tab)
    cp, rv, nz = O.assemble_matrix(O.LAPLACE, xyz, cn, cd, V.n_free, V.n_dirichlet, tabd, alpha=0.5)
    b_ref = O.assemble_vector(O.SOURCE_CONST, xyz, cn, cd, V.n_free, V.n_dirichlet, tabd, alpha=2.0, f_const=[1.5])
    eng = E.Engine(0)
    eng.set_mesh(xyz, cn); eng.set_space(cd, V.n_free, V.n_dirichlet); eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    eng.matrix_symbolic(); eng.vector_symbolic()
    cpg, rvg = eng.matrix_pattern()
    assert np.array_equal(cpg, cp) and np.array_equal(rvg, rv)
    got, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=0.5), E.FORM_SOURCE_CONST, dict(alpha=2.0, f_const=[1.5]))
    assert eng.info(5) == 5, "permuted Q1 hexahedra must take the fused cell kernel"
    assert_values_close(got, nz); assert_values_close(b, b_ref)
    got2, b2 = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=0.5), E.FORM_SOURCE_CONST, dict(alpha=2.0, f_const=[1.5]))
    assert got.tobytes() == got2.tobytes() and b.tobytes() == b2.tobytes()
    only = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.5)
    assert eng.info(5) == 5 and only.tobytes() == got.tobytes()
    os.environ["GTK_DISABLE_Q1CELLS"] = "1"
    try:
        ref2 = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.5)
        assert eng.info(5) != 5
    finally:
        del os.environ["GTK_DISABLE_Q1CELLS"]
    assert_values_close(only, ref2, tol=1e-13)
    # a subset of active cells (zeros elsewhere) and a coefficient field (not this kernel's business: generic path)
    eng.set_active_cells(3, mesh.n_cells // 2)
    part = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.5)
    os.environ["GTK_DISABLE_Q1CELLS"] = "1"
    try:
        assert_values_close(part, eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.5), tol=1e-13)
    finally:
        del os.environ["GTK_DISABLE_Q1CELLS"]
    eng.set_active_cells(0, mesh.n_cells)
    kq = 1.0 + rng.random((mesh.n_cells, 8))
    withk = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.5, coef_qp=kq)
    assert eng.info(5) != 5
    assert_values_close(withk, O.assemble_matrix(O.LAPLACE, xyz, cn, cd, V.n_free, V.n_dirichlet, tabd, alpha=0.5, coef_qp=kq)[2])
    eng.close()
