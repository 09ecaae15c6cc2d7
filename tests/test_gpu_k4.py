"""K4: the warp-per-cell isotropic-elasticity kernel for 3D vector elements with <= 10 scalar shape functions (P1 / P2
tetrahedra, BASELINE config 4) against the oracle and against the generic per-entry kernel (1e-14)."""
import os

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, make_engine, oracle_matrix, problem

pytestmark = pytest.mark.gpu
E = gtk_b200.engine


@pytest.mark.parametrize("cells,order,bc,warp", [((3, 2, 2), 2, [1], 0.15), ((4, 3, 3), 1, "boundary", 0.2), ((2, 2, 3), 2, None, 0.0),
                                                 ((5, 4, 3), 2, [2, 5], 0.1)])
def test_k4_matches_oracle_and_generic_kernel(cells, order, bc, warp):
    mesh, V, tab = problem(cells, order=order, bc=bc, n_comp=3, simplexify=True, warp=warp)
    assert tab.w.size == (11 if order == 2 else 4)                 # Strang degree 4 / 2
    cp, rv, nz = oracle_matrix(O.ELASTICITY, mesh, V, tab, alpha=0.75, lam=1.3, mu=0.7)
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    cpg, rvg = eng.matrix_pattern()
    assert np.array_equal(cpg, cp) and np.array_equal(rvg, rv)
    got = eng.matrix_numeric(E.FORM_ELASTICITY_ISO, alpha=0.75, lam=1.3, mu=0.7)
    assert eng.info(5) == 4, "K4 must take P1/P2 tetrahedra with 3 components"
    assert_values_close(got, nz)
    assert got.tobytes() == eng.matrix_numeric(E.FORM_ELASTICITY_ISO, alpha=0.75, lam=1.3, mu=0.7).tobytes()
    os.environ["GTK_DISABLE_K4"] = "1"
    try:
        generic = eng.matrix_numeric(E.FORM_ELASTICITY_ISO, alpha=0.75, lam=1.3, mu=0.7)
        assert eng.info(5) == 0
    finally:
        del os.environ["GTK_DISABLE_K4"]
    assert_values_close(got, generic, tol=1e-13)                    # same closed forms and q order; (α dV) folded in earlier
    # numeric-active subset (multi-GPU halo semantics): inactive cells contribute zeros
    nc = mesh.n_cells
    eng.set_active_cells(nc // 3, nc // 2)
    part = eng.matrix_numeric(E.FORM_ELASTICITY_ISO, alpha=0.75, lam=1.3, mu=0.7)
    os.environ["GTK_DISABLE_K4"] = "1"
    try:
        assert_values_close(part, eng.matrix_numeric(E.FORM_ELASTICITY_ISO, alpha=0.75, lam=1.3, mu=0.7), tol=1e-13)
    finally:
        del os.environ["GTK_DISABLE_K4"]
    eng.close()
