"""free + Dirichlet column blocks side by side (matrix slots) and the linear-problem right-hand side
b <- b - Ad*xd (problems.jl:363-387, 439-453) on the device, against the oracle."""
import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, make_engine, problem, tab_dict

E = gtk_b200.engine
GT = gtk_b200.gt
pytestmark = pytest.mark.gpu

CASES = [((9, 7), 1, "boundary", 0.2), ((6, 5, 4), 1, [1, 4], 0.15), ((12, 10, 8), 1, "boundary", 0.0), ((4, 3, 3), 2, [2, 5], 0.1)]


@pytest.mark.parametrize("cells,order,bc,warp", CASES)
def test_slots_and_matvec_add(cells, order, bc, warp):
    mesh, V, tab = problem(cells, order=order, bc=bc, warp=warp)
    rng = np.random.default_rng(11)
    xd = rng.standard_normal(V.n_dirichlet)
    (cpA, rvA, nzA), (cpD, rvD, nzD), b_ref = O.linear_problem_rhs(
        O.LAPLACE, O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tab_dict(tab), xd,
        params_a=dict(alpha=1.0), params_l=dict(f_const=[1.0]))
    eng = make_engine(mesh, V, tab)
    eng.select_matrix(0)
    assert eng.matrix_symbolic(E.FREE, E.FREE) == rvA.size
    nz_a, b0 = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    eng.select_matrix(1)
    assert eng.matrix_symbolic(E.FREE, E.DIRICHLET) == rvD.size
    cp_d, rv_d = eng.matrix_pattern()
    assert np.array_equal(cp_d, cpD) and np.array_equal(rv_d, rvD)
    nz_d = eng.matrix_numeric(E.FORM_LAPLACE, alpha=1.0)
    assert_values_close(nz_d, nzD)
    b = eng.matvec_add(-1.0, xd, 1.0)
    assert_values_close(b, b_ref)
    # bitwise: Julia's mul! order and roundings applied to the engine's own Ad and b
    assert b.tobytes() == O.spmatmul_add(cp_d, rv_d, nz_d, xd, -1.0, 1.0, b0).tobytes()
    # general alpha/beta (beta == 0 zero-fills, other beta scales first)
    b1 = eng.matvec_add(0.5, xd, 0.0)
    assert b1.tobytes() == O.spmatmul_add(cp_d, rv_d, nz_d, xd, 0.5, 0.0, b).tobytes()
    b2 = eng.matvec_add(2.0, xd, -0.25)
    assert b2.tobytes() == O.spmatmul_add(cp_d, rv_d, nz_d, xd, 2.0, -0.25, b1).tobytes()
    # slot 0 is untouched: pattern and values still there, update_matrix! stays numeric-only and bit-identical
    eng.select_matrix(0)
    cp_a, rv_a = eng.matrix_pattern()
    assert np.array_equal(cp_a, cpA) and np.array_equal(rv_a, rvA)
    assert eng.copy_nzval().tobytes() == nz_a.tobytes()
    assert eng.matrix_numeric(E.FORM_LAPLACE, alpha=1.0).tobytes() == nz_a.tobytes()
    # A*x through the same call on the free x free block
    x = rng.standard_normal(V.n_free)
    y = eng.matvec_add(1.0, x, 0.0)
    assert y.tobytes() == O.spmatmul_add(cp_a, rv_a, nz_a, x, 1.0, 0.0, b2).tobytes()
    eng.close()


def test_linear_problem_mirror_solves_manufactured_poisson():
    """GT-style driver: u = x + 2y + 3z is harmonic, so with its boundary values as Dirichlet data and f = 0 the
    discrete solution is exact (Q1 reproduces linears): A x = b - Ad xd gives the nodal values."""
    import scipy.sparse.linalg as spla
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (6, 5, 4))
    Om = GT.interior(mesh)
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=GT.boundary(mesh))
    dOm = GT.measure(Om, 2)
    g = lambda X: X[..., 0] + 2 * X[..., 1] + 3 * X[..., 2]
    xd = g(V.data.dirichlet_dof_nodes)
    a = lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dOm)
    l = lambda v: GT.integrate(lambda x: v(x) * 0.0, dOm)
    x0, A, b = GT.linear_problem(xd, a, l, V)
    assert x0.shape == (V.num_free_dofs(),) and not x0.any()
    x = spla.spsolve(A.to_scipy().tocsc(), b)
    assert np.abs(x - g(V.data.free_dof_nodes)).max() < 1e-12
