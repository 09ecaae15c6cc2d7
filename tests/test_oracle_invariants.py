"""Pin the oracle with the reference's own known-answer invariants.
The reference holds no golden matrices; these are the checks its test-suite makes."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import gt_oracle as O
from util import oracle_matrix, oracle_vector, problem, tab_dict


def _csc(colptr, rowval, nzval, m, n):
    return sp.csc_matrix((nzval, rowval.astype(np.int64) - 1, colptr.astype(np.int64) - 1), shape=(m, n))


def test_mass_and_source_sums_2x2_square():
    # test/problems_tests.jl:48-57: Q1 on a 2x2 unit square: sum(M) ≈ 1, sum(b) ≈ 1
    mesh, V, tab = problem((2, 2), bc=None)
    cp, rv, nz = oracle_matrix(O.MASS, mesh, V, tab)
    assert abs(nz.sum() - 1.0) < 1e-14
    b = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[1.0])
    assert abs(b.sum() - 1.0) < 1e-14


@pytest.mark.parametrize("cells,domain", [((4, 3), (0, 2, 0, 1)), ((3, 4, 2), (0, 1, 0, 2, 0, 3))])
def test_mass_sum_is_volume_and_laplace_rowsums_vanish(cells, domain):
    mesh, V, tab = problem(cells, bc=None, warp=0.2, domain=domain)
    vol = np.prod([domain[2 * d + 1] - domain[2 * d] for d in range(len(cells))])
    cp, rv, nz = oracle_matrix(O.MASS, mesh, V, tab)
    assert abs(nz.sum() - vol) < 1e-12 * vol
    cp, rv, nz = oracle_matrix(O.LAPLACE, mesh, V, tab)
    A = _csc(cp, rv, nz, V.n_free, V.n_free)
    assert np.abs(A.sum(axis=1)).max() < 1e-12 * np.abs(nz).max()
    assert abs(A - A.T).max() == 0.0          # dot() is bitwise commutative ⇒ Ke bitwise symmetric (A.6)


def test_quadrature_weights_sum_and_tabulator_identity():
    # test/integration_tests.jl:28-34 ; test/space_tests.jl:218-220, 250-252
    for D in (1, 2, 3):
        for deg in (1, 2, 4, 6):
            x, w = O.tensor_gauss(D, deg)
            assert abs(w.sum() - 1.0) < 1e-14
    for D, kind in ((2, "Q"), (3, "Q"), (2, "P"), (3, "P")):
        for order in (1, 2, 3):
            nodes = [[e[d] / order for d in range(D)] for e in O.monomial_exponents(D, order, kind)]
            N, dN = O.tabulate(D, order, kind, nodes)
            assert np.abs(N - np.eye(len(nodes))).max() < 1e-10
            assert np.abs(dN.sum(axis=1)).max() < 1e-9      # partition of unity


def test_dof_counts():
    # 2x2x2 Q1 cube with full Dirichlet boundary: 1 free, 26 Dirichlet
    o = O.q1_space((0, 1, 0, 1, 0, 1), (2, 2, 2), "boundary")
    assert (o["n_free"], o["n_dirichlet"]) == (1, 26)
    # corners are numbered first (topology.jl:1072-1078): without BC node 1 is dof 1, the far corner dof 8
    o = O.q1_space((0, 1, 0, 1, 0, 1), (2, 2, 2), None)
    assert o["cell_dofs"][0, 0] == 1 and o["cell_dofs"][-1, -1] == 8
    assert o["cell_dofs"].max() == 27


def test_pattern_counts_match_closed_form():
    # nnz = (3m-2)^D with m free nodes per direction (Q1, full Dirichlet boundary)
    for cells in ((8, 8), (6, 6, 6)):
        mesh, V, tab = problem(cells, bc="boundary")
        cp, rv, nz = oracle_matrix(O.LAPLACE, mesh, V, tab)
        m = cells[0] - 1
        assert rv.size == (3 * m - 2) ** len(cells)
        assert cp[-1] - 1 == rv.size
        # rows sorted inside each column, explicit zeros kept
        for j in range(V.n_free):
            seg = rv[cp[j] - 1: cp[j + 1] - 1]
            assert np.all(np.diff(seg) > 0)


def test_manufactured_poisson_solution():
    """test/problems_tests.jl:86-105: u = x+y is reproduced exactly by Q1 (el2 ≈ 0) with
    b = -Ad*xd (problems.jl:439-453)."""
    mesh, V, tab = problem((6, 5), bc="boundary", warp=0.15)
    cp, rv, nz = oracle_matrix(O.LAPLACE, mesh, V, tab)
    A = _csc(cp, rv, nz, V.n_free, V.n_free)
    cpd, rvd, nzd = oracle_matrix(O.LAPLACE, mesh, V, tab, fd=(O.FREE, O.DIRICHLET))
    Ad = _csc(cpd, rvd, nzd, V.n_free, V.n_dirichlet)
    u = lambda x: x[:, 0] + x[:, 1]
    xd = u(V.dirichlet_dof_nodes)
    x = spla.spsolve(A.tocsc(), -Ad @ xd)
    assert np.abs(x - u(V.free_dof_nodes)).max() < 1e-12


def test_sparse_combines_duplicates_in_input_order_and_keeps_zeros():
    I = np.array([2, 1, 2, 2, 1], dtype=np.int32)
    J = np.array([1, 1, 1, 1, 2], dtype=np.int32)
    V = np.array([1e16, 3.0, 1.0, -1e16, 0.0])
    cp, rv, nz = O.sparse_csc(I, J, V, 2, 2)
    assert cp.tolist() == [1, 3, 4] and rv.tolist() == [1, 2, 1]
    assert nz.tolist() == [3.0, (1e16 + 1.0) - 1e16, 0.0]      # left-to-right; explicit zero stays
    b = O.dense_vector(np.array([2, 2, 1], dtype=np.int32), np.array([1e16, 1.0, 5.0]), 3)
    assert b.tolist() == [5.0, 1e16 + 1.0, 0.0]


@pytest.mark.parametrize("cells,order,simplexify,n_comp", [((3, 2), 2, False, 1), ((2, 2, 2), 2, False, 1), ((2, 2, 2), 3, False, 1),
                                                           ((3, 3), 2, True, 1), ((2, 2, 2), 2, True, 3)])
def test_high_order_invariants(cells, order, simplexify, n_comp):
    """sum(M) = n_comp |Omega| (problems_tests.jl:56-57) and Laplacian row sums vanish without BCs, for the elements of
    BASELINE configs 3/4 with the reference's face-complex dof numbering (hostprep -> refnumbering.py)."""
    import scipy.sparse as sp
    from util import oracle_matrix, problem
    mesh, V, tab = problem(cells, order=order, bc=None, n_comp=n_comp, simplexify=simplexify)
    cp, rv, nz = oracle_matrix(O.MASS, mesh, V, tab)
    assert abs(nz.sum() - n_comp) < 1e-10
    cp, rv, nz = oracle_matrix(O.LAPLACE, mesh, V, tab)
    A = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
    assert abs(A - A.T).max() == 0.0
    assert np.abs(np.asarray(A.sum(axis=1))).max() < 1e-10 * np.abs(nz).max()
    # every dof is shared consistently: number of dofs of the conforming space
    if not simplexify:
        assert V.n_free == n_comp * int(np.prod([order * c + 1 for c in cells]))


@pytest.mark.parametrize("cells,order", [((3, 2, 2), 2), ((2, 2, 2), 3), ((4, 3), 3)])
def test_c_oracle_matches_numpy_oracle_on_high_order_elements(cells, order):
    """The C restatement is the timed CPU baseline of bench.py's high_order entry (config 3 element): same pattern and
    values as the numpy oracle on the reference's own high-order dof numbering."""
    import c_oracle
    from util import oracle_matrix, oracle_vector, problem, tab_dict, assert_values_close
    mesh, V, tab = problem(cells, order=order, bc=[1, 4], warp=0.1)
    cp, rv, nz = oracle_matrix(O.LAPLACE, mesh, V, tab)
    b = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[1.0])
    out = c_oracle.assemble(1, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tab_dict(tab), nthreads=2,
                            nnz_cap=mesh.n_cells * V.cell_dofs.shape[1] ** 2)
    assert np.array_equal(out[0], cp) and np.array_equal(out[1], rv)
    assert_values_close(out[2], nz)
    assert_values_close(out[3], b)


def test_oracle_reproduces_the_reference_plaplacian_golden():
    """The ONE numeric known-answer the reference's tests hold on the assembly path: the L2 norm of the discrete
    p-Laplacian solution, 0.09133166701839236 (test/problems_ext_tests.jl:148-172; same forms in
    test/assembly_tests.jl:675-701).  It depends on the whole oracle chain — cartesian_mesh coordinates, Dirichlet
    partition, Q1 tabulation, Gauss rule, J / dV / Jᵀ\\∇̂N, DiscreteField gradients, element vectors and matrices,
    COO → CSC compression and the scalar integral — so reproducing it to 1e-10 (the reference's own tolerance; we get
    ~1e-17) pins the oracle against a number the reference holds."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    mesh, V, tab = problem((10, 10), order=1, bc="boundary", domain=(0, 1, 0, 1))
    td = tab_dict(tab)
    X, CN, CD = mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs
    x = np.random.default_rng(0).random(V.n_free)           # GT.rand_field: random free values, zero Dirichlet values
    xd = np.zeros(V.n_dirichlet)
    q = 3
    R = lambda x: O.assemble_vector_plaplace_residual(X, CN, CD, V.n_free, td, x, xd, q)
    for it in range(60):
        r = R(x)
        nr = np.linalg.norm(r)
        if nr < 1e-13:
            break
        cp, rv, nz = O.assemble_matrix_plaplace_jacobian(X, CN, CD, V.n_free, td, x, xd, q)
        dx = spl.spsolve(sp.csc_matrix((nz, rv - 1, cp - 1), shape=(V.n_free, V.n_free)), -r)
        t = 1.0
        while np.linalg.norm(R(x + t * dx)) >= (1 - 1e-4 * t) * nr and t > 1e-8:
            t *= 0.5
        x = x + t * dx
    else:
        raise AssertionError("Newton did not converge")
    uhl2 = np.sqrt(O.assemble_scalar_field(O.SCALAR_L2SQ, X, CN, CD, td, x, xd))
    assert abs(uhl2 - 0.09133166701839236) < 1.0e-10, uhl2


def test_strang_tet_rules_match_the_literal_tables_and_are_exact():
    """quadrature.jl:41-48, 500-635: tetrahedra of degree 1..5 use the Strang-Fix rules (degree 4 = the 11-point rule of
    BASELINE config 4, negative centroid weight).  The product builds them from their orbits, the oracle restates the
    tables literally: equal bit for bit, in the reference's point order, and exact for all monomials up to the degree."""
    import itertools
    import math
    import gtk_b200
    H = gtk_b200.hostprep
    for degree in range(1, 6):
        q = H.quadrature(3, True, degree)
        xo, wo = O.strang_tet(degree)
        assert q.coordinates.tobytes() == xo.tobytes() and q.weights.tobytes() == wo.tobytes()
        for i, j, k in itertools.product(range(degree + 1), repeat=3):
            if i + j + k <= degree:
                exact = math.factorial(i) * math.factorial(j) * math.factorial(k) / math.factorial(i + j + k + 3)
                got = (wo * xo[:, 0] ** i * xo[:, 1] ** j * xo[:, 2] ** k).sum()
                assert abs(got - exact) < 1e-15
    assert H.quadrature(3, True, 4).weights.size == 11 and H.quadrature(3, True, 4).weights[0] < 0
    assert H.quadrature(3, True, 6).weights.size == 64            # beyond the tables: Duffy
    assert H.quadrature(2, True, 2).weights.size == 4             # triangles: always Duffy


def test_c_oracle_elasticity_equals_the_numpy_oracle():
    """the timed CPU baseline of config 4 (C port, vector-valued P2 tets, isotropic elasticity, Strang degree-4 rule)
    reproduces the numpy oracle bit for bit"""
    import c_oracle
    mesh, V, tab = problem((3, 2, 2), order=2, bc=[1], n_comp=3, simplexify=True, warp=0.1)
    cp, rv, nz = oracle_matrix(O.ELASTICITY, mesh, V, tab, alpha=1.0, lam=1.3, mu=0.7)
    c_oracle.set_vector_space(3, 1.3, 0.7)
    try:
        out = c_oracle.assemble(3, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tab_dict(tab), nthreads=3,
                                nnz_cap=mesh.n_cells * 900)
    finally:
        c_oracle.set_vector_space(1)
    assert tab.w.size == 11
    assert np.array_equal(out[0], cp) and np.array_equal(out[1], rv) and out[2].tobytes() == nz.tobytes()
