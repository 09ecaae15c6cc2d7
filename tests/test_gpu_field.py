"""DiscreteField parameters on the device: forms that depend on ∇u_h, scalar integrals, nodal
interpolation — and the reference's ONLY numeric golden on this path, the p-Laplacian L2 norm of
test/problems_ext_tests.jl:148-172."""
import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from gtk_b200 import gt as GT
from util import assert_values_close, make_engine, problem, tab_dict

pytestmark = pytest.mark.gpu
E = gtk_b200.engine

REFERENCE_GOLDEN_UHL2 = 0.09133166701839236      # /root/reference/test/problems_ext_tests.jl:172


def test_plaplacian_reference_golden_l2_norm():
    """Transcription of test/problems_ext_tests.jl:148-172: n = 10 Q1 quads, zero Dirichlet data on the whole boundary,
    random free start, q = 3, res = ∫ ∇v⋅flux(∇u) − v, jac = ∫ ∇v⋅dflux(∇du,∇u), Newton, then
    `abs(uhl2 - 0.09133166701839236) < 1e-10` with uhl2 = sqrt(sum(∫ abs2(uh)))."""
    n = 10
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (n, n))
    Ω = GT.interior(mesh)
    Γ = GT.boundary(mesh)
    order = 1
    degree = 2 * order
    V = GT.lagrange_space(Ω, order, dirichlet_boundary=Γ)
    dΩ = GT.measure(Ω, degree)
    uh = GT.rand_field(np.float64, V, rng=np.random.default_rng(1234))
    q = 3
    flux, dflux = GT.plaplacian_flux(q), GT.plaplacian_dflux(q)
    res = lambda u: lambda v: GT.integrate(lambda x: GT.dot(GT.grad(v, x), GT.call(flux, GT.grad(u, x))) - v(x), dΩ)
    jac = lambda u: lambda du, v: GT.integrate(lambda x: GT.dot(GT.grad(v, x), GT.call(dflux, GT.grad(du, x), GT.grad(u, x))), dΩ)
    prob = GT.nonlinear_problem(uh, res, jac)
    launches0 = prob.residual_cache.engine.info(1)
    sol = GT.newton_solve(prob)
    assert prob.residual_cache.engine.info(1) > launches0          # every Newton step re-assembled on the device
    uh = GT.solution_field(uh, sol)
    prob.close()
    uhl2 = np.sqrt(GT.integrate(lambda x: GT.abs2(uh(x)), dΩ).sum())
    tol = 1.0e-10
    assert abs(uhl2 - REFERENCE_GOLDEN_UHL2) < tol, uhl2
    # a different random start converges to the same discrete solution
    uh2 = GT.rand_field(np.float64, V, rng=np.random.default_rng(99))
    prob2 = GT.nonlinear_problem(uh2, res, jac)
    uh2 = GT.solution_field(uh2, GT.newton_solve(prob2))
    prob2.close()
    assert abs(np.sqrt(GT.integrate(lambda x: GT.abs2(uh2(x)), dΩ).sum()) - REFERENCE_GOLDEN_UHL2) < tol


@pytest.mark.parametrize("cells,order,q,warp", [((6, 5), 1, 3, 0.2), ((4, 3), 2, 3, 0.15), ((3, 3, 2), 1, 3, 0.2),
                                                ((5, 4), 1, 4, 0.0), ((3, 2, 2), 2, 2.5, 0.1)])
def test_plaplace_residual_and_jacobian_match_oracle(cells, order, q, warp):
    mesh, V, tab = problem(cells, order=order, warp=warp)
    rng = np.random.default_rng(3)
    x, xd = rng.random(V.n_free), rng.random(V.n_dirichlet)
    X, CN, CD, td = mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, tab_dict(tab)
    eng = make_engine(mesh, V, tab)
    eng.field_set_values(x, xd)
    eng.matrix_symbolic()
    cp, rv = eng.matrix_pattern()
    nz = eng.matrix_numeric(E.FORM_PLAPLACE_JACOBIAN, exponent=q, alpha=1.0)
    cpo, rvo, nzo = O.assemble_matrix_plaplace_jacobian(X, CN, CD, V.n_free, td, x, xd, q)
    assert np.array_equal(cp, cpo) and np.array_equal(rv, rvo)
    assert_values_close(nz, nzo)
    b = eng.vector_assemble(E.FORM_PLAPLACE_RESIDUAL, exponent=q, f_const=[1.0])
    assert_values_close(b, O.assemble_vector_plaplace_residual(X, CN, CD, V.n_free, td, x, xd, q))
    fq = rng.random((mesh.n_cells, tab.w.size))
    b = eng.vector_assemble(E.FORM_PLAPLACE_RESIDUAL, exponent=q, f_qp=fq, alpha=-0.5)
    assert_values_close(b, O.assemble_vector_plaplace_residual(X, CN, CD, V.n_free, td, x, xd, q, alpha=-0.5, f_qp=fq))
    # update_*! with new parameters: values change, repeated calls are byte-identical
    x2 = rng.random(V.n_free)
    eng.field_set_values(x2)
    nz2 = eng.matrix_numeric(E.FORM_PLAPLACE_JACOBIAN, exponent=q)
    assert_values_close(nz2, O.assemble_matrix_plaplace_jacobian(X, CN, CD, V.n_free, td, x2, xd, q)[2])
    assert nz2.tobytes() == eng.matrix_numeric(E.FORM_PLAPLACE_JACOBIAN, exponent=q).tobytes()
    fv, dv = eng.field_get_values()
    assert np.array_equal(fv, x2) and np.array_equal(dv, xd)
    eng.field_axpy_free(0.5, x)
    assert np.array_equal(eng.field_get_values()[0], x2 + 0.5 * x)
    eng.close()


@pytest.mark.parametrize("cells,order,warp", [((7, 5), 1, 0.2), ((3, 4), 3, 0.1), ((3, 2, 4), 2, 0.2)])
def test_scalar_integrals_match_oracle(cells, order, warp):
    mesh, V, tab = problem(cells, order=order, warp=warp, bc=[1])
    rng = np.random.default_rng(5)
    x, xd = rng.random(V.n_free), rng.random(V.n_dirichlet)
    X, CN, CD, td = mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, tab_dict(tab)
    D, nq = mesh.D, tab.w.size
    eng = make_engine(mesh, V, tab)
    eng.field_set_values(x, xd)
    vol = eng.scalar_assemble(E.SCALAR_VOLUME)
    assert abs(vol - O.assemble_scalar_field(O.SCALAR_VOLUME, X, CN, CD, td, x, xd)) <= 1e-13 and abs(vol - 1.0) < 1e-12
    g = rng.random((mesh.n_cells, nq)); dg = rng.random((mesh.n_cells, nq, D))
    for kind, okind, gq in ((E.SCALAR_L2SQ, O.SCALAR_L2SQ, None), (E.SCALAR_L2SQ, O.SCALAR_L2SQ, g),
                            (E.SCALAR_H1SQ, O.SCALAR_H1SQ, None), (E.SCALAR_H1SQ, O.SCALAR_H1SQ, dg)):
        got = eng.scalar_assemble(kind, **({} if gq is None else dict(f_qp=gq)))
        ref = O.assemble_scalar_field(okind, X, CN, CD, td, x, xd, g_qp=gq)
        assert abs(got - ref) <= 1e-12 * abs(ref), (kind, got, ref)
        assert got == eng.scalar_assemble(kind, **({} if gq is None else dict(f_qp=gq)))       # bit-reproducible
    eng.close()


def test_field_forms_reject_what_they_do_not_cover():
    mesh, V, tab = problem((3, 3), order=1, n_comp=2)
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    with pytest.raises(E.UnsupportedFormError):
        eng.matrix_numeric(E.FORM_PLAPLACE_JACOBIAN, exponent=3)
    with pytest.raises(E.UnsupportedFormError):
        eng.vector_assemble(E.FORM_PLAPLACE_RESIDUAL, exponent=3)
    with pytest.raises(E.UnsupportedFormError):
        eng.scalar_assemble(999)
    eng.close()
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (3, 3))
    V = GT.lagrange_space(GT.interior(mesh), 1, dirichlet_boundary=GT.boundary(mesh))
    dΩ = GT.measure(GT.interior(mesh), 2)
    uh = GT.zero_field(np.float64, V)
    opaque = lambda g: g            # an opaque closure handed to GT.call cannot be recognised: explicit error
    with pytest.raises(GT.UnsupportedFormError):
        GT.assemble_vector(lambda v: GT.integrate(lambda x: GT.dot(GT.grad(v, x), GT.call(opaque, GT.grad(uh, x))) - v(x), dΩ),
                           np.float64, V, parameters=(uh,))
    with pytest.raises(GT.UnsupportedFormError):
        GT.integrate(lambda x: uh(x) * uh(x) * uh(x), dΩ).sum()


def test_dirichlet_interpolation_and_solution_field_on_device():
    """interpolate_dirichlet / interpolate_free / solution_field (space.jl:1876-1897, 2000-2060; problems.jl:501-526):
    dof-node coordinates come from the device kernel (last cell wins, sequential tabulator sum) and equal the oracle's
    literal loop bit for bit."""
    mesh = GT.cartesian_mesh((0, 1, 0, 2, 0, 1), (3, 4, 2))
    V = GT.lagrange_space(GT.interior(mesh), 2, dirichlet_boundary=GT.boundary(mesh, ["2-face-1", "2-face-4"]))
    g = lambda x: 1.0 + x[0] - 2.0 * x[1] * x[2]
    xf, xdc = GT.dof_coordinates(V)
    Mn = GT._reference_node_tabulation(V)
    xf_o, xd_o = O.space_dof_coordinates(mesh.node_coordinates, mesh.cell_nodes, V.data.cell_dofs, V.data.n_free,
                                         V.data.n_dirichlet, Mn, 1)
    assert np.array_equal(xf, xf_o) and np.array_equal(xdc, xd_o)
    assert np.allclose(xf, V.data.free_dof_nodes, atol=1e-14) and np.allclose(xdc, V.data.dirichlet_dof_nodes, atol=1e-14)
    xd = GT.interpolate_dirichlet(g, V)
    assert np.array_equal(xd, g(xdc.T))
    uh = GT.interpolate(g, V)
    assert np.array_equal(uh.free_values, g(xf.T)) and np.array_equal(uh.dirichlet_values, xd)
    vals = GT.solution_field(V, uh.free_values, xd)
    lat = np.array(O._lattice(3, 2), dtype=np.float64) / 2
    X = mesh.node_coordinates[mesh.cell_nodes.astype(np.int64) - 1]
    xl = X[:, None, 0, :] + lat[None] * (X[:, -1, :] - X[:, 0, :])[:, None, :]
    assert np.allclose(vals, g(np.moveaxis(xl, -1, 0)))
    # g is a quadratic in the Q2 space: the interpolant reproduces it, so both error norms vanish (manufactured-solution
    # check of test/problems_ext_tests.jl:139-146 with the integrals on the device)
    dΩ = GT.measure(GT.interior(mesh), 4)
    u = GT.analytical_field(g, GT.interior(mesh), gradient=lambda x: np.stack([np.ones_like(x[0]), -2.0 * x[2], -2.0 * x[1]]))
    el2 = np.sqrt(GT.integrate(lambda x: GT.abs2(u(x) - uh(x)), dΩ).sum())
    eh1 = np.sqrt(GT.integrate(lambda x: GT.dot(GT.grad(u, x) - GT.grad(uh, x), GT.grad(u, x) - GT.grad(uh, x)), dΩ).sum())
    assert el2 < 1e-10 and eh1 < 1e-10
    # vector-valued space: every component dof of a node gets the node's coordinate, values by component
    Vv = GT.lagrange_space(GT.interior(mesh), 1, dirichlet_boundary=GT.boundary(mesh), tensor_size=(3,))
    gv = lambda x: np.stack([x[0], 2 * x[1], -x[2]])
    xdv = GT.interpolate_dirichlet(gv, Vv)
    comp = GT._dof_component(Vv, False)
    ref = gv(Vv.data.dirichlet_dof_nodes.T)
    assert np.allclose(xdv, ref[comp, np.arange(xdv.size)])
    uhd = GT.zero_field(np.float64, Vv)
    GT.interpolate_dirichlet(gv, uhd)
    assert np.array_equal(uhd.dirichlet_values, xdv) and not uhd.free_values.any()
