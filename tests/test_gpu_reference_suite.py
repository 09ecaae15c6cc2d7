"""The reference's own tests for this path, transcribed to the host mirror (galerkintoolkit.jl_b200/gt.py) and run on the
GPU engine: test/problems_tests.jl (known answers sum(b) = sum(M) = 1, reuse + update_*!, the manufactured Poisson problem
solved through the linear-problem path) and the sizes pinned by test/assembly_tests.jl:58-73.
Julia's `≈` is rtol = sqrt(eps) = 1.5e-8; the gates here are 1e-10 (the order-3 tabulation, a monomial Vandermonde solve
as in the reference's `tabulator`, space.jl:960-970, is itself only good to ~1e-13)."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

import gtk_b200

GT = gtk_b200.gt
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("domain,cells,k", [((0, 1, 0, 1), (2, 2), 1), ((0, 1, 0, 1), (4, 3), 2), ((0, 1, 0, 1, 0, 1), (3, 2, 2), 1),
                                            ((0, 1, 0, 1, 0, 1), (2, 2, 2), 3)])
def test_problems_tests_jl(domain, cells, k):
    """test/problems_tests.jl:9-105 (cells = (2,2), k = 1 is the reference's own case)"""
    T = np.float64
    mesh = GT.cartesian_mesh(domain, cells)
    degree = 2 * k
    Om = GT.interior(mesh)
    Gd = GT.boundary(mesh)
    dOm = GT.measure(Om, degree)
    grad, dot = GT.grad, GT.dot

    V = GT.lagrange_space(Om, k)
    a = lambda u, v: GT.integrate(lambda x: dot(grad(u, x), grad(v, x)), dOm)
    A = GT.assemble_matrix(a, T, V, V)
    assert A.m == A.n == V.num_free_dofs()
    assert abs(A.sum()) < 1e-11 * np.abs(A.nzval).max()            # constants are in the kernel of the Laplacian (no BC)

    l = lambda v: GT.integrate(lambda x: v(x), dOm)
    a = lambda u, v: GT.integrate(lambda x: u(x) * v(x), dOm)

    b = GT.assemble_vector(l, T, V)
    assert np.isclose(b.sum(), 1.0, rtol=0, atol=1e-10)             # @test sum(b) ≈ 1       (:57)

    A = GT.assemble_matrix(a, T, V, V)
    assert np.isclose(A.sum(), 1.0, rtol=0, atol=1e-10)             # @test sum(A) ≈ 1       (:60)

    b, bcache = GT.assemble_vector(l, T, V, reuse=True)
    assert np.isclose(b.sum(), 1.0, rtol=0, atol=1e-10)
    b0 = b.copy()
    b[:] = 0.0                                                      # fill!(b, 0.0); update_vector!(b, bcache)   (:66-71)
    for _ in range(3):
        GT.update_vector(b, bcache)
    assert np.isclose(b.sum(), 1.0, rtol=0, atol=1e-10) and b.tobytes() == b0.tobytes()
    bcache.engine.close()

    A, Acache = GT.assemble_matrix(a, T, V, V, reuse=True)
    assert np.isclose(A.sum(), 1.0, rtol=0, atol=1e-10)
    nz0 = A.nzval.copy()
    A.nzval[:] = 0.0                                                # fill!(A, 0.0); update_matrix!(A, Acache)   (:76-81)
    for _ in range(2):
        GT.update_matrix(A, Acache)
    assert np.isclose(A.sum(), 1.0, rtol=0, atol=1e-10) and A.nzval.tobytes() == nz0.tobytes()
    Acache.engine.close()

    # manufactured solution g = sum(x): f = -Δg = 0, Dirichlet data interpolated, linear problem, LU, L2 error (:86-103)
    g = lambda x: sum(x[d] for d in range(len(cells)))
    V = GT.lagrange_space(Om, k, dirichlet_boundary=Gd)
    xd = GT.interpolate_dirichlet(g, V)
    a = lambda u, v: GT.integrate(lambda x: dot(grad(u, x), grad(v, x)), dOm)
    f = GT.analytical_field(lambda x: 0.0 * x[0])
    l = lambda v: GT.integrate(lambda x: v(x) * f(x), dOm)
    x0, A, rhs = GT.linear_problem(xd, a, l, V)
    assert x0.shape == (V.num_free_dofs(),)
    x = spla.spsolve(A.to_scipy().tocsc(), rhs) if V.num_free_dofs() > 1 else rhs / A.nzval[0]
    x = np.atleast_1d(x)
    eh = x - g(V.data.free_dof_nodes.T)                             # nodal error; u_h - g vanishes on Γd by interpolation
    m = lambda u, v: GT.integrate(lambda x: u(x) * v(x), dOm)
    M = GT.assemble_matrix(m, T, V, V).to_scipy()
    el2 = float(np.sqrt(max(eh @ (M @ eh), 0.0)))
    assert el2 + 1.0 == pytest.approx(1.0, abs=1e-10)               # @test el2 + 1.0 ≈ 1.0     (:103)
    uh = GT.solution_field(V, x, xd)
    assert uh.shape == V.face_dofs().shape


def test_assembly_tests_jl_sizes():
    """test/assembly_tests.jl:40-73: 2x2 mesh, full Dirichlet boundary: 1 free dof, 8 Dirichlet dofs; vector lengths and the
    shape of the free x Dirichlet block."""
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (2, 2))
    Om = GT.interior(mesh)
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=GT.boundary(mesh))
    dOm = GT.measure(Om, 2)
    assert V.num_free_dofs() == 1 and V.num_dirichlet_dofs() == 8                      # :72-73
    l = lambda v: GT.integrate(lambda x: v(x), dOm)
    b = GT.assemble_vector(l, np.float64, V)
    assert b.size == V.num_free_dofs()                                                   # :58
    bd = GT.assemble_vector(l, np.float64, V, free_or_dirichlet=GT.DIRICHLET)
    assert bd.size == V.num_dirichlet_dofs()                                             # :63
    a = lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dOm)
    Ad = GT.assemble_matrix(a, np.float64, V, V, free_or_dirichlet=(GT.FREE, GT.DIRICHLET))
    assert (Ad.m, Ad.n) == (V.num_free_dofs(), V.num_dirichlet_dofs())                    # :68
    assert np.isclose(b.sum() + bd.sum(), 1.0, rtol=0, atol=1e-14)


def test_assembly_options_index_type_int64():
    """test/assembly_tests.jl:120-135: `assembly_options = (; matrix = (; index_type = Int))` gives the same matrix with
    64-bit colptr / rowval; unsupported options raise instead of being ignored."""
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (5, 4))
    Ω = GT.interior(mesh)
    V = GT.lagrange_space(Ω, 2, dirichlet_boundary=GT.boundary(mesh))
    dΩ = GT.measure(Ω, 4)
    a = lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dΩ)
    A32 = GT.assemble_matrix(a, np.float64, V, V)
    A64 = GT.assemble_matrix(a, np.float64, V, V, assembly_options=dict(index_type=np.int64))
    assert A32.colptr.dtype == np.int32 and A64.colptr.dtype == np.int64 and A64.rowval.dtype == np.int64
    assert np.array_equal(A32.colptr, A64.colptr) and np.array_equal(A32.rowval, A64.rowval)
    assert A32.nzval.tobytes() == A64.nzval.tobytes()
    with pytest.raises(GT.UnsupportedFormError):
        GT.assemble_matrix(a, np.float64, V, V, assembly_options=dict(eltype=np.float32))
