"""Boundary integrals (Neumann / Robin terms, the north-star's "source/Neumann terms"): the faces of a boundary domain
go to the engine as a mesh of (D-1)-cells embedded in D dimensions (gtk_set_manifold_dim), dV = sqrt(det(JᵀJ)) w with
the D x (D-1) Jacobian; sums of integrals continue one COO vector (accumulate).  Against the oracle on identical inputs."""
import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, tab_dict

E = gtk_b200.engine
H = gtk_b200.hostprep
GT = gtk_b200.gt
pytestmark = pytest.mark.gpu


def _bend(mesh):
    """smooth global map: boundary faces become curved / non axis-aligned (non-affine faces)"""
    X = mesh.node_coordinates
    Y = X.copy()
    Y[:, 0] += 0.08 * np.sin(2.0 * X[:, 1]) + (0.05 * X[:, -1] ** 2 if mesh.D == 3 else 0.0)
    Y[:, 1] += 0.06 * X[:, 0] * X[:, 1]
    mesh.node_coordinates[:] = Y


def _face_engine(mesh, V, fp):
    eng = E.Engine(0)
    eng.set_mesh(mesh.node_coordinates, fp.face_nodes)
    eng.set_manifold_dim(mesh.D - 1)
    eng.set_space(fp.face_dofs, V.n_free, V.n_dirichlet, V.n_comp)
    eng.set_tabulation(fp.tab.w, fp.tab.N, fp.tab.dN, fp.tab.M, fp.tab.dM)
    return eng


CASES = [
    # cells, order, dirichlet sides, neumann sides, n_comp   (negative order = simplexified mesh)
    ((6, 5), -1, [3], [4, 2], 1),            # P1 triangles
    ((4, 3), -2, None, None, 1),             # P2 triangles
    ((4, 3, 3), -1, [1], [2, 6], 1),         # P1 tets
    ((3, 2, 2), -2, [5], [2, 4], 3),         # config 4 element: P2 x 3 on tets, traction on two sides
    ((7, 5), 1, [3], [4, 2], 1),
    ((4, 3), 2, None, None, 1),
    ((5, 4, 3), 1, [1], [2, 6], 1),
    ((3, 3, 2), 2, [5], None, 1),
    ((3, 2, 2), 3, None, [4], 1),
    ((4, 3, 3), 1, [1], [2, 3], 3),
]


@pytest.mark.parametrize("cells,order,diri,neu,n_comp", CASES)
def test_neumann_vector_parity(cells, order, diri, neu, n_comp):
    D = len(cells)
    simplexify, order = order < 0, abs(order)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells, simplexify=simplexify)
    _bend(mesh)
    V = H.lagrange_space(mesh, order, diri, n_comp)
    fp = H.face_problem(V, neu, 2 * order)
    tabd = tab_dict(fp.tab)
    args = (mesh.node_coordinates, fp.face_nodes, fp.face_dofs, V.n_free, V.n_dirichlet, tabd)
    eng = _face_engine(mesh, V, fp)
    eng.vector_symbolic(E.FREE)
    g = [1.5, -0.5, 2.0][:n_comp]
    b = eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=g, alpha=0.75)
    assert_values_close(b, O.assemble_vector(O.SOURCE_CONST, *args, n_comp=n_comp, f_const=g, alpha=0.75))
    rng = np.random.default_rng(2)
    gq = rng.standard_normal((fp.face_nodes.shape[0], fp.tab.w.size, n_comp))
    bq = eng.vector_assemble(E.FORM_SOURCE_QP, f_qp=gq)
    assert_values_close(bq, O.assemble_vector(O.SOURCE_QP, *args, n_comp=n_comp, f_qp=gq))
    assert eng.vector_assemble(E.FORM_SOURCE_QP, f_qp=gq).tobytes() == bq.tobytes()
    gn = rng.standard_normal((mesh.n_nodes, n_comp))
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_NODAL, f_nodal=gn), O.assemble_vector(O.SOURCE_NODAL, *args, n_comp=n_comp, f_nodal=gn))
    # continue a vector that already holds the volume integral: one COO vector in the reference
    tabv = H.measure_tabulation(V, 2 * order)
    b_vol = O.assemble_vector(O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet,
                              tab_dict(tabv), n_comp=n_comp, f_const=[1.0] * n_comp)
    eng.set_vector(b_vol)
    b_sum = eng.vector_assemble(E.FORM_SOURCE_QP, f_qp=gq, accumulate=True)
    assert_values_close(b_sum, O.assemble_vector(O.SOURCE_QP, *args, n_comp=n_comp, f_qp=gq, b0=b_vol))
    # total of a constant flux over Γ = |Γ| * g for spaces without Dirichlet dofs (partition of unity on the faces)
    if diri is None and n_comp == 1:
        area = O.assemble_vector(O.SOURCE_CONST, *args, f_const=[1.0]).sum()
        assert abs(eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=[1.0]).sum() - area) <= 1e-12 * area
    # Dirichlet rows of the same boundary integral
    if V.n_dirichlet:
        eng.vector_symbolic(E.DIRICHLET)
        assert_values_close(eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=g),
                            O.assemble_vector(O.SOURCE_CONST, *args, n_comp=n_comp, f_const=g, free_or_dirichlet=O.DIRICHLET))
    eng.close()


@pytest.mark.parametrize("cells,order", [((6, 5), 1), ((4, 3, 3), 1), ((3, 2, 2), 2)])
def test_robin_boundary_mass_matrix(cells, order):
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells)
    _bend(mesh)
    V = H.lagrange_space(mesh, order, [1])
    fp = H.face_problem(V, [2, 2 * D], 2 * order)
    colptr, rowval, nzval = O.assemble_matrix(O.MASS, mesh.node_coordinates, fp.face_nodes, fp.face_dofs, V.n_free, V.n_dirichlet,
                                              tab_dict(fp.tab), alpha=2.5)
    eng = _face_engine(mesh, V, fp)
    assert eng.matrix_symbolic() == rowval.size
    cp, rv = eng.matrix_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    assert_values_close(eng.matrix_numeric(E.FORM_MASS, alpha=2.5), nzval)
    with pytest.raises(E.UnsupportedFormError):
        eng.matrix_numeric(E.FORM_LAPLACE)          # physical gradients on a face: explicit error, no fallback
    eng.close()


def test_gt_mirror_volume_plus_neumann():
    """l(v) = ∫_Ω f v dΩ + ∫_Γ g v dΓ through the GT-style host API, analytical f and g."""
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (5, 4, 3))
    Om = GT.interior(mesh)
    Gd = GT.boundary(mesh, ["2-face-1"])
    Gn = GT.boundary(mesh, ["2-face-2", "2-face-6"])
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=Gd)
    dOm, dGn = GT.measure(Om, 2), GT.measure(Gn, 2)
    f = GT.AnalyticalField(lambda x: x[0] + 2.0 * x[1] * x[2])
    g = GT.AnalyticalField(lambda x: 1.0 + x[0] * x[1])
    l = lambda v: GT.integrate(lambda x: f(x) * v(x), dOm) + GT.integrate(lambda x: g(x) * v(x), dGn)
    b = GT.assemble_vector(l, np.float64, V)
    Vd, m = V.data, mesh
    tabv = H.measure_tabulation(Vd, 2)
    xq = np.einsum("qn,cnd->cqd", tabv.M, m.node_coordinates[m.cell_nodes.astype(np.int64) - 1])
    b_ref = O.assemble_vector(O.SOURCE_QP, m.node_coordinates, m.cell_nodes, Vd.cell_dofs, Vd.n_free, Vd.n_dirichlet, tab_dict(tabv),
                              f_qp=(xq[..., 0] + 2.0 * xq[..., 1] * xq[..., 2])[..., None])
    fp = H.face_problem(Vd, [2, 6], 2)
    xf = np.einsum("qn,cnd->cqd", fp.tab.M, m.node_coordinates[fp.face_nodes.astype(np.int64) - 1])
    b_ref = O.assemble_vector(O.SOURCE_QP, m.node_coordinates, fp.face_nodes, fp.face_dofs, Vd.n_free, Vd.n_dirichlet, tab_dict(fp.tab),
                              f_qp=(1.0 + xf[..., 0] * xf[..., 1])[..., None], b0=b_ref)
    assert_values_close(b, b_ref)
