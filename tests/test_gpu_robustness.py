"""State handling of the C ABI: stale inputs, stale plans and out-of-range exchange plans are errors, never silent
out-of-bounds work on the device."""
import numpy as np
import pytest

import gtk_b200
from gtk_b200 import gt as GT
from util import make_engine, problem

pytestmark = pytest.mark.gpu
E = gtk_b200.engine


def test_fused_call_requires_one_measure():
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (3, 3, 3))
    Ω, Γ = GT.interior(mesh), GT.boundary(mesh, ["2-face-2"])
    V = GT.lagrange_space(Ω, 1, dirichlet_boundary=GT.boundary(mesh, ["2-face-1"]))
    dΩ, dΓ = GT.measure(Ω, 2), GT.measure(Γ, 2)
    a = lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dΩ)
    l = lambda v: GT.integrate(lambda x: GT.analytical_field(lambda x: x[0])(x) * v(x), dΓ)
    with pytest.raises(GT.UnsupportedFormError):
        GT.assemble_matrix_and_vector(a, l, np.float64, V, V)
    with pytest.raises(GT.UnsupportedFormError):
        GT.assemble_matrix_and_vector_with_free_and_dirichlet_columns(a, l, np.float64, V, V)


def test_new_mesh_invalidates_space_and_tabulation():
    mesh, V, tab = problem((3, 3, 3))
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    eng.matrix_numeric(E.FORM_LAPLACE)
    big, Vb, _ = problem((5, 5, 5))
    eng.set_mesh(big.node_coordinates, big.cell_nodes)
    with pytest.raises(E.GtkError):           # the old (smaller) cell_dofs must not be read with the new cell count
        eng.matrix_symbolic()
    eng.set_space(Vb.cell_dofs, Vb.n_free, Vb.n_dirichlet)
    eng.matrix_symbolic()
    with pytest.raises(E.GtkError):           # tabulation belongs to the previous setup
        eng.matrix_numeric(E.FORM_MASS)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    eng.matrix_numeric(E.FORM_MASS)
    eng.close()


def test_exchange_plan_is_range_checked_and_dies_with_its_pattern():
    mesh, V, tab = problem((3, 3, 3))
    eng = make_engine(mesh, V, tab)
    with pytest.raises(E.GtkError):
        eng.comm_set_exchange(1, [0], [], [], [])            # before the symbolic phase
    nnz = eng.matrix_symbolic()
    with pytest.raises(E.GtkError):
        eng.comm_set_exchange(1, [nnz], [], [], [])
    with pytest.raises(E.GtkError):
        eng.comm_set_exchange(1, [], [V.n_free], [], [])
    with pytest.raises(E.GtkError):
        eng.comm_set_exchange(1, [], [], [-1], [])
    eng.comm_set_exchange(1, [0, 1], [0], [2], [1])
    assert eng.comm_ghost_info(0) == 3 and eng.comm_ghost_info(1) == 2
    eng.matrix_symbolic()                                    # new pattern: the plan of the old one is gone
    assert eng.comm_ghost_info(0) == 0 and eng.comm_ghost_info(1) == 0
    eng.close()


def test_widening_the_active_cells_reclassifies_affine_layers():
    """the affine kernel may only be used if EVERY active cell layer is affine"""
    mesh, V, tab = problem((6, 6, 6))
    X = mesh.node_coordinates.copy()
    top = X[:, 2] > 0.7
    inner = ~gtk_b200.hostprep.boundary_node_mask(mesh)
    X[top & inner, 0] += 0.03 * np.sin(7 * X[top & inner, 1])            # non-affine cells in the upper layers only
    eng = make_engine(mesh, V, tab)
    eng.update_coordinates(X)
    eng.matrix_symbolic()
    eng.set_active_cells(0, 36 * 2)
    eng.assemble_matrix_and_vector(E.FORM_LAPLACE, {}, E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    assert eng.info(5) == 2                                              # affine kernel on the affine layers
    eng.set_active_cells(0, 36 * 6)
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, {}, E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    assert eng.info(5) in (1, 6)                                         # the general kernel now (everywhere, or on the marked tiles)
    import gt_oracle as O
    from util import tab_dict
    ref = O.assemble_matrix(O.LAPLACE, X, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tab_dict(tab))
    assert np.abs(nz - ref[2]).max() <= 1e-12 * np.abs(ref[2]).max()
    eng.close()


def test_block_and_sum_entry_points_reject_inconsistent_state():
    """gtk_set_parts / gtk_matrix_numeric_blocks / gtk_matrix_sum_*: wrong sizes, missing tables and mixed-up contexts are errors"""
    import importlib
    MF = importlib.import_module("galerkintoolkit_jl_b200.multifield")
    H = gtk_b200.hostprep
    mesh = H.cartesian_mesh((0, 1, 0, 1), (3, 3))
    V = H.lagrange_space(mesh, 1, [1])
    bp = MF.skeleton_problem([V], 2)
    eng = E.Engine(0)
    eng.set_mesh(bp.node_coordinates, bp.face_nodes)
    eng.set_manifold_dim(1)
    eng.set_space(bp.super_dofs, bp.n_free, bp.n_dirichlet, 1)
    with pytest.raises(E.GtkError):                       # the parts must add up to the super element of gtk_set_space
        eng.set_parts(bp.w, bp.M, bp.dM, bp.parts[:1], bp.n_sides, bp.face_var)
    bad = bp.face_var.copy(); bad[0, 0] = 99
    with pytest.raises(E.GtkError):                       # variant index outside the tables
        eng.set_parts(bp.w, bp.M, bp.dM, bp.parts, bp.n_sides, bad)
    with pytest.raises(E.GtkError):                       # a failed gtk_set_parts leaves no parts behind
        eng.matrix_symbolic(); eng.matrix_numeric_blocks([(0, 0, E.BLOCK_MASS, 1.0)])
    eng.set_parts(bp.w, bp.M, bp.dM, bp.parts, bp.n_sides, bp.face_var)
    eng.matrix_symbolic()
    with pytest.raises(E.GtkError):                       # part index out of range
        eng.matrix_numeric_blocks([(0, 7, E.BLOCK_MASS, 1.0)])
    with pytest.raises(E.UnsupportedFormError):           # gradients on faces without the cells around
        eng.matrix_numeric_blocks([(0, 0, E.BLOCK_LAPLACE, 1.0)])
    with pytest.raises(E.GtkError):                       # the plain numeric entry points refuse a context that holds parts
        eng.matrix_numeric(E.FORM_MASS)
    nz = eng.matrix_numeric_blocks([(0, 0, E.BLOCK_MASS, 1.0)])
    assert np.isfinite(nz).all()
    # sums: sources of different sizes, the destination among the sources, numeric before symbolic
    mesh2, V2, tab2 = problem((3, 3, 3))
    other = make_engine(mesh2, V2, tab2)
    other.matrix_symbolic(); other.matrix_numeric(E.FORM_MASS)
    total = E.Engine(0)
    with pytest.raises(E.GtkError):
        total.matrix_sum_symbolic([eng, other])
    with pytest.raises(E.GtkError):
        eng.matrix_sum_symbolic([eng])
    with pytest.raises(E.GtkError):
        total.matrix_sum_numeric([eng])
    total.matrix_sum_symbolic([eng, eng])                 # the same integral twice: same pattern, doubled values
    cp, rv = total.matrix_pattern()
    ecp, erv = eng.matrix_pattern()
    assert np.array_equal(cp, ecp) and np.array_equal(rv, erv)
    assert np.array_equal(total.matrix_sum_numeric([eng, eng]), 2.0 * nz)
    for e in (eng, other, total):
        e.close()
