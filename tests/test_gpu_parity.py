"""GPU parity: libgtkasm (through the C ABI / ctypes) vs the oracle on identical inputs.
Pattern (colptr/rowval) bit-exact, values within 1e-12 norm-relative."""
import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, make_engine, oracle_matrix, oracle_vector, problem

E = gtk_b200.engine
pytestmark = pytest.mark.gpu

MAT_CASES = [
    # cells, bc, warp, form
    ((64, 64), "boundary", 0.0, "laplace"),          # BASELINE config 1
    ((64, 64), None, 0.0, "laplace"),
    ((7, 5), [1, 4], 0.2, "laplace"),
    ((9, 6), "boundary", 0.2, "mass"),
    ((8, 8, 8), "boundary", 0.0, "laplace"),
    ((16, 16, 16), "boundary", 0.0, "laplace"),
    ((12, 9, 7), "boundary", 0.2, "laplace"),        # ragged + non-affine cells
    ((6, 5, 4), None, 0.2, "laplace"),
    ((5, 6, 7), [1, 6], 0.15, "mass"),
    ((2, 2, 2), "boundary", 0.0, "laplace"),         # 1 free dof
    ((32, 32, 32), "boundary", 0.1, "laplace"),
]
FORMS = {"laplace": (O.LAPLACE, E.FORM_LAPLACE), "mass": (O.MASS, E.FORM_MASS), "elasticity": (O.ELASTICITY, E.FORM_ELASTICITY_ISO)}


@pytest.mark.parametrize("cells,bc,warp,form", MAT_CASES)
def test_matrix_parity(cells, bc, warp, form):
    mesh, V, tab = problem(cells, bc=bc, warp=warp)
    oform, gform = FORMS[form]
    colptr, rowval, nzval = oracle_matrix(oform, mesh, V, tab, alpha=1.0)
    eng = make_engine(mesh, V, tab)
    nnz = eng.matrix_symbolic()
    assert nnz == rowval.size
    cp, rv = eng.matrix_pattern()
    assert cp.dtype == np.int32 and rv.dtype == np.int32
    assert np.array_equal(cp, colptr)
    assert np.array_equal(rv, rowval)
    nz = eng.matrix_numeric(gform, alpha=1.0)
    assert_values_close(nz, nzval)
    # update_matrix!: re-assembly on the cached pattern is bit-identical
    nz2 = eng.matrix_numeric(gform, alpha=1.0)
    assert nz2.tobytes() == nz.tobytes()
    eng.close()


@pytest.mark.parametrize("cells,n_comp", [((6, 5), 2), ((5, 4, 3), 3)])
def test_elasticity_and_vector_mass(cells, n_comp):
    mesh, V, tab = problem(cells, bc=[1], n_comp=n_comp, warp=0.2)
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    cp, rv = eng.matrix_pattern()
    for form, kw in (("elasticity", dict(lam=1.3, mu=0.7)), ("mass", {}), ("laplace", {})):
        oform, gform = FORMS[form]
        okw = dict(kw)
        colptr, rowval, nzval = oracle_matrix(oform, mesh, V, tab, alpha=0.5, **okw)
        assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
        nz = eng.matrix_numeric(gform, alpha=0.5, **kw)
        assert_values_close(nz, nzval)
    eng.close()


HIGH_ORDER_CASES = [
    # cells, order, simplexify, n_comp, bc, warp, form, params
    ((7, 5), 2, False, 1, "boundary", 0.2, "laplace", {}),
    ((4, 3, 3), 2, False, 1, [1, 4], 0.15, "laplace", {}),
    ((3, 3, 2), 3, False, 1, "boundary", 0.1, "laplace", {}),          # config 3 element (Q3 hex, 64 dofs, 64 points)
    ((3, 2, 2), 3, False, 1, None, 0.0, "mass", {}),
    ((5, 4), 2, True, 1, "boundary", 0.2, "laplace", {}),              # P2 triangles, Duffy rule
    ((4, 3, 3), 1, True, 1, [2], 0.2, "laplace", {}),                  # P1 tets
    ((3, 2, 2), 2, True, 3, [1], 0.15, "elasticity", dict(lam=1.0, mu=1.0)),   # config 4 element (P2 x 3 on tets)
]


@pytest.mark.parametrize("cells,order,simplexify,n_comp,bc,warp,form,params", HIGH_ORDER_CASES)
def test_high_order_and_simplex_parity(cells, order, simplexify, n_comp, bc, warp, form, params):
    """The elements of BASELINE configs 3 and 4 (small meshes) with the reference's own face-complex dof numbering
    (refnumbering.py, checked against the oracle's literal restatement in tests/test_host_side.py)."""
    mesh, V, tab = problem(cells, order=order, bc=bc, n_comp=n_comp, simplexify=simplexify, warp=warp)
    oform, gform = FORMS[form]
    colptr, rowval, nzval = oracle_matrix(oform, mesh, V, tab, alpha=1.25, **params)
    eng = make_engine(mesh, V, tab)
    assert eng.matrix_symbolic() == rowval.size
    cp, rv = eng.matrix_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    nz = eng.matrix_numeric(gform, alpha=1.25, **params)
    assert_values_close(nz, nzval)
    assert eng.matrix_numeric(gform, alpha=1.25, **params).tobytes() == nz.tobytes()
    f = [1.0, -2.0, 0.5][:n_comp]
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=f), oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=f))
    eng.close()


def test_free_dirichlet_blocks():
    """Ad = free rows x Dirichlet columns (problems.jl:363-387) and the other selections."""
    mesh, V, tab = problem((6, 5, 4), bc=[1, 3, 6], warp=0.1)
    eng = make_engine(mesh, V, tab)
    for fd in [(E.FREE, E.DIRICHLET), (E.DIRICHLET, E.FREE), (E.DIRICHLET, E.DIRICHLET), (E.FREE, E.FREE)]:
        colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, fd=fd)
        nnz = eng.matrix_symbolic(*fd)
        cp, rv = eng.matrix_pattern()
        assert nnz == rowval.size and np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
        assert_values_close(eng.matrix_numeric(E.FORM_LAPLACE), nzval)
    eng.close()


VEC_CASES = [((64, 64), "boundary", 0.0), ((7, 5), [2], 0.2), ((16, 16, 16), "boundary", 0.0), ((6, 5, 4), None, 0.2)]


@pytest.mark.parametrize("cells,bc,warp", VEC_CASES)
def test_vector_parity(cells, bc, warp):
    mesh, V, tab = problem(cells, bc=bc, warp=warp)
    eng = make_engine(mesh, V, tab)
    eng.vector_symbolic(E.FREE)
    b_ref = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[1.0])
    b = eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=[1.0])
    assert_values_close(b, b_ref)
    rng = np.random.default_rng(1)
    fn = rng.standard_normal(mesh.n_nodes)
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_NODAL, f_nodal=fn, alpha=2.0),
                        oracle_vector(O.SOURCE_NODAL, mesh, V, tab, f_nodal=fn, alpha=2.0))
    fq = rng.standard_normal((mesh.n_cells, tab.w.size, 1))
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_QP, f_qp=fq), oracle_vector(O.SOURCE_QP, mesh, V, tab, f_qp=fq))
    b2 = eng.vector_assemble(E.FORM_SOURCE_QP, f_qp=fq)
    b3 = eng.vector_assemble(E.FORM_SOURCE_QP, f_qp=fq)
    assert b2.tobytes() == b3.tobytes()
    if V.n_dirichlet:
        eng.vector_symbolic(E.DIRICHLET)
        assert_values_close(eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=[1.0]),
                            oracle_vector(O.SOURCE_CONST, mesh, V, tab, fd=O.DIRICHLET, f_const=[1.0]))
    eng.close()


def test_matrix_and_vector_fused_matches_separate():
    mesh, V, tab = problem((16, 16, 16), bc="boundary", warp=0.1)
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    _, _, nz_ref = oracle_matrix(O.LAPLACE, mesh, V, tab)
    b_ref = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[1.0])
    assert_values_close(nz, nz_ref)
    assert_values_close(b, b_ref)
    runs = [eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0])) for _ in range(3)]
    for nz_i, b_i in runs:      # bit-reproducible
        assert nz_i.tobytes() == nz.tobytes() and b_i.tobytes() == b.tobytes()
    eng.close()


def test_unsupported_form_raises_and_state_errors():
    mesh, V, tab = problem((4, 4), bc="boundary")
    eng = make_engine(mesh, V, tab)
    with pytest.raises(E.GtkError):
        eng.matrix_numeric(E.FORM_LAPLACE)            # numeric before symbolic
    eng.matrix_symbolic()
    with pytest.raises(E.UnsupportedFormError):
        eng.matrix_numeric(77)
    with pytest.raises(E.UnsupportedFormError):
        eng.matrix_numeric(E.FORM_ELASTICITY_ISO, lam=1.0, mu=1.0)   # scalar space
    eng.vector_symbolic()
    with pytest.raises(E.UnsupportedFormError):
        eng.vector_assemble(55)
    eng.close()


def test_empty_and_all_dirichlet():
    """1x1 cell: every dof is on the boundary -> empty free system (ragged/empty edge case)."""
    mesh, V, tab = problem((1, 1), bc="boundary")
    assert V.n_free == 0
    eng = make_engine(mesh, V, tab)
    assert eng.matrix_symbolic() == 0
    cp, rv = eng.matrix_pattern()
    assert cp.tolist() == [1] and rv.size == 0
    assert eng.matrix_numeric(E.FORM_LAPLACE).size == 0
    # Dirichlet x Dirichlet block of the same space is the full 4x4 element matrix
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, fd=(O.DIRICHLET, O.DIRICHLET))
    assert eng.matrix_symbolic(E.DIRICHLET, E.DIRICHLET) == 16
    cp, rv = eng.matrix_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    assert_values_close(eng.matrix_numeric(E.FORM_LAPLACE), nzval)
    eng.close()


COEF_CASES = [((7, 6), 1, False), ((5, 4, 3), 1, False), ((3, 3, 2), 2, False), ((3, 2, 2), 3, False), ((4, 3, 3), 1, True), ((12, 10, 8), 1, False)]


@pytest.mark.parametrize("cells,order,simplexify", COEF_CASES)
def test_coefficient_fields(cells, order, simplexify):
    """∫ κ ∇u·∇v and ∫ κ u v with κ a nodal field (DiscreteField parameter of update_matrix!) or sampled
    at the quadrature points (AnalyticalField): generic kernel, element-GEMM path, and the sweep kernels stepping aside."""
    mesh, V, tab = problem(cells, order=order, bc=[1], simplexify=simplexify, warp=0.15)
    rng = np.random.default_rng(7)
    kn = 1.0 + rng.random(mesh.n_nodes)
    kq = 1.0 + rng.random((mesh.n_cells, tab.w.size))
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    for form, oform in ((E.FORM_LAPLACE, O.LAPLACE), (E.FORM_MASS, O.MASS)):
        for kw in (dict(coef_nodal=kn), dict(coef_qp=kq)):
            _, _, ref = oracle_matrix(oform, mesh, V, tab, alpha=0.5, **kw)
            nz = eng.matrix_numeric(form, alpha=0.5, **kw)
            assert eng.info(5) in (0, 3)                 # never the constant-coefficient sweep kernels
            assert_values_close(nz, ref)
            assert eng.matrix_numeric(form, alpha=0.5, **kw).tobytes() == nz.tobytes()
    # update_matrix! with a new coefficient on the cached pattern; κ ≡ 1 reproduces the plain form
    _, _, plain = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=0.5)
    assert_values_close(eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.5, coef_nodal=np.ones(mesh.n_nodes)), plain)
    with pytest.raises(E.GtkError):
        eng.matrix_numeric(E.FORM_LAPLACE, coef_nodal=kn, coef_qp=kq)
    eng.close()
