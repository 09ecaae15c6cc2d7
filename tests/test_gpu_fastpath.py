"""The structured Q1-hex sweep kernel (csrc/fastq1.cu) against the oracle AND against the generic
sort/segmented-reduce path of the same library, on meshes that stress its index logic."""
import os

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, make_engine, oracle_matrix, oracle_vector, problem

E = gtk_b200.engine
pytestmark = pytest.mark.gpu

CASES = [
    ((16, 8, 9), "boundary", 0.0),      # exactly one footprint
    ((17, 9, 8), "boundary", 0.2),      # one node past the footprint in x and y
    ((33, 20, 21), "boundary", 0.2),    # several footprints and z-segments, ragged
    ((5, 4, 3), None, 0.2),             # no BC: corner dofs numbered first (non-monotone columns)
    ((9, 7, 30), [1, 4, 5], 0.1),       # partial Dirichlet boundary, long z sweep
    ((2, 2, 2), "boundary", 0.0),
    ((1, 1, 1), None, 0.0),
    ((40, 3, 2), [3], 0.15),
    ((48, 48, 48), "boundary", 0.1),
]


@pytest.mark.parametrize("cells,bc,warp", CASES)
def test_fast_path_matches_oracle_and_generic(cells, bc, warp):
    mesh, V, tab = problem(cells, bc=bc, warp=warp)
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=1.5)
    b_ref = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[2.0], alpha=0.5)
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    cp, rv = eng.matrix_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.5), E.FORM_SOURCE_CONST, dict(f_const=[2.0], alpha=0.5))
    fast = 2 if warp == 0.0 else 1       # 2: exactly-affine kernel (Cartesian coordinates), 1: general sweep kernel
    assert eng.info(5) == fast, "the structured fast path was not taken"
    assert eng.info(0) <= 2, "the fast path is a single kernel launch (+ one classification kernel per coordinate upload)"
    assert_values_close(nz, nzval)
    assert_values_close(b, b_ref)
    # matrix-only and vector-only entry points use the same kernel
    assert_values_close(eng.matrix_numeric(E.FORM_LAPLACE, alpha=1.5), nzval)
    assert eng.info(5) == fast and eng.info(0) == 1
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=[2.0], alpha=0.5), b_ref)
    # bit-reproducible
    nz2, b2 = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.5), E.FORM_SOURCE_CONST, dict(f_const=[2.0], alpha=0.5))
    assert nz2.tobytes() == nz.tobytes() and b2.tobytes() == b.tobytes()
    eng.close()
    if fast == 2:   # the general sweep kernel on the same (affine) mesh
        os.environ["GTK_DISABLE_AFFINE"] = "1"
        try:
            eng = make_engine(mesh, V, tab)
            eng.matrix_symbolic()
            nz_s, b_s = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.5), E.FORM_SOURCE_CONST, dict(f_const=[2.0], alpha=0.5))
            assert eng.info(5) == 1
            eng.close()
        finally:
            del os.environ["GTK_DISABLE_AFFINE"]
        assert_values_close(nz, nz_s)
        assert_values_close(b, b_s)
    # generic path of the same library (sort-based symbolic phase, staged element matrices, segmented reduction)
    os.environ["GTK_DISABLE_FASTPATH"] = "1"
    os.environ["GTK_DISABLE_DMMA"] = "1"
    try:
        eng = make_engine(mesh, V, tab)
        eng.matrix_symbolic()
        cp_g, rv_g = eng.matrix_pattern()
        nz_g, b_g = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.5), E.FORM_SOURCE_CONST, dict(f_const=[2.0], alpha=0.5))
        assert eng.info(5) == 0
        eng.close()
    finally:
        del os.environ["GTK_DISABLE_FASTPATH"]
        del os.environ["GTK_DISABLE_DMMA"]
    assert np.array_equal(cp_g, cp) and np.array_equal(rv_g, rv)
    assert_values_close(nz, nz_g)
    assert_values_close(b, b_g)


def _graded(mesh, cells):
    """non-uniform tensor-product spacing: still exactly affine cells"""
    X = mesh.node_coordinates
    for d in range(3):
        X[:, d] = X[:, d] ** (1.0 + 0.35 * d) * (1.0 + d) - 0.3 * d
    return mesh


@pytest.mark.parametrize("cells,bc", [((16, 8, 8), "boundary"), ((8, 8, 4), [1, 6])])
def test_affine_kernel_on_sheared_meshes(cells, bc):
    """A global shear with dyadic factors keeps every cell an exact parallelepiped with NON-orthogonal edges: the affine kernel
    must take its full six-coefficient branch (the three mixed terms B01, B02, B12 are non-zero), not the orthogonal-cell
    shortcut a Cartesian mesh takes."""
    mesh, V, tab = problem(cells, bc=bc)
    X = mesh.node_coordinates
    X[:, 0] += 0.5 * X[:, 1] + 0.25 * X[:, 2]
    X[:, 1] += 0.5 * X[:, 2]
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=1.25)
    b_ref = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[2.0])
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.25), E.FORM_SOURCE_CONST, dict(f_const=[2.0]))
    assert eng.info(5) == 2, "a dyadic shear keeps the cells exactly affine"
    assert_values_close(nz, nzval)
    assert_values_close(b, b_ref)
    # half of the mesh sheared, the other half Cartesian: warps see orthogonal and non-orthogonal layers / patches side by side
    mesh2, V2, tab2 = problem(cells, bc=bc)
    Y = mesh2.node_coordinates
    upper = Y[:, 2] >= 0.5
    Y[upper, 0] += 0.5 * (Y[upper, 2] - 0.5)
    ref2 = oracle_matrix(O.LAPLACE, mesh2, V2, tab2, alpha=1.25)
    eng2 = make_engine(mesh2, V2, tab2)
    eng2.matrix_symbolic()
    nz2 = eng2.matrix_numeric(E.FORM_LAPLACE, alpha=1.25)
    assert eng2.info(5) == 2
    assert_values_close(nz2, ref2[2])
    eng.close(); eng2.close()


@pytest.mark.parametrize("cells,bc", [((18, 9, 11), "boundary"), ((7, 5, 6), None), ((33, 17, 20), [2, 3])])
def test_affine_kernel_on_graded_and_mixed_meshes(cells, bc):
    mesh, V, tab = problem(cells, bc=bc)
    _graded(mesh, cells)
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=0.75)
    b_ref = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[3.0])
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic()
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=0.75), E.FORM_SOURCE_CONST, dict(f_const=[3.0]))
    assert eng.info(5) == 2, "graded Cartesian meshes are exactly affine"
    assert_values_close(nz, nzval)
    assert_values_close(b, b_ref)
    import scipy.sparse as sp
    A = sp.csc_matrix((nz, rowval.astype(np.int64) - 1, colptr.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
    assert abs(A - A.T).max() == 0.0, "the affine kernel keeps A bitwise symmetric"
    # move ONE interior node: the classification must notice and the general kernel must take over
    X = mesh.node_coordinates.copy()
    inner = np.flatnonzero(~gtk_b200.hostprep.boundary_node_mask(mesh))
    k = inner[len(inner) // 2]
    X[k] += 1e-9
    mesh.node_coordinates[:] = X
    _, _, nzval2 = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=0.75)
    eng.update_coordinates(X)
    nz2 = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.75)
    assert eng.info(5) in (1, 6), "non-affine cells must be handled by the general kernel (1: everywhere, 6: on their tiles only)"
    assert_values_close(nz2, nzval2)
    assert np.abs(nz2 - nz).max() > 0
    eng.close()


@pytest.mark.parametrize("cells,n_moved", [((40, 30, 50), 1), ((40, 30, 50), 12), ((24, 40, 30), 300)])   # 300: most tiles -> general kernel
def test_mixed_mesh_affine_kernel_plus_general_kernel_on_marked_tiles(cells, n_moved):
    """A mostly affine mesh with a few displaced nodes: the affine kernel runs everywhere and the general kernel recomputes
    the columns of the tiles that hold a node of a non-affine cell (fast-path id 6) — one moved node must not cost 3x.
    Matrix and vector against the oracle; forcing the general kernel everywhere gives the same values to rounding."""
    mesh, V, tab = problem(cells, bc="boundary", domain=(0, 1, 0, 0.8, 0, 1.3))
    rng = np.random.default_rng(n_moved)
    inner = np.flatnonzero(~gtk_b200.hostprep.boundary_node_mask(mesh))
    moved = rng.choice(inner, size=n_moved, replace=False)
    h = 1.0 / max(cells)
    mesh.node_coordinates[moved] += 0.2 * h * rng.uniform(-1, 1, size=(n_moved, 3))
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=1.25)
    b_ref = oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[2.0])
    eng = make_engine(mesh, V, tab)
    eng.matrix_symbolic(); eng.vector_symbolic()
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.25), E.FORM_SOURCE_CONST, dict(f_const=[2.0]))
    assert eng.info(5) == (6 if n_moved <= 12 else 1), "few non-affine cells: mixed mode; many: the general kernel everywhere"
    assert_values_close(nz, nzval); assert_values_close(b, b_ref)
    nz2, b2 = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.25), E.FORM_SOURCE_CONST, dict(f_const=[2.0]))
    assert nz.tobytes() == nz2.tobytes() and b.tobytes() == b2.tobytes()
    os.environ["GTK_DISABLE_MIXED"] = "1"
    try:
        eng.update_coordinates(mesh.node_coordinates)      # forces a new classification
        nz_g, b_g = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.25), E.FORM_SOURCE_CONST, dict(f_const=[2.0]))
        assert eng.info(5) == 1
    finally:
        del os.environ["GTK_DISABLE_MIXED"]
    assert_values_close(nz, nz_g, tol=1e-13); assert_values_close(b, b_g, tol=1e-13)
    eng.close()


def test_fast_path_declines_what_it_does_not_cover():
    # shuffled cell order: not a structured topology any more -> generic path, same answer
    mesh, V, tab = problem((6, 5, 4), bc="boundary", warp=0.1)
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab)
    perm = np.random.default_rng(3).permutation(mesh.n_cells)
    eng = gtk_b200.engine.Engine(0)
    eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes[perm])
    eng.set_space(V.cell_dofs[perm], V.n_free, V.n_dirichlet)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    eng.matrix_symbolic()
    cp, rv = eng.matrix_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    nz = eng.matrix_numeric(E.FORM_LAPLACE)
    assert eng.info(5) == 5             # not the sweep kernels (1, 2): the fused Q1 cell kernel of unstructured meshes
    assert_values_close(nz, nzval)      # other summation order than the oracle's cell order: still within 1e-12
    # mass on the structured mesh: fast path declines (LAPLACE only)
    eng2 = make_engine(mesh, V, tab)
    eng2.matrix_symbolic()
    eng2.matrix_numeric(E.FORM_MASS)
    assert eng2.info(5) == 0
    eng.close(); eng2.close()


def test_full_size_config2_invariants():
    """BASELINE config 2 at full size (128^3): size-independent properties instead of an oracle run:
    nnz = (3m-2)^3, row sums of the Dirichlet-reduced Laplacian are >= 0 and vanish for rows not touching the
    boundary, A is bitwise symmetric, sum(b) = volume of the union of interior-node supports."""
    n = 128
    mesh, V, tab = problem((n, n, n), bc="boundary")
    eng = make_engine(mesh, V, tab)
    nnz = eng.matrix_symbolic()
    assert nnz == (3 * (n - 1) - 2) ** 3 == 54439939
    cp, rv = eng.matrix_pattern()
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    assert eng.info(5) == 2
    import scipy.sparse as sp
    A = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
    assert abs(A - A.T).max() == 0.0
    rs = np.asarray(A.sum(axis=1)).ravel()
    interior = np.diff(cp) == 27
    assert np.abs(rs[interior]).max() < 1e-12 * np.abs(nz).max()
    assert rs.min() > -1e-12 * np.abs(nz).max()
    # diagonal of the uniform-mesh Q1 Laplacian: 8 cells * h/3 each
    h = 1.0 / n
    assert np.allclose(A.diagonal(), 8 * h / 3, rtol=1e-12)
    # Σ_i b_i = ∫ Σ_i N_i = volume minus the part carried by boundary shape functions
    assert abs(b.sum() - (1 - h) ** 3) < 1e-12
    nz2, b2 = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    assert nz2.tobytes() == nz.tobytes() and b2.tobytes() == b.tobytes()
    eng.close()


STRUCT_CASES = [((12, 9, 7), "boundary"), ((5, 4, 3), None), ((9, 7, 11), [1, 4, 5]), ((1, 1, 1), None), ((2, 2, 2), "boundary"),
                ((3, 1, 1), [2]), ((24, 24, 24), "boundary")]


@pytest.mark.parametrize("cells,bc", STRUCT_CASES)
def test_structured_symbolic_equals_sort_based(cells, bc):
    """The lattice-derived pattern (no COO, no sort) is bit-identical to the sort-based one and to the oracle's; the
    sweep tables built from it give bit-identical values; N_coo agrees; forms outside the sweep kernels build the
    generic plan lazily on the same pattern."""
    mesh, V, tab = problem(cells, bc=bc, warp=0.15)
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab)
    eng = make_engine(mesh, V, tab)
    nnz = eng.matrix_symbolic()
    cp, rv = eng.matrix_pattern()
    assert nnz == rowval.size and np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    ncoo = eng.info(4)
    nz = eng.matrix_numeric(E.FORM_LAPLACE)
    fast = eng.info(5)
    assert fast in (1, 2)      # 2 when the mesh has no interior node to displace
    assert_values_close(nz, nzval)
    # lazily built generic plan: MASS is not a sweep-kernel form
    _, _, mass = oracle_matrix(O.MASS, mesh, V, tab)
    assert_values_close(eng.matrix_numeric(E.FORM_MASS), mass)
    assert eng.info(5) == 0
    cp2, rv2 = eng.matrix_pattern()
    assert np.array_equal(cp2, cp) and np.array_equal(rv2, rv)
    assert eng.matrix_numeric(E.FORM_LAPLACE).tobytes() == nz.tobytes() and eng.info(5) == fast
    fn = np.random.default_rng(3).standard_normal(mesh.n_nodes)
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_NODAL, f_nodal=fn), oracle_vector(O.SOURCE_NODAL, mesh, V, tab, f_nodal=fn))
    assert_values_close(eng.vector_assemble(E.FORM_SOURCE_CONST, f_const=[1.0]), oracle_vector(O.SOURCE_CONST, mesh, V, tab, f_const=[1.0]))
    eng.close()
    os.environ["GTK_DISABLE_STRUCT_SYMBOLIC"] = "1"
    try:
        eng = make_engine(mesh, V, tab)
        assert eng.matrix_symbolic() == nnz
        cp3, rv3 = eng.matrix_pattern()
        assert eng.info(4) == ncoo
        nz3 = eng.matrix_numeric(E.FORM_LAPLACE)
        assert eng.info(5) == fast
        eng.close()
    finally:
        del os.environ["GTK_DISABLE_STRUCT_SYMBOLIC"]
    assert np.array_equal(cp3, cp) and np.array_equal(rv3, rv)
    assert nz3.tobytes() == nz.tobytes()


def test_structured_symbolic_rejects_scrambled_topology():
    """A mesh whose cells are permuted is not a lattice in the reference's numbering: the structured phase must step
    aside and the sort-based phase must still give the oracle's pattern."""
    mesh, V, tab = problem((6, 5, 4), bc="boundary", warp=0.1)
    perm = np.random.default_rng(5).permutation(mesh.n_cells)
    mesh.cell_nodes[:] = mesh.cell_nodes[perm]
    V.cell_dofs[:] = V.cell_dofs[perm]
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab)
    eng = make_engine(mesh, V, tab)
    assert eng.matrix_symbolic() == rowval.size
    cp, rv = eng.matrix_pattern()
    assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
    assert_values_close(eng.matrix_numeric(E.FORM_LAPLACE), nzval)
    assert eng.info(5) != 1 and eng.info(5) != 2
    eng.close()
