"""multi-field spaces (CartesianProductSpace block offsets) and skeleton integrals.

CPU: the oracle's literal restatement of the generated multi-field / skeleton loops (compiler.jl:1826-2000, assembly.jl:
321-416) against known answers — the reference's own (`sum(b)+1 ≈ 1` for ∫_Λ jump(v), test/assembly_tests.jl:420-427) and
structural ones (blocks of V × V equal the single-field matrix bitwise, the Stokes coupling blocks are minus each other's
transpose) — and the package's vectorised input preparation (multifield.py) against the oracle's loops.
GPU: libgtkasm's block kernels against the oracle on identical inputs, through the C ABI and through the GT mirror."""
import importlib

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close

E = gtk_b200.engine
H = gtk_b200.hostprep
GT = gtk_b200.gt
MF = importlib.import_module("galerkintoolkit_jl_b200.multifield")


# ---- integrands written the way the reference's tests write them ---------------------------------------------------
def jump_u(p, f=0): return p.u(f, 2) - p.u(f, 1)          # jump(u,p) = u[2](p) - u[1](p)   (test/assembly_tests.jl:332)
def jump_v(p, f=0): return p.v(f, 2) - p.v(f, 1)


INTEGRANDS = {
    "jumpjump": lambda p: jump_u(p) * jump_v(p),                                   # test/assembly_tests.jl:397-401
    "sum_mass": lambda p: (p.v(0) + p.v(1)) * (p.u(0) + p.u(1)),                   # test/assembly_tests.jl:409-412
    "jump12": lambda p: jump_u(p, 0) * jump_v(p, 1),                               # test/assembly_tests.jl:413-415
    # docs/src/src_jl/example_stokes.jl: ∇(v,x)⋅∇(u,x) - div(v,x)*p(x) + q(x)*div(u,x)
    "stokes": lambda p: O.frobenius(p.grad_v(0), p.grad_u(0)) - p.div_v(0) * p.u(1) + p.v(1) * p.div_u(0),
}


def _oracle_fields(bp, spaces, with_gradients):
    """the oracle's field descriptors from a BlockProblem (one per field; tables of side 0 serve both sides)"""
    out = []
    for f, s in enumerate(spaces):
        part = bp.parts[bp.part_index(f, 0)]
        out.append(dict(cell_dofs=s.cell_dofs, n_free=s.n_free, n_dirichlet=s.n_dirichlet, n_comp=s.n_comp,
                        N=part["N"], dN=part["dN"] if with_gradients else None))
    return out


def _oracle_matrix(bp, spaces, mesh, name, alpha=1.0, fd=(O.FREE, O.FREE)):
    nf = bp.face_nodes.shape[0]
    if bp.n_sides == 2:
        sides = [[(int(bp.side_cells[i, a]), int(bp.face_var[i, a])) for a in range(2)] for i in range(nf)]
        geo = None
    else:
        sides = [[(i + 1, 0)] for i in range(nf)]
        geo = (mesh.cell_nodes, bp.dM)
    return O.assemble_matrix_multifield(mesh.D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                        _oracle_fields(bp, spaces, geo is not None), INTEGRANDS[name], alpha=alpha,
                                        free_or_dirichlet=fd, cell_geometry=geo)


def _warp(mesh, amount=0.15, seed=3):
    rng = np.random.default_rng(seed)
    inner = ~H.boundary_node_mask(mesh)
    h = 1.0 / np.array(mesh.cells_per_dir)
    mesh.node_coordinates[inner] += amount * h * rng.uniform(-1, 1, size=(int(inner.sum()), mesh.D))


def _stokes_spaces(cells, warp=True):
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells)
    if warp:
        _warp(mesh)
    V = H.lagrange_space(mesh, 2, "boundary", D)
    Q = H.lagrange_space(mesh, 1, None, 1)
    return mesh, [V, Q]


# ---- CPU: host preparation against the oracle's loops ---------------------------------------------------------------
@pytest.mark.parametrize("cells", [(3, 2), (4, 4), (2, 3, 2)])
def test_skeleton_inputs_equal_the_literal_restatement(cells):
    D = len(cells)
    dom = tuple([0, 1] * D)
    mesh = H.cartesian_mesh(dom, cells)
    V = H.lagrange_space(mesh, 2, [1])
    bp = MF.skeleton_problem([V], 4)
    coords, cn = O.cartesian_chain(dom, cells)
    sf = O.skeleton_faces(cn, coords.shape[0], D)
    # number of interior faces of a Cartesian mesh
    assert len(sf) == sum(int(np.prod([c - (1 if k == d else 0) for k, c in enumerate(cells)])) for d in range(D))
    assert len(sf) == bp.face_nodes.shape[0]
    q = H.quadrature(D - 1, False, 4)
    tabs = O.reference_map_tables(D, q.coordinates)
    for i, f in enumerate(sf):
        assert list(bp.face_nodes[i]) == f["nodes"]
        assert [s[0] for s in f["sides"]] == list(bp.side_cells[i])
        for a, (cell, lface, perm) in enumerate(f["sides"]):
            No, _ = O.tabulate(D, 2, "Q", np.array(tabs[lface - 1][perm - 1]))
            assert np.abs(No - bp.parts[a]["N"][bp.face_var[i, a]]).max() < 1e-13
            # the mapped points are the same physical points on both sides: face geometry == cell geometry there
            Mf, _ = H.tabulate(D - 1, 1, "Q", q.coordinates)
            xf = Mf @ mesh.node_coordinates[bp.face_nodes[i] - 1]
            Mc, _ = H.tabulate(D, 1, "Q", np.array(tabs[lface - 1][perm - 1]))
            xc = Mc @ mesh.node_coordinates[mesh.cell_nodes[cell - 1] - 1]
            assert np.abs(xf - xc).max() < 1e-13
    # super dof table: [side 0 dofs, side 1 dofs]
    nld = V.cell_dofs.shape[1]
    assert np.array_equal(bp.super_dofs[:, :nld], V.cell_dofs[bp.side_cells[:, 0] - 1])
    assert np.array_equal(bp.super_dofs[:, nld:], V.cell_dofs[bp.side_cells[:, 1] - 1])


def test_block_offsets_of_a_product_space():
    mesh, (V, Q) = _stokes_spaces((3, 2), warp=False)
    dofs, nfree, ndiri, fo, do = MF.offset_dofs([V, Q])
    assert nfree == V.n_free + Q.n_free and ndiri == V.n_dirichlet + Q.n_dirichlet
    assert list(fo) == O.monolithic_offsets([V.n_free, Q.n_free]) and list(do) == O.monolithic_offsets([V.n_dirichlet, Q.n_dirichlet])
    assert np.array_equal(dofs[0], V.cell_dofs)
    assert np.array_equal(dofs[1], np.where(Q.cell_dofs > 0, Q.cell_dofs + V.n_free, Q.cell_dofs - V.n_dirichlet))


# ---- CPU: oracle known answers ---------------------------------------------------------------------------------------
def test_oracle_jump_of_a_continuous_space_vanishes():
    """test/assembly_tests.jl:420-427: b = assemble_vector(∫_Λ jump(v)); @test sum(b)+1 ≈ 1 — and, entry by entry, the jump of a
    continuous basis function is zero, so b ≈ 0 and the jump-jump matrix holds (explicitly stored) zeros only"""
    mesh = H.cartesian_mesh((0, 1, 0, 1), (4, 4))
    _warp(mesh)
    V = H.lagrange_space(mesh, 1, [1, 3])
    bp = MF.skeleton_problem([V], 2)
    sides = [[(int(bp.side_cells[i, a]), int(bp.face_var[i, a])) for a in range(2)] for i in range(bp.face_nodes.shape[0])]
    b = O.assemble_vector_multifield(2, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                     _oracle_fields(bp, [V], False), lambda p: jump_v(p))
    assert b.shape == (V.n_free,) and sum(b) + 1 == pytest.approx(1.0, abs=1e-14) and np.abs(b).max() < 1e-15
    cp, rv, nz = _oracle_matrix(bp, [V], mesh, "jumpjump")
    assert cp.shape == (V.n_free + 1,) and np.abs(nz).max() < 1e-15
    # pattern: two free dofs are coupled iff they belong to two cells (possibly the same) that share an interior face
    pairs = set()
    for c0, c1 in bp.side_cells:
        d = [x for x in list(V.cell_dofs[c0 - 1]) + list(V.cell_dofs[c1 - 1]) if x > 0]
        pairs |= {(r, c) for r in d for c in d}
    got = {(int(rv[k]), c + 1) for c in range(V.n_free) for k in range(cp[c] - 1, cp[c + 1] - 1)}
    assert got == pairs


def test_oracle_product_space_blocks_equal_the_single_field_matrix():
    """a((u1,u2),(v1,v2)) = ∫ (v1+v2)*(u1+u2) over V × V: every block of the monolithic matrix is the mass matrix of V, bit for bit"""
    mesh = H.cartesian_mesh((0, 1, 0, 1), (3, 3))
    _warp(mesh)
    V = H.lagrange_space(mesh, 1, [2])
    bp = MF.volume_problem([V, V], 2)
    cp, rv, nz = _oracle_matrix(bp, [V, V], mesh, "sum_mass")
    tab = H.measure_tabulation(V, 2)
    cp1, rv1, nz1 = O.assemble_matrix(O.MASS, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet,
                                      dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM))
    n = V.n_free
    assert cp.shape == (2 * n + 1,)
    for cb in range(2):
        for c in range(n):
            lo, hi = cp[cb * n + c] - 1, cp[cb * n + c + 1] - 1
            lo1, hi1 = cp1[c] - 1, cp1[c + 1] - 1
            assert np.array_equal(rv[lo:hi], np.concatenate([rv1[lo1:hi1], rv1[lo1:hi1] + n]))
            assert np.array_equal(nz[lo:hi], np.concatenate([nz1[lo1:hi1], nz1[lo1:hi1]]))


def test_oracle_stokes_blocks():
    """K = vector Laplacian of V (bitwise the single-field assembly), the (p, v) block is minus the transpose of the (u, q)
    block, the (p, q) block is stored and zero (block_mask all true, assembly.jl:321-333)"""
    import scipy.sparse as sp
    mesh, (V, Q) = _stokes_spaces((2, 2))
    bp = MF.volume_problem([V, Q], 4)
    cp, rv, nz = _oracle_matrix(bp, [V, Q], mesh, "stokes")
    n, m = V.n_free, Q.n_free
    A = sp.csc_matrix((nz, rv - 1, cp - 1), shape=(n + m, n + m))
    tab = H.measure_tabulation(V, 4)
    cp1, rv1, nz1 = O.assemble_matrix(O.LAPLACE, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet,
                                      dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM), n_comp=2)
    K1 = sp.csc_matrix((nz1, rv1 - 1, cp1 - 1), shape=(n, n))
    assert abs(A[:n, :n] - K1).max() == 0.0
    B_uq = A[:n, n:].toarray()          # rows u, columns q:  ∫ q div(u)
    B_pv = A[n:, :n].toarray()          # rows p, columns v: -∫ div(v) p
    assert np.abs(B_uq).max() > 1e-3 and np.abs(B_pv + B_uq.T).max() < 1e-15
    assert abs(A[n:, n:]).max() == 0.0 and A[n:, n:].nnz > 0      # explicit zeros are part of the pattern
    # div of the constant-pressure mode integrates the velocity flux through the boundary: zero for interior (free) velocities
    assert np.abs(B_uq @ np.ones(m)).max() < 1e-13


def test_recognition_of_block_forms_without_a_gpu():
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (3, 3))
    Om = GT.interior(mesh)
    V = GT.lagrange_space(Om, 2, dirichlet_boundary=GT.boundary(mesh), tensor_size=(2,))
    Q = GT.lagrange_space(Om, 1)
    X = V * Q
    assert X.num_free_dofs() == V.num_free_dofs() + Q.num_free_dofs()
    dO = GT.measure(Om, 4)
    a = lambda up, vq: GT.integrate(lambda x: GT.dot(GT.grad(vq[0], x), GT.grad(up[0], x)) - GT.div(vq[0], x) * up[1](x)
                                    + vq[1](x) * GT.div(up[0], x), dO)
    term, meas, _ = a(GT._form_arguments(X, 2), GT._form_arguments(X, 1)).contributions[0]
    bp = GT._block_problem(X, meas)
    assert sorted(GT.recognise_blocks(term, bp, False)) == [(0, 0, E.BLOCK_LAPLACE, 1.0), (0, 1, E.BLOCK_DIVU_VALV, 1.0),
                                                            (1, 0, E.BLOCK_VALU_DIVV, -1.0)]
    dL = GT.measure(GT.skeleton(mesh), 2)
    W = GT.lagrange_space(Om, 1)
    jump = lambda u, p: u[2](p) - u[1](p)
    aj = lambda u, v: GT.integrate(lambda p: jump(u, p) * jump(v, p), dL)
    term, meas, _ = aj(GT._form_arguments(W, 2), GT._form_arguments(W, 1)).contributions[0]
    bs = GT._block_problem(W, meas)
    assert sorted(GT.recognise_blocks(term, bs, True)) == [(0, 0, E.BLOCK_MASS, 1.0), (0, 1, E.BLOCK_MASS, -1.0),
                                                           (1, 0, E.BLOCK_MASS, -1.0), (1, 1, E.BLOCK_MASS, 1.0)]
    with pytest.raises(GT.UnsupportedFormError):     # an unrestricted argument on a skeleton measure
        t, m, _ = GT.integrate(lambda p: GT.FormArgument(W, 2)(p) * GT.FormArgument(W, 1)(p), dL).contributions[0]
        GT.recognise_blocks(t, bs, True)
    with pytest.raises(GT.UnsupportedFormError):     # gradient of u times value of v: no such block kernel
        t, m, _ = GT.integrate(lambda p: GT.dot(GT.grad(GT.FormArgument(V, 2, 0), p), GT.FormArgument(V, 1, 0)(p)), dO).contributions[0]
        GT.recognise_blocks(t, bp, False)


# ---- GPU: block kernels against the oracle -----------------------------------------------------------------------------
def _engine(bp):
    eng = E.Engine(0)
    eng.set_mesh(bp.node_coordinates, bp.face_nodes)
    if bp.manifold_dim != bp.node_coordinates.shape[1]:
        eng.set_manifold_dim(bp.manifold_dim)
    eng.set_space(bp.super_dofs, bp.n_free, bp.n_dirichlet, 1)
    eng.set_parts(bp.w, bp.M, bp.dM, bp.parts, bp.n_sides, bp.face_var)
    return eng


def _check(eng, blocks, expect, fd=(E.FREE, E.FREE)):
    cp, rv, nz = expect
    eng.matrix_symbolic(*fd)
    gcp, grv = eng.matrix_pattern()
    assert np.array_equal(gcp, cp) and np.array_equal(grv, rv)          # bit-exact pattern, explicit zeros included
    gnz = eng.matrix_numeric_blocks(blocks)
    assert_values_close(gnz, nz)
    assert np.array_equal(eng.matrix_numeric_blocks(blocks), gnz)       # re-assembly: byte-identical
    return gnz


@pytest.mark.gpu
@pytest.mark.parametrize("cells", [(3, 2), (2, 2, 2)])
def test_gpu_stokes_blocks_parity(cells):
    mesh, spaces = _stokes_spaces(cells)
    bp = MF.volume_problem(spaces, 4)
    blocks = [(0, 0, E.BLOCK_LAPLACE, 1.0), (1, 0, E.BLOCK_VALU_DIVV, -1.0), (0, 1, E.BLOCK_DIVU_VALV, 1.0)]
    eng = _engine(bp)
    _check(eng, blocks, _oracle_matrix(bp, spaces, mesh, "stokes"))
    # free x Dirichlet columns of the same form (Ad of a linear problem, problems.jl:363-380)
    _check(eng, blocks, _oracle_matrix(bp, spaces, mesh, "stokes", fd=(O.FREE, O.DIRICHLET)), fd=(E.FREE, E.DIRICHLET))
    assert eng.lib.gtk_info(eng.h, 5) == 7
    eng.close()


@pytest.mark.gpu
def test_gpu_product_space_mass_parity():
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (3, 2, 2), simplexify=True)
    V = H.lagrange_space(mesh, 2, [1])
    W = H.lagrange_space(mesh, 1, [2, 3])
    bp = MF.volume_problem([V, W], 4)
    eng = _engine(bp)
    _check(eng, [(pu, pv, E.BLOCK_MASS, 0.5) for pu in range(2) for pv in range(2)],
           _oracle_matrix(bp, [V, W], mesh, "sum_mass", alpha=0.5))
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("cells,order", [((4, 4), 1), ((3, 2), 2), ((3, 2, 2), 1), ((2, 2, 2), 2)])
def test_gpu_skeleton_jump_parity(cells, order):
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells)
    _warp(mesh)
    V = H.lagrange_space(mesh, order, [1, 3])
    bp = MF.skeleton_problem([V], 2 * order)
    eng = _engine(bp)
    blocks = [(0, 0, E.BLOCK_MASS, 1.0), (0, 1, E.BLOCK_MASS, -1.0), (1, 0, E.BLOCK_MASS, -1.0), (1, 1, E.BLOCK_MASS, 1.0)]
    cp, rv, nz = _oracle_matrix(bp, [V], mesh, "jumpjump")
    eng.matrix_symbolic()
    gcp, grv = eng.matrix_pattern()
    assert np.array_equal(gcp, cp) and np.array_equal(grv, rv)
    gnz = eng.matrix_numeric_blocks(blocks)
    # a continuous space has no jumps: the values are rounding-level zeros on both sides; compare against the entry scale h^(D-1)
    scale = (1.0 / max(cells)) ** (D - 1)
    assert np.abs(gnz - nz).max() < 1e-12 * scale and np.abs(gnz).max() < 1e-12 * scale
    # one side only: a non-trivial matrix (∫_Λ u[1] v[1]), values to 1e-12
    one = eng.matrix_numeric_blocks([(0, 0, E.BLOCK_MASS, 1.0)])
    sides = [[(int(bp.side_cells[i, a]), int(bp.face_var[i, a])) for a in range(2)] for i in range(bp.face_nodes.shape[0])]
    ref = O.assemble_matrix_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                       _oracle_fields(bp, [V], False), lambda p: p.u(0, 1) * p.v(0, 1))
    assert np.array_equal(ref[0], gcp) and np.array_equal(ref[1], grv)
    assert_values_close(one, ref[2])
    # ∫_Λ jump(v): zero vector; ∫_Λ v[2]: against the oracle
    b = eng.vector_assemble_blocks([(0, -1.0, 1.0), (1, 1.0, 1.0)])
    assert b.shape == (V.n_free,) and np.abs(b).max() < 1e-12 * scale
    b2 = eng.vector_assemble_blocks([(1, 1.0, 1.0)])
    rb = O.assemble_vector_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                      _oracle_fields(bp, [V], False), lambda p: p.v(0, 2))
    assert_values_close(b2, rb)
    eng.close()


@pytest.mark.gpu
def test_gpu_skeleton_two_fields_parity():
    """test/assembly_tests.jl:407-416: V² = V × V, ∫_Λ jump(u1) jump(v2): one non-zero field block, all of them stored"""
    mesh = H.cartesian_mesh((0, 1, 0, 1), (4, 3))
    _warp(mesh)
    V = H.lagrange_space(mesh, 1, [1, 3])
    bp = MF.skeleton_problem([V, V], 2)
    eng = _engine(bp)
    pu = [bp.part_index(0, 0), bp.part_index(0, 1)]
    pv = [bp.part_index(1, 0), bp.part_index(1, 1)]
    blocks = [(pu[a], pv[b], E.BLOCK_MASS, (-1.0 if a == 0 else 1.0) * (-1.0 if b == 0 else 1.0)) for a in range(2) for b in range(2)]
    cp, rv, nz = _oracle_matrix(bp, [V, V], mesh, "jump12")
    eng.matrix_symbolic()
    gcp, grv = eng.matrix_pattern()
    assert np.array_equal(gcp, cp) and np.array_equal(grv, rv) and gcp.shape == (2 * V.n_free + 1,)
    gnz = eng.matrix_numeric_blocks(blocks)
    assert np.abs(gnz - nz).max() < 1e-13
    eng.close()


@pytest.mark.gpu
def test_gpu_reference_tests_through_the_gt_mirror():
    """test/assembly_tests.jl:366-427 transcribed: jump-jump matrix, V² = V × V products, ∫_Λ jump(v)"""
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (4, 4))
    Om, Lam = GT.interior(mesh), GT.skeleton(mesh)
    Gdiri = GT.boundary(mesh, group_names=["1-face-1", "1-face-3"])
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=Gdiri)
    dO, dL = GT.measure(Om, 2), GT.measure(Lam, 2)
    jump = lambda u, p: u[2](p) - u[1](p)
    A = GT.assemble_matrix(lambda u, v: GT.integrate(lambda p: jump(u, p) * jump(v, p), dL), float, V, V)
    assert A.m == V.num_free_dofs() and A.n == V.num_free_dofs() and np.abs(A.nzval).max() < 1e-14
    V2 = V * V
    a2 = lambda u, v: GT.integrate(lambda q: (v[0](q) + v[1](q)) * (u[0](q) + u[1](q)), dO)
    A2, cache = GT.assemble_matrix(a2, float, V2, V2, reuse=True)
    assert A2.m == 2 * V.num_free_dofs()
    M = GT.assemble_matrix(lambda u, v: GT.integrate(lambda q: u(q) * v(q), dO), float, V, V)
    S = A2.to_scipy()
    n = V.num_free_dofs()
    for i in range(2):
        for j in range(2):
            assert abs(S[i * n:(i + 1) * n, j * n:(j + 1) * n] - M.to_scipy()).max() < 1e-15
    before = A2.nzval.copy()
    GT.update_matrix(A2, cache)
    assert np.array_equal(before, A2.nzval)
    cache.engine.close()
    b = GT.assemble_vector(lambda v: GT.integrate(lambda p: jump(v, p), dL), float, V)
    assert b.shape == (n,) and sum(b) + 1 == pytest.approx(1.0, abs=1e-13)
    with pytest.raises(GT.UnsupportedFormError):      # no CPU fallback for what the block kernels do not cover
        GT.assemble_matrix(lambda u, v: GT.integrate(lambda p: GT.dot(GT.grad(u[1], p), GT.grad(v[2], p)), dL), float, V, V)


@pytest.mark.gpu
def test_gpu_stokes_lid_driven_cavity_solves():
    """docs/src/src_jl/example_stokes.jl at 6 x 6: the assembled saddle-point system is solvable (one pressure dof pinned) and the
    discrete velocity is divergence-free against every pressure test function"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (6, 6))
    Om = GT.interior(mesh)
    V = GT.lagrange_space(Om, 2, dirichlet_boundary=GT.boundary(mesh), tensor_size=(2,))
    Q = GT.lagrange_space(Om, 1)
    X = V * Q
    dO = GT.measure(Om, 4)
    a = lambda up, vq: GT.integrate(lambda x: GT.dot(GT.grad(vq[0], x), GT.grad(up[0], x)) - GT.div(vq[0], x) * up[1](x)
                                    + vq[1](x) * GT.div(up[0], x), dO)
    A = GT.assemble_matrix(a, float, X, X)
    Ad = GT.assemble_matrix(a, float, X, X, free_or_dirichlet=(GT.FREE, GT.DIRICHLET))
    n, m = V.num_free_dofs(), Q.num_free_dofs()
    assert A.m == n + m and Ad.n == V.num_dirichlet_dofs()
    # lid velocity (1, 0) on the side y = 1
    xd = np.zeros(V.num_dirichlet_dofs())
    Xd = V.data.dirichlet_dof_nodes
    comp = GT._dof_component(V, False)
    xd[(np.abs(Xd[:, 1] - 1.0) < 1e-12) & (comp == 0)] = 1.0
    S = A.to_scipy().tolil()
    rhs = -(Ad.to_scipy() @ xd)
    S[n, :] = 0.0; S[:, n] = 0.0; S[n, n] = 1.0; rhs[n] = 0.0          # pin one pressure dof (GT.last_dof() in the example)
    x = spla.spsolve(sp.csc_matrix(S), rhs)
    assert np.isfinite(x).all() and np.abs(x[:n]).max() > 1e-2
    full = A.to_scipy() @ x + Ad.to_scipy() @ xd
    assert np.abs(np.delete(full, n)).max() < 1e-10


# ---- sums of integrals over different domains in ONE matrix (gtk_matrix_sum_*) ------------------------------------------
def _volume_coo(form, mesh, V, degree, alpha=1.0):
    tab = H.measure_tabulation(V, degree)
    be = O.element_matrices(form, mesh.node_coordinates, mesh.cell_nodes, dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM),
                            n_comp=V.n_comp, alpha=alpha)
    return O.coo_matrix(be, V.cell_dofs, V.cell_dofs)


@pytest.mark.gpu
@pytest.mark.parametrize("cells", [(4, 4), (3, 2, 2)])
def test_gpu_sum_of_integrals_over_three_domains(cells):
    """test/assembly_tests.jl:366-403: a(u,v) = ∫_Ω u v + ∫_Γ u v + ∫_Λ jump(u) jump(v) assembled as ONE matrix: the union
    pattern (the skeleton term couples all dofs of two neighbouring cells) and the sum of all triplets"""
    D = len(cells)
    mesh = GT.cartesian_mesh(tuple([0, 1] * D), cells)
    _warp(mesh)
    Om, Lam = GT.interior(mesh), GT.skeleton(mesh)
    names = ["1-face-1", "1-face-3"] if D == 2 else ["2-face-1", "2-face-3"]
    Gam = GT.boundary(mesh, group_names=["1-face-2", "1-face-4"] if D == 2 else ["2-face-2", "2-face-6"])
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=GT.boundary(mesh, group_names=names))
    dO, dG, dL = GT.measure(Om, 2), GT.measure(Gam, 2), GT.measure(Lam, 2)
    jump = lambda u, p: u[2](p) - u[1](p)
    a = lambda u, v: (GT.integrate(lambda q: u(q) * v(q), dO) + 0.5 * GT.integrate(lambda q: u(q) * v(q), dG)
                      + GT.integrate(lambda p: jump(u, p) * jump(v, p), dL))
    A, cache = GT.assemble_matrix(a, float, V, V, reuse=True)
    # the oracle: triplets of the three contributions, concatenated, one compress
    Vd = V.data
    coo = [_volume_coo(O.MASS, mesh, Vd, 2)]
    fp = H.face_problem(Vd, [2, 4] if D == 2 else [2, 6], 2)
    bef = O.element_matrices(O.MASS, mesh.node_coordinates, fp.face_nodes, dict(w=fp.tab.w, N=fp.tab.N, dN=fp.tab.dN, M=fp.tab.M, dM=fp.tab.dM), alpha=0.5)
    coo.append(O.coo_matrix(bef, fp.face_dofs, fp.face_dofs))
    bp = MF.skeleton_problem([Vd], 2)
    sides = [[(int(bp.side_cells[i, s]), int(bp.face_var[i, s])) for s in range(2)] for i in range(bp.face_nodes.shape[0])]
    coo.append(O.assemble_matrix_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                            _oracle_fields(bp, [Vd], False), INTEGRANDS["jumpjump"], return_coo=True))
    cp, rv, nz = O.assemble_matrix_sum(coo, Vd.n_free, Vd.n_free)
    assert np.array_equal(A.colptr, cp) and np.array_equal(A.rowval, rv)
    assert_values_close(A.nzval, nz)
    # wider than the volume pattern alone
    M = GT.assemble_matrix(lambda u, v: GT.integrate(lambda q: u(q) * v(q), dO), float, V, V)
    assert A.nzval.size > M.nzval.size
    before = A.nzval.copy()
    A.nzval[:] = 0.0
    GT.update_matrix(A, cache)                    # update_matrix!: every integral re-assembled, merged again
    assert np.array_equal(before, A.nzval)
    for e, _ in cache.params["parts"]:
        e.close()
    cache.engine.close()


@pytest.mark.gpu
def test_gpu_poisson_with_robin_term_as_one_matrix():
    """a(u,v) = ∫_Ω ∇u⋅∇v + ∫_Γ u v on 3D Q1 hexahedra: the volume part runs the structured sweep kernel, the Robin part the
    boundary-face kernel, merged on the device; pattern = the volume pattern"""
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (6, 5, 4))
    Om = GT.interior(mesh)
    Gam = GT.boundary(mesh, group_names=["2-face-2", "2-face-4"])
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=GT.boundary(mesh, group_names=["2-face-1"]))
    dO, dG = GT.measure(Om, 2), GT.measure(Gam, 2)
    a = lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dO) + 3.0 * GT.integrate(lambda x: u(x) * v(x), dG)
    A = GT.assemble_matrix(a, float, V, V)
    Vd = V.data
    fp = H.face_problem(Vd, [2, 4], 2)
    bef = O.element_matrices(O.MASS, mesh.node_coordinates, fp.face_nodes, dict(w=fp.tab.w, N=fp.tab.N, dN=fp.tab.dN, M=fp.tab.M, dM=fp.tab.dM), alpha=3.0)
    cp, rv, nz = O.assemble_matrix_sum([_volume_coo(O.LAPLACE, mesh, Vd, 2), O.coo_matrix(bef, fp.face_dofs, fp.face_dofs)], Vd.n_free, Vd.n_free)
    assert np.array_equal(A.colptr, cp) and np.array_equal(A.rowval, rv)
    assert_values_close(A.nzval, nz)
    K = GT.assemble_matrix(lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dO), float, V, V)
    assert np.array_equal(K.colptr, A.colptr) and np.array_equal(K.rowval, A.rowval)


# ---- gradients and unit normals on skeleton faces: interior-penalty terms (GTK_BLOCK_IP) ---------------------------------
def _ip(p, su, sv, c):
    """c0 ((1/h) v n_sv)⋅(u n_su) + c1 (v n_sv)⋅∇u + c2 ∇v⋅(u n_su) with u on the cell around su, v on sv (masked elsewhere)"""
    fv, fu = p.v(0, sv), p.u(0, su)
    return (O.frobenius((c[0] / p.h) * (fv * p.n(sv)), fu * p.n(su)) + c[1] * O.frobenius(fv * p.n(sv), p.grad_u(0, su))
            + c[2] * O.frobenius(p.grad_v(0, sv), fu * p.n(su)))


def _skeleton_oracle(bp, V, mesh, integrand, **kw):
    sides = [[(int(bp.side_cells[i, a]), int(bp.face_var[i, a])) for a in range(2)] for i in range(bp.face_nodes.shape[0])]
    return O.assemble_matrix_multifield(mesh.D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                        _oracle_fields(bp, [V], True), integrand,
                                        skeleton_geometry=(bp.cell_nodes, bp.dM_cell, bp.ref_normals), **kw)


@pytest.mark.parametrize("cells", [(3, 3), (2, 2, 2)])
def test_unit_normals_of_the_two_cells_around_a_face_are_opposite(cells):
    """map_unit_normal (accessors.jl:1026-1035) at the mapped face points: unit length, n[1] = -n[2], n[1] points from the first
    cell around into the second — on a warped mesh, through the tables multifield.py hands to the engine"""
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells)
    _warp(mesh)
    V = H.lagrange_space(mesh, 1, None)
    bp = MF.skeleton_problem([V], 2, gradients=True)
    X = mesh.node_coordinates
    for i in range(bp.face_nodes.shape[0]):
        c = [X[mesh.cell_nodes[bp.side_cells[i, a] - 1] - 1].mean(axis=0) for a in range(2)]
        for q in range(bp.w.size):
            n = []
            for a in range(2):
                var = bp.face_var[i, a]
                J = O.point_geometry(X, mesh.cell_nodes[bp.side_cells[i, a] - 1][None, :], bp.dM_cell[var][q])[0]
                n.append(O.map_unit_normal(J, bp.ref_normals[var]))
            assert abs(np.linalg.norm(n[0]) - 1.0) < 1e-14 and np.abs(n[0] + n[1]).max() < 1e-13
            assert np.dot(n[0], c[1] - c[0]) > 0.0


def test_oracle_interior_penalty_terms_cancel_on_a_continuous_space():
    """test/assembly_tests.jl:329-340: 'The skeleton terms are not needed. They are added just to make sure that they are computed
    correctly' — on a continuous space jump(u, n) of every basis function vanishes, so the assembled term is zero to rounding"""
    mesh = H.cartesian_mesh((0, 1, 0, 1), (3, 3))
    _warp(mesh)
    V = H.lagrange_space(mesh, 1, "boundary")
    bp = MF.skeleton_problem([V], 2, gradients=True)
    full = lambda p: sum(_ip(p, su, sv, (1.0, -0.5, -0.5)) for su in (1, 2) for sv in (1, 2))
    cp, rv, nz = _skeleton_oracle(bp, V, mesh, full)
    one = _skeleton_oracle(bp, V, mesh, lambda p: _ip(p, 1, 1, (1.0, -0.5, -0.5)))
    assert np.abs(one[2]).max() > 0.1 and np.abs(nz).max() < 1e-14


@pytest.mark.gpu
@pytest.mark.parametrize("cells,order", [((4, 3), 1), ((3, 2), 2), ((2, 2, 2), 1)])
def test_gpu_interior_penalty_blocks_parity(cells, order):
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells)
    _warp(mesh)
    V = H.lagrange_space(mesh, order, [1])
    bp = MF.skeleton_problem([V], 2 * order, gradients=True)
    eng = _engine(bp)
    eng.set_skeleton_cells(bp.cell_nodes, bp.side_cells, bp.dM_cell, bp.ref_normals)
    eng.matrix_symbolic()
    gcp, grv = eng.matrix_pattern()
    # two one-sided blocks with different coefficients: nothing cancels, every term is exercised
    ca, cb = (2.0, -0.5, 0.25), (1.0, 0.3, -0.5)
    ref = _skeleton_oracle(bp, V, mesh, lambda p: 1.0 * _ip(p, 1, 1, ca) + 0.7 * _ip(p, 2, 1, cb))
    assert np.array_equal(ref[0], gcp) and np.array_equal(ref[1], grv)
    got = eng.matrix_numeric_blocks([(0, 0, E.BLOCK_IP, 1.0, ca), (1, 0, E.BLOCK_IP, 0.7, cb)])
    assert_values_close(got, ref[2])
    assert np.array_equal(eng.matrix_numeric_blocks([(0, 0, E.BLOCK_IP, 1.0, ca), (1, 0, E.BLOCK_IP, 0.7, cb)]), got)
    # the full interior-penalty term on this continuous space: zero to rounding
    full = eng.matrix_numeric_blocks([(pu, pv, E.BLOCK_IP, 1.0, (1.0, -0.5, -0.5)) for pu in range(2) for pv in range(2)])
    assert np.abs(full).max() < 1e-12 * np.abs(got).max()
    with pytest.raises(E.UnsupportedFormError):          # IP needs scalar parts with gradient tables
        eng2 = _engine(MF.skeleton_problem([V], 2 * order))
        eng2.matrix_symbolic()
        eng2.matrix_numeric_blocks([(0, 0, E.BLOCK_IP, 1.0, ca)])
    eng.close()


@pytest.mark.gpu
def test_gpu_reference_poisson_with_skeleton_terms():
    """test/assembly_tests.jl:311-360 transcribed: Poisson with Dirichlet data u = x + y, the interior-penalty skeleton terms added
    to the form 'just to make sure that they are computed correctly'; the discrete solution is the exact one (tol 1e-10)"""
    import scipy.sparse.linalg as spla
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (4, 4))
    Om, Lam, Gd = GT.interior(mesh), GT.skeleton(mesh), GT.boundary(mesh)
    V = GT.lagrange_space(Om, 1, dirichlet_boundary=Gd)
    dO, dL = GT.measure(Om, 2), GT.measure(Lam, 2)
    n = GT.unit_normal(mesh, 1)
    h = GT.face_diameter_field(Lam)
    gamma = GT.uniform_quantity(1.0)
    jump = lambda u, n_, x: u[2](x) * n_[2](x) + u[1](x) * n_[1](x)
    mean = lambda f, u, x: 0.5 * (f(u[1], x) + f(u[2], x))
    a = lambda u, v: (GT.integrate(lambda q: GT.dot(GT.grad(u, q), GT.grad(v, q)), dO)
                      + GT.integrate(lambda x: GT.dot((gamma / h(x)) * jump(v, n, x), jump(u, n, x)) - GT.dot(jump(v, n, x), mean(GT.grad, u, x))
                                     - GT.dot(mean(GT.grad, v, x), jump(u, n, x)), dL))
    A = GT.assemble_matrix(a, float, V, V)
    Ad = GT.assemble_matrix(a, float, V, V, free_or_dirichlet=(GT.FREE, GT.DIRICHLET))
    xd = V.data.dirichlet_dof_nodes.sum(axis=1)                     # interpolate_dirichlet!(u, uhd), u = sum
    x = spla.spsolve(A.to_scipy().tocsc(), -(Ad.to_scipy() @ xd))   # l(v) = 0
    assert np.abs(x - V.data.free_dof_nodes.sum(axis=1)).max() < 1e-10
    # the skeleton term alone widens the pattern and contributes nothing
    K = GT.assemble_matrix(lambda u, v: GT.integrate(lambda q: GT.dot(GT.grad(u, q), GT.grad(v, q)), dO), float, V, V)
    assert A.nzval.size > K.nzval.size and abs(A.to_scipy() - K.to_scipy()).max() < 1e-12


# ---- discontinuous spaces, Nitsche terms on boundary faces: the reference's interior-penalty example -------------------
def _boundary_oracle_sides(bp):
    return [[(int(bp.side_cells[i, 0]), int(bp.face_var[i, 0]))] for i in range(bp.face_nodes.shape[0])]


def test_discontinuous_space_numbering_and_boundary_inputs():
    """lagrange_space(Ω, k; continuous=false): dof = (cell-1) n_ldofs + local dof (space.jl:860-882); boundary faces carry the one
    cell around with an outward normal"""
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (3, 2, 2))
    _warp(mesh)
    V = H.discontinuous_lagrange_space(mesh, 1)
    assert V.n_free == mesh.n_cells * 8 and V.n_dirichlet == 0
    assert np.array_equal(V.cell_dofs, np.arange(1, mesh.n_cells * 8 + 1).reshape(-1, 8))
    assert np.allclose(V.free_dof_nodes, mesh.node_coordinates[mesh.cell_nodes.astype(np.int64) - 1].reshape(-1, 3))
    bp = MF.boundary_problem([V], None, 2)
    assert bp.face_nodes.shape[0] == 2 * (3 * 2 + 3 * 2 + 2 * 2) and bp.n_sides == 1
    X = mesh.node_coordinates
    centre = X.mean(axis=0)
    for i in range(bp.face_nodes.shape[0]):
        cell, var = int(bp.side_cells[i, 0]), int(bp.face_var[i, 0])
        J = O.point_geometry(X, mesh.cell_nodes[cell - 1][None, :], bp.dM_cell[var][0])[0]
        n = O.map_unit_normal(J, bp.ref_normals[var])
        fc = X[bp.face_nodes[i] - 1].mean(axis=0)
        assert abs(np.linalg.norm(n) - 1) < 1e-14 and np.dot(n, fc - centre) > 0.0      # outward


@pytest.mark.parametrize("cells", [(3, 2), (2, 2, 2)])
def test_boundary_inputs_on_simplices(cells):
    """the simplex around a boundary face comes from the face complex (boundary_faces names the parent hexahedron); its unit
    normal points out of the domain"""
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells, simplexify=True)
    _warp(mesh)
    V = H.discontinuous_lagrange_space(mesh, 1)
    bp = MF.boundary_problem([V], None, 2)
    X = mesh.node_coordinates
    centre = X.mean(axis=0)
    assert bp.face_nodes.shape[0] == H.boundary_faces(mesh, None)[0].shape[0]
    for i in range(bp.face_nodes.shape[0]):
        cell, var = int(bp.side_cells[i, 0]), int(bp.face_var[i, 0])
        assert set(bp.face_nodes[i]) <= set(mesh.cell_nodes[cell - 1])
        J = O.point_geometry(X, mesh.cell_nodes[cell - 1][None, :], bp.dM_cell[var][0])[0]
        n = O.map_unit_normal(J, bp.ref_normals[var])
        assert abs(np.linalg.norm(n) - 1) < 1e-14 and np.dot(n, X[bp.face_nodes[i] - 1].mean(axis=0) - centre) > 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("cells,simplexify", [((3, 3), False), ((2, 2, 2), False), ((3, 2), True), ((2, 2, 2), True)])
def test_gpu_nitsche_terms_parity(cells, simplexify):
    """(γ/h) v u - v n⋅∇u - n⋅∇v u on boundary faces and the right-hand side (γ/h) v g - n⋅∇v g, on a discontinuous space
    (quadrilaterals / hexahedra and triangles / tetrahedra)"""
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells, simplexify=simplexify)
    _warp(mesh)
    V = H.discontinuous_lagrange_space(mesh, 1)
    sides_sel = [1, 4] if D == 2 else [2, 5]
    bp = MF.boundary_problem([V], sides_sel, 2)
    eng = _engine(bp)
    eng.set_skeleton_cells(bp.cell_nodes, bp.side_cells, bp.dM_cell, bp.ref_normals)
    fields_o = _oracle_fields(bp, [V], True)
    geo = (bp.cell_nodes, bp.dM_cell, bp.ref_normals)
    sides = _boundary_oracle_sides(bp)
    gamma = 0.7
    mat = lambda p: (gamma / p.h) * p.v(0) * p.u(0) - O.frobenius(p.v(0) * p.n(1), p.grad_u(0)) - O.frobenius(p.n(1), p.grad_v(0)) * p.u(0)
    ref = O.assemble_matrix_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides, fields_o, mat,
                                       skeleton_geometry=geo)
    eng.matrix_symbolic()
    gcp, grv = eng.matrix_pattern()
    assert np.array_equal(ref[0], gcp) and np.array_equal(ref[1], grv)
    got = eng.matrix_numeric_blocks([(0, 0, E.BLOCK_IP, 1.0, (gamma, -1.0, -1.0))])
    assert_values_close(got, ref[2])
    xq = MF.face_point_coordinates(bp)
    g = np.sin(xq[..., 0]) + 2.0 * xq[..., 1]
    vec = lambda p: (gamma / p.h) * p.v(0) * p.g - O.frobenius(p.n(1), p.grad_v(0)) * p.g
    rb = O.assemble_vector_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides, fields_o, vec,
                                      skeleton_geometry=geo, point_data=g)
    b = eng.vector_assemble_blocks([(0, 1.0, (0.0, gamma, -1.0))], g_qp=g)
    assert_values_close(b, rb)
    b2 = eng.vector_assemble_blocks([(0, 1.0, (0.0, gamma, -1.0))], g_qp=g, accumulate=True)     # continues the same COO vector
    assert_values_close(b2, 2.0 * rb)
    eng.close()


@pytest.mark.gpu
def test_gpu_reference_interior_penalty_example():
    """docs/src/src_jl/example_hello_world_dg.jl transcribed: symmetric interior penalty on a DISCONTINUOUS Q1 space, 4 x 4 x 4
    hexahedra — Laplace operator on Ω + interior penalty on Λ + Nitsche terms on Γ in ONE matrix, right-hand side with the
    Nitsche data terms; the example's own check: the L2 error against g = sum(x) is below 1e-9"""
    import scipy.sparse.linalg as spla
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (4, 4, 4))
    D = 3
    n = GT.unit_normal(mesh, D - 1)
    Om, Gd, Lam = GT.interior(mesh), GT.boundary(mesh), GT.skeleton(mesh)
    h_L, h_G = GT.face_diameter_field(Lam), GT.face_diameter_field(Gd)
    g = GT.AnalyticalField(lambda x: x[0] + x[1] + x[2], Om)
    f = GT.AnalyticalField(lambda x: 0.0 * x[0], Om)
    mean = lambda fn, u, x: 0.5 * (fn(u[1], x) + fn(u[2], x))
    jump = lambda u, n_, x: u[2](x) * n_[2](x) + u[1](x) * n_[1](x)
    k = 1
    gamma = GT.uniform_quantity(k * (k + 1) / 10)
    V = GT.lagrange_space(Om, k, continuous=False)
    dO, dL, dG = GT.measure(Om, 2 * k), GT.measure(Lam, 2 * k), GT.measure(Gd, 2 * k)
    grad, dot = GT.grad, GT.dot
    a = lambda u, v: (GT.integrate(lambda x: dot(grad(u, x), grad(v, x)), dO)
                      + GT.integrate(lambda x: dot((gamma / h_L(x)) * jump(v, n, x), jump(u, n, x)) - dot(jump(v, n, x), mean(grad, u, x))
                                     - dot(mean(grad, v, x), jump(u, n, x)), dL)
                      + GT.integrate(lambda x: (gamma / h_G(x)) * v(x) * u(x) - dot(v(x) * n(x), grad(u, x)) - dot(n(x), grad(v, x)) * u(x), dG))
    l = lambda v: (GT.integrate(lambda x: v(x) * f(x), dO)
                   + GT.integrate(lambda x: (gamma / h_G(x)) * v(x) * g(x) - dot(n(x), grad(v, x)) * g(x), dG))
    A = GT.assemble_matrix(a, float, V, V)
    b = GT.assemble_vector(l, float, V)
    assert A.m == A.n == 64 * 8 == b.size
    S = A.to_scipy().tocsc()
    assert abs(S - S.T).max() < 1e-12 * abs(S).max()                 # the symmetric interior penalty method
    x = spla.spsolve(S, b)
    Xdof = V.data.free_dof_nodes
    assert np.abs(x - Xdof.sum(axis=1)).max() < 1e-9                 # nodal error of the exact (linear) solution
    uh = GT.solution_field(V, x)
    el2 = np.sqrt(GT.integrate(lambda y: GT.abs2(uh(y) - g(y)), dO).sum())
    assert el2 < 1.0e-9                                              # the example's own assertion


def test_oracle_reproduces_the_reference_interior_penalty_example():
    """docs/src/src_jl/example_hello_world_dg.jl on the CPU ORACLE (3 x 3 x 3 cells; the example's assertion holds on any mesh since
    g = sum(x) is in the space): Laplace operator + interior penalty + Nitsche terms pushed as ONE COO allocation, Nitsche data
    on the right-hand side; the discrete solution is g — `@assert el2 < 1.0e-9`.  A second known answer of the reference (after the
    p-Laplacian norm) that pins the restatement: normals, gradients on faces, face diameters, discontinuous numbering."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    cells = (3, 3, 3)
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), cells)
    V = H.discontinuous_lagrange_space(mesh, 1)
    gamma = 1 * (1 + 1) / 10
    coo = [_volume_coo(O.LAPLACE, mesh, V, 2)]
    bs = MF.skeleton_problem([V], 2, gradients=True)
    sides = [[(int(bs.side_cells[i, a]), int(bs.face_var[i, a])) for a in range(2)] for i in range(bs.face_nodes.shape[0])]
    jn = lambda p, w: (p.u(0, 2) * p.n(2) + p.u(0, 1) * p.n(1)) if w == "u" else (p.v(0, 2) * p.n(2) + p.v(0, 1) * p.n(1))
    mg = lambda p, w: 0.5 * ((p.grad_u(0, 1) + p.grad_u(0, 2)) if w == "u" else (p.grad_v(0, 1) + p.grad_v(0, 2)))
    ip = lambda p: O.frobenius((gamma / p.h) * jn(p, "v"), jn(p, "u")) - O.frobenius(jn(p, "v"), mg(p, "u")) - O.frobenius(mg(p, "v"), jn(p, "u"))
    coo.append(O.assemble_matrix_multifield(3, mesh.node_coordinates, bs.face_nodes, dict(w=bs.w, dM=bs.dM), sides, _oracle_fields(bs, [V], True), ip,
                                            skeleton_geometry=(bs.cell_nodes, bs.dM_cell, bs.ref_normals), return_coo=True))
    bb = MF.boundary_problem([V], None, 2)
    bsides = _boundary_oracle_sides(bb)
    geo = (bb.cell_nodes, bb.dM_cell, bb.ref_normals)
    nit = lambda p: (gamma / p.h) * p.v(0) * p.u(0) - O.frobenius(p.v(0) * p.n(1), p.grad_u(0)) - O.frobenius(p.n(1), p.grad_v(0)) * p.u(0)
    coo.append(O.assemble_matrix_multifield(3, mesh.node_coordinates, bb.face_nodes, dict(w=bb.w, dM=bb.dM), bsides, _oracle_fields(bb, [V], True), nit,
                                            skeleton_geometry=geo, return_coo=True))
    cp, rv, nz = O.assemble_matrix_sum(coo, V.n_free, V.n_free)
    gq = MF.face_point_coordinates(bb).sum(axis=2)
    rhs = lambda p: (gamma / p.h) * p.v(0) * p.g - O.frobenius(p.n(1), p.grad_v(0)) * p.g
    b = O.assemble_vector_multifield(3, mesh.node_coordinates, bb.face_nodes, dict(w=bb.w, dM=bb.dM), bsides, _oracle_fields(bb, [V], True), rhs,
                                     skeleton_geometry=geo, point_data=gq)
    A = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
    assert abs(A - A.T).max() < 1e-13
    x = spla.spsolve(A, b)
    err = x - V.free_dof_nodes.sum(axis=1)
    assert np.abs(err).max() < 1e-9
    # el2 = sqrt(∫ (uh - g)^2): with the nodal error e the integrand is the P/Q1 interpolant of e — bound it by max|e| sqrt(|Ω|)
    assert np.abs(err).max() * 1.0 < 1e-9


# ---- simplexified meshes: skeleton inputs and face terms on triangles / tetrahedra -------------------------------------------
@pytest.mark.parametrize("cells", [(3, 2), (2, 2, 2)])
def test_skeleton_inputs_on_simplices_equal_the_literal_restatement(cells):
    D = len(cells)
    dom = tuple([0, 1] * D)
    mesh = H.cartesian_mesh(dom, cells, simplexify=True)
    V = H.lagrange_space(mesh, 2, [1])
    bp = MF.skeleton_problem([V], 4, gradients=True)
    coords, cn, fc = O.simplex_face_complex(dom, cells)
    sf = O.skeleton_faces(cn, coords.shape[0], D, fc=fc, simplex=True)
    assert len(sf) == bp.face_nodes.shape[0] > 0
    q = H.quadrature(D - 1, True, 4)
    tabs = O.reference_map_tables(D, q.coordinates, simplex=True)
    refn = {2: [(0.0, -1.0), (-1.0, 0.0), (2 ** -0.5, 2 ** -0.5)],
            3: [(0.0, 0.0, -1.0), (0.0, -1.0, 0.0), (-1.0, 0.0, 0.0), (3 ** -0.5, 3 ** -0.5, 3 ** -0.5)]}[D]      # domain.jl:428, 460
    for i, f in enumerate(sf):
        assert list(bp.face_nodes[i]) == f["nodes"] and [s[0] for s in f["sides"]] == list(bp.side_cells[i])
        for a, (cell, lface, perm) in enumerate(f["sides"]):
            var = bp.face_var[i, a]
            No, _ = O.tabulate(D, 2, "P", np.array(tabs[lface - 1][perm - 1]))
            assert np.abs(No - bp.parts[a]["N"][var]).max() < 1e-13
            assert np.allclose(bp.ref_normals[var], refn[lface - 1], atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("cells,order", [((3, 3), 1), ((2, 2, 2), 2)])
def test_gpu_face_terms_on_simplices_parity(cells, order):
    """jump products and interior-penalty blocks on triangles / tetrahedra (P1 / P2): same kernels, simplex tables"""
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells, simplexify=True)
    _warp(mesh)
    V = H.lagrange_space(mesh, order, [1])
    bp = MF.skeleton_problem([V], 2 * order, gradients=True)
    eng = _engine(bp)
    eng.set_skeleton_cells(bp.cell_nodes, bp.side_cells, bp.dM_cell, bp.ref_normals)
    eng.matrix_symbolic()
    gcp, grv = eng.matrix_pattern()
    ca, cb = (2.0, -0.5, 0.25), (1.0, 0.3, -0.5)
    ref = _skeleton_oracle(bp, V, mesh, lambda p: 1.0 * _ip(p, 1, 1, ca) + 0.7 * _ip(p, 2, 1, cb))
    assert np.array_equal(ref[0], gcp) and np.array_equal(ref[1], grv)
    got = eng.matrix_numeric_blocks([(0, 0, E.BLOCK_IP, 1.0, ca), (1, 0, E.BLOCK_IP, 0.7, cb)])
    assert_values_close(got, ref[2])
    full = eng.matrix_numeric_blocks([(pu, pv, E.BLOCK_IP, 1.0, (1.0, -0.5, -0.5)) for pu in range(2) for pv in range(2)])
    assert np.abs(full).max() < 1e-12 * np.abs(got).max()             # continuous space: the four blocks cancel
    one = eng.matrix_numeric_blocks([(0, 1, E.BLOCK_MASS, 1.0)])      # ∫_Λ u[1] v[2]
    sides = [[(int(bp.side_cells[i, a]), int(bp.face_var[i, a])) for a in range(2)] for i in range(bp.face_nodes.shape[0])]
    r1 = O.assemble_matrix_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), sides,
                                      _oracle_fields(bp, [V], False), lambda p: p.u(0, 1) * p.v(0, 2))
    assert_values_close(one, r1[2])
    eng.close()


def test_recognition_of_the_interior_penalty_example_without_a_gpu():
    """docs/src/src_jl/example_hello_world_dg.jl: what the three integrals of `a` and the Nitsche integral of `l` are recognised as"""
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2))
    n = GT.unit_normal(mesh, 2)
    Om, Gd, Lam = GT.interior(mesh), GT.boundary(mesh), GT.skeleton(mesh)
    h_L, h_G = GT.face_diameter_field(Lam), GT.face_diameter_field(Gd)
    g = GT.AnalyticalField(lambda x: x[0] + x[1] + x[2], Om)
    mean = lambda fn, u, x: 0.5 * (fn(u[1], x) + fn(u[2], x))
    jump = lambda u, n_, x: u[2](x) * n_[2](x) + u[1](x) * n_[1](x)
    gamma = GT.uniform_quantity(0.2)
    V = GT.lagrange_space(Om, 1, continuous=False)
    assert V.num_free_dofs() == 64 and V.num_dirichlet_dofs() == 0
    dL, dG = GT.measure(Lam, 2), GT.measure(Gd, 2)
    grad, dot = GT.grad, GT.dot
    u, v = GT._form_arguments(V, 2), GT._form_arguments(V, 1)
    ip = GT.integrate(lambda x: dot((gamma / h_L(x)) * jump(v, n, x), jump(u, n, x)) - dot(jump(v, n, x), mean(grad, u, x))
                      - dot(mean(grad, v, x), jump(u, n, x)), dL).contributions[0][0]
    bs = GT._block_problem(V, dL)
    blocks = sorted(GT.recognise_blocks(ip, bs, "skeleton"))
    assert [b[:3] for b in blocks] == [(pu, pv, E.BLOCK_IP) for pu in range(2) for pv in range(2)]
    assert all(b[4] == (0.2, -0.5, -0.5) for b in blocks)
    nit = GT.integrate(lambda x: (gamma / h_G(x)) * v(x) * u(x) - dot(v(x) * n(x), grad(u, x)) - dot(n(x), grad(v, x)) * u(x), dG).contributions[0][0]
    assert GT._is_blocks_case(V, dG, nit)
    bb = GT._block_problem(V, dG)
    assert GT.recognise_blocks(nit, bb, "boundary") == [(0, 0, E.BLOCK_IP, 1.0, (0.2, -1.0, -1.0))]
    rhs = GT.integrate(lambda x: (gamma / h_G(x)) * v(x) * g(x) - dot(n(x), grad(v, x)) * g(x), dG).contributions[0][0]
    vb, gq = GT.recognise_vblocks(rhs, bb, "boundary", V)
    assert vb == [(0, 1.0, (0.0, 0.2, -1.0))] and gq.shape == (bb.face_nodes.shape[0], bb.w.size)
    assert np.allclose(gq, MF.face_point_coordinates(bb).sum(axis=2))
    # a continuous space with a plain Robin term keeps the trace-space path; with normals it goes to the block kernels
    W = GT.lagrange_space(Om, 1)
    w2, w1 = GT._form_arguments(W, 2), GT._form_arguments(W, 1)
    robin = GT.integrate(lambda x: w2(x) * w1(x), dG).contributions[0][0]
    assert not GT._is_blocks_case(W, dG, robin) and GT._is_blocks_case(W, dG, nit)
    with pytest.raises(GT.UnsupportedFormError):          # a gradient-gradient product on faces is not a recognised face term
        GT.recognise_blocks(GT.integrate(lambda x: dot(grad(u, x), grad(v, x)), dG).contributions[0][0], bb, "boundary")


def test_reference_integration_tests_with_unit_normals():
    """test/integration_tests.jl:86-106, 158-160 on the ORACLE's face machinery (domain (0,2)^2, 8 x 8 cells, degree 2):
    `∫(x->norm(n(x)),dΓ) ≈ 4` over the sides "1-face-2", "1-face-4" and `∫(x->norm(n[1](x)+n[2](x)),dΛ) + 1 ≈ 1`"""
    mesh = H.cartesian_mesh((0, 2, 0, 2), (8, 8))
    V = H.lagrange_space(mesh, 1, None)
    X = mesh.node_coordinates

    def integrate(bp, f):
        total = 0.0
        for i in range(bp.face_nodes.shape[0]):
            for q in range(bp.w.size):
                Jf = O.point_geometry(X, bp.face_nodes[i:i + 1], bp.dM[q])
                dV = float(O.change_of_measure(Jf)[0] * bp.w[q])                      # weight(::MeshFace), accessors.jl:1000-1007
                normals = []
                for a in range(bp.n_sides):
                    cell, var = int(bp.side_cells[i, a]), int(bp.face_var[i, a])
                    Jc = O.point_geometry(X, mesh.cell_nodes[cell - 1][None, :], bp.dM_cell[var][q])[0]
                    normals.append(O.map_unit_normal(Jc, bp.ref_normals[var]))
                total += f(normals) * dV
        return total

    bb = MF.boundary_problem([V], [2, 4], 2)
    assert integrate(bb, lambda n: np.linalg.norm(n[0])) == pytest.approx(4.0, rel=1e-13)
    bs = MF.skeleton_problem([V], 2, gradients=True)
    assert integrate(bs, lambda n: np.linalg.norm(n[0] + n[1])) + 1.0 == pytest.approx(1.0, abs=1e-13)
    assert integrate(bs, lambda n: 1.0) == pytest.approx(2 * 7 * 2.0, rel=1e-13)       # total length of the interior faces


# ---- Nitsche terms without the 1/h scaling (test/issue_224.jl) -------------------------------------------------------------------
def _issue_224_forms(mesh, V):
    Om, Gam = GT.interior(mesh), GT.boundary(mesh)
    dO, dG = GT.measure(Om, 2), GT.measure(Gam, 2)
    n = GT.unit_normal(mesh, mesh.D - 1)
    g = GT.AnalyticalField(lambda x: sum(x[k] for k in range(mesh.D)), Om)
    grad, dot = GT.grad, GT.dot
    a = lambda u, v: (GT.integrate(lambda x: v(x) * u(x) - dot(v(x) * n(x), grad(u, x)) - dot(n(x), grad(v, x)) * u(x), dG)
                      + GT.integrate(lambda x: dot(grad(u, x), grad(v, x)), dO))
    l = lambda v: (GT.integrate(lambda x: v(x) * g(x) - dot(n(x), grad(v, x)) * g(x), dG) + GT.integrate(lambda x: v(x) * 0, dO))
    return a, l, g, dO, dG


def test_recognition_of_nitsche_terms_without_the_penalty_scaling():
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2), simplexify=True)
    V = GT.lagrange_space(GT.interior(mesh), 1)
    a, l, g, dO, dG = _issue_224_forms(mesh, V)
    term = a(GT._form_arguments(V, 2), GT._form_arguments(V, 1)).contributions[0][0]
    assert GT._is_blocks_case(V, dG, term)
    bb = GT._block_problem(V, dG)
    assert GT.recognise_blocks(term, bb, "boundary") == [(0, 0, E.BLOCK_IP_NOH, 1.0, (1.0, -1.0, -1.0))]
    lt = l(GT._form_arguments(V, 1)).contributions[0][0]
    vb, gq = GT.recognise_vblocks(lt, bb, "boundary", V)
    assert vb == [(0, 1.0, (1.0, 0.0, -1.0))] and gq.shape[0] == bb.face_nodes.shape[0]


@pytest.mark.gpu
@pytest.mark.parametrize("cells,simplexify", [((3, 3), False), ((2, 2, 2), True)])
def test_gpu_nitsche_without_scaling_parity(cells, simplexify):
    D = len(cells)
    mesh = H.cartesian_mesh(tuple([0, 1] * D), cells, simplexify=simplexify)
    _warp(mesh)
    V = H.lagrange_space(mesh, 1, None)
    bp = MF.boundary_problem([V], None, 2)
    eng = _engine(bp)
    eng.set_skeleton_cells(bp.cell_nodes, bp.side_cells, bp.dM_cell, bp.ref_normals)
    mat = lambda p: p.v(0) * p.u(0) - O.frobenius(p.v(0) * p.n(1), p.grad_u(0)) - O.frobenius(p.n(1), p.grad_v(0)) * p.u(0)
    ref = O.assemble_matrix_multifield(D, mesh.node_coordinates, bp.face_nodes, dict(w=bp.w, dM=bp.dM), _boundary_oracle_sides(bp),
                                       _oracle_fields(bp, [V], True), mat, skeleton_geometry=(bp.cell_nodes, bp.dM_cell, bp.ref_normals))
    eng.matrix_symbolic()
    gcp, grv = eng.matrix_pattern()
    assert np.array_equal(ref[0], gcp) and np.array_equal(ref[1], grv)
    assert_values_close(eng.matrix_numeric_blocks([(0, 0, E.BLOCK_IP_NOH, 1.0, (1.0, -1.0, -1.0))]), ref[2])
    eng.close()


@pytest.mark.gpu
def test_gpu_reference_issue_224_nitsche_on_tetrahedra():
    """test/issue_224.jl transcribed (on the simplexified 2 x 2 x 2 mesh instead of its single hand-made tetrahedron): weak Dirichlet
    conditions through v u - v n⋅∇u - n⋅∇v u on Γ, continuous P1, exact solution sum(x): `@test sqrt(sum(int)) < 1.0e-10`"""
    import scipy.sparse.linalg as spla
    mesh = GT.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2), simplexify=True)
    V = GT.lagrange_space(GT.interior(mesh), 1)
    a, l, g, dO, dG = _issue_224_forms(mesh, V)
    A = GT.assemble_matrix(a, float, V, V)
    b = GT.assemble_vector(l, float, V)
    x = spla.spsolve(A.to_scipy().tocsc(), b)
    assert np.abs(x - V.data.free_dof_nodes.sum(axis=1)).max() < 1e-10
    uh = GT.solution_field(V, x)
    assert np.sqrt(GT.integrate(lambda y: GT.abs2(uh(y) - g(y)), dO).sum()) < 1.0e-10


@pytest.mark.parametrize("simplexify", [True, False])
def test_oracle_reproduces_the_reference_issue_224_known_answer(simplexify):
    """test/issue_224.jl (tetrahedra) and test/issue_230.jl (hexahedra, there from Gmsh) on the CPU ORACLE (2 x 2 x 2 Cartesian mesh,
    simplexified or not): Nitsche terms without penalty scaling on Γ + Laplace operator, data right-hand side; the discrete solution
    is sum(x): `@test sqrt(sum(int)) < 1.0e-10` / `@assert sqrt(sum(int)) < 1.0e-9`"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2), simplexify=simplexify)
    V = H.lagrange_space(mesh, 1, None)
    bb = MF.boundary_problem([V], None, 2)
    geo = (bb.cell_nodes, bb.dM_cell, bb.ref_normals)
    sides, flds = _boundary_oracle_sides(bb), _oracle_fields(bb, [V], True)
    nit = lambda p: p.v(0) * p.u(0) - O.frobenius(p.v(0) * p.n(1), p.grad_u(0)) - O.frobenius(p.n(1), p.grad_v(0)) * p.u(0)
    coo = [O.assemble_matrix_multifield(3, mesh.node_coordinates, bb.face_nodes, dict(w=bb.w, dM=bb.dM), sides, flds, nit,
                                        skeleton_geometry=geo, return_coo=True), _volume_coo(O.LAPLACE, mesh, V, 2)]
    cp, rv, nz = O.assemble_matrix_sum(coo, V.n_free, V.n_free)
    gq = MF.face_point_coordinates(bb).sum(axis=2)
    rhs = lambda p: p.v(0) * p.g - O.frobenius(p.n(1), p.grad_v(0)) * p.g
    b = O.assemble_vector_multifield(3, mesh.node_coordinates, bb.face_nodes, dict(w=bb.w, dM=bb.dM), sides, flds, rhs,
                                     skeleton_geometry=geo, point_data=gq)
    A = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
    x = spla.spsolve(A, b)
    assert np.abs(x - V.free_dof_nodes.sum(axis=1)).max() < 1e-10
