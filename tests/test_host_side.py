"""CPU tests of the host side: input preparation vs the oracle's literal restatement, form
recognition, the C-ABI library (loads, exports every declared symbol, fails loudly without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200

H, GT, E = gtk_b200.hostprep, gtk_b200.GT, gtk_b200.engine


@pytest.mark.parametrize("cells,domain", [((4, 3), (0, 1, 0, 1)), ((3, 2, 4), (0, 1, 0, 2, 0, 1)), ((2, 2, 2), (0, 1, 0, 1, 0, 1)), ((1, 1), (0, 1, 0, 1))])
@pytest.mark.parametrize("bc", [None, "boundary", [1]])
@pytest.mark.parametrize("n_comp", [1, 2])
def test_hostprep_matches_literal_numbering(cells, domain, bc, n_comp):
    o = O.q1_space(domain, cells, bc, n_comp=n_comp)
    mesh = H.cartesian_mesh(domain, cells)
    V = H.lagrange_space(mesh, 1, bc, n_comp)
    assert np.array_equal(mesh.cell_nodes, o["cell_nodes"])
    assert np.array_equal(mesh.node_coordinates, o["coords"])
    assert np.array_equal(V.cell_dofs, o["cell_dofs"])
    assert (V.n_free, V.n_dirichlet) == (o["n_free"], o["n_dirichlet"])


def test_simplexified_mesh_matches_literal():
    for cells, dom in (((3, 2), (0, 1, 0, 1)), ((2, 3, 2), (0, 1, 0, 1, 0, 1))):
        coords, cn = O.cartesian_chain(dom, cells, simplexify=True)
        mesh = H.cartesian_mesh(dom, cells, simplexify=True)
        assert np.array_equal(mesh.cell_nodes, np.array(cn, dtype=np.int32))
        assert np.array_equal(mesh.node_coordinates, coords)


def test_closed_form_q1_dofs_full_dirichlet():
    # free id of interior node (i,j,k) = 1+(i-1)+(n1-1)(j-1)+(n1-1)(n2-1)(k-1)
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (5, 4, 3))
    V = H.lagrange_space(mesh, 1, "boundary")
    cell = 1 + 1 * 5 + 1 * 20          # cell (1,1,1) 0-based: all 8 nodes... node (1..2,1..2,1..2)
    d = V.cell_dofs[cell]
    assert d[0] == 1 + 0 + 4 * 0 + 12 * 0
    assert d[1] == 2 and d[2] == 1 + 4 and d[4] == 1 + 12
    assert V.n_free == 4 * 3 * 2


def test_tabulation_and_quadrature_match_oracle():
    for D in (2, 3):
        q = H.quadrature(D, False, 2)
        x, w = O.tensor_gauss(D, 2)
        assert np.array_equal(q.coordinates, x) and np.array_equal(q.weights, w)
        N, dN = H.tabulate(D, 1, "Q", q.coordinates)
        N2, dN2 = O.tabulate(D, 1, "Q", x)
        assert np.allclose(N, N2, rtol=0, atol=1e-15) and np.allclose(dN, dN2, rtol=0, atol=1e-15)


def test_partition_slab_inputs_cover_the_global_mesh():
    full = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (4, 4, 6))
    parts = [H.cartesian_mesh((0, 1, 0, 1, 0, 1), (4, 4, 6), z_cell_range=r) for r in ((0, 2), (2, 6))]
    assert np.array_equal(np.vstack([p.cell_nodes for p in parts]), full.cell_nodes)


# ---- form recognition -----------------------------------------------------------------------
def _space():
    mesh = GT.cartesian_mesh((0, 1, 0, 1), (4, 4))
    om = GT.interior(mesh)
    V = GT.lagrange_space(om, 1, dirichlet_boundary=GT.boundary(mesh))
    return mesh, om, V, GT.measure(om, 2)


def test_form_recognition():
    mesh, om, V, dom = _space()
    u, v = GT.FormArgument(V, 2), GT.FormArgument(V, 1)
    t = GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dom).contributions[0][0]
    assert GT.recognise_bilinear(t) == (E.FORM_LAPLACE, dict(alpha=1.0))
    t = GT.integrate(lambda x: 3.0 * (u(x) * v(x)), dom).contributions[0][0]
    assert GT.recognise_bilinear(t) == (E.FORM_MASS, dict(alpha=3.0))
    t = GT.integrate(lambda x: GT.isotropic_elasticity(2.0, 1.0)(u, v, x), dom).contributions[0][0]
    assert GT.recognise_bilinear(t)[0] == E.FORM_ELASTICITY_ISO
    f = GT.analytical_field(lambda x: x[0] + x[1])
    form, params = GT.recognise_linear(GT.integrate(lambda x: f(x) * v(x), dom).contributions[0][0], V, dom)
    assert form == E.FORM_SOURCE_QP and params["f_qp"].shape == (16, 4, 1)
    form, params = GT.recognise_linear(GT.integrate(lambda x: 2.0 * v(x), dom).contributions[0][0], V, dom)
    assert form == E.FORM_SOURCE_CONST and params["alpha"] == 2.0
    # variable coefficient κ(x) ∇u·∇v / κ(x) u v: κ sampled by the host at the quadrature points
    kappa = GT.analytical_field(lambda x: 1.0 + x[0] * x[1])
    t = GT.integrate(lambda x: 2.0 * (kappa(x) * GT.dot(GT.grad(u, x), GT.grad(v, x))), dom).contributions[0][0]
    form, params = GT.recognise_bilinear(t, V, dom)
    assert form == E.FORM_LAPLACE and params["alpha"] == 2.0 and params["coef_qp"].shape == (16, 4)
    xq = GT.quadrature_point_coordinates(V, dom)
    assert np.allclose(params["coef_qp"], 1.0 + xq[..., 0] * xq[..., 1])
    t = GT.integrate(lambda x: kappa(x) * (u(x) * v(x)), dom).contributions[0][0]
    assert GT.recognise_bilinear(t, V, dom)[0] == E.FORM_MASS


def test_unsupported_forms_raise_instead_of_falling_back():
    mesh, om, V, dom = _space()
    u, v = GT.FormArgument(V, 2), GT.FormArgument(V, 1)
    with pytest.raises(GT.UnsupportedFormError):       # convection-like, non-symmetric
        GT.recognise_bilinear(GT.integrate(lambda x: u(x) * GT.dot(GT.grad(v, x), GT.grad(v, x)), dom).contributions[0][0])
    with pytest.raises(GT.UnsupportedFormError):
        GT.recognise_bilinear(GT.integrate(lambda x: u(x) + v(x), dom).contributions[0][0])
    with pytest.raises(GT.UnsupportedFormError):
        GT.recognise_linear(GT.integrate(lambda x: GT.dot(GT.grad(v, x), GT.grad(v, x)), dom).contributions[0][0], V, dom)
    # boundary measures exist (Neumann / Robin terms), but forms with physical gradients on Γ are refused
    dG = GT.measure(GT.boundary(mesh), 2)
    with pytest.raises(GT.UnsupportedFormError):
        GT.assemble_matrix(lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dG), np.float64, V, V)


# ---- the C-ABI library ----------------------------------------------------------------------
def test_library_exports_every_declared_symbol(built_library):
    header = open(os.path.join(os.path.dirname(gtk_b200.PACKAGE_DIR), "include", "gtk_assembly.h")).read()
    declared = set(re.findall(r"\b(gtk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(E.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(built_library, name), name
    assert built_library.gtk_version() >= 100


def test_ctypes_prototypes_match_the_header(built_library):
    """Every entry point has a ctypes prototype whose arity equals the C declaration's (a wrong arity silently
    passes garbage in the upper halves of 64-bit arguments)."""
    header = open(os.path.join(os.path.dirname(gtk_b200.PACKAGE_DIR), "include", "gtk_assembly.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(gtk_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", header)
    assert len(protos) == len(E.ABI_SYMBOLS)
    for name, args in protos:
        n = 0 if args.strip() in ("", "void") else len(args.split(","))
        fn = getattr(built_library, name)
        assert fn.argtypes is not None, f"{name} has no ctypes prototype"
        assert len(fn.argtypes) == n, f"{name}: header has {n} parameters, ctypes prototype {len(fn.argtypes)}"


def test_engine_fails_loudly_without_gpu(built_library):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.GtkError):
        E.Engine(0)
    with pytest.raises(E.GtkError):
        mesh, om, V, dom = _space()
        GT.assemble_matrix(lambda u, v: GT.integrate(lambda x: GT.dot(GT.grad(u, x), GT.grad(v, x)), dom), float, V, V)


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    pat = re.compile(r"^\s*(import\s+(gt_oracle|oracle)|from\s+(gt_oracle|oracle)[\s.]|#include\s+\".*oracle|.*dlopen.*oracle|.*CDLL.*oracle)", re.M)
    for root, _, files in os.walk(gtk_b200.PACKAGE_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".jl")):
                src = open(os.path.join(root, f)).read()
                assert not pat.search(src), f


def test_boundary_faces_follow_cartesian_mesh_order_and_neumann_total():
    """cartesian_mesh.jl:117-168: boundary faces cell-major / local face ascending, group = local face id; a constant
    flux over Γ integrates to |Γ| (oracle, partition of unity on the faces)."""
    import gt_oracle as O
    H = gtk_b200.hostprep
    mesh = H.cartesian_mesh((0, 2, 0, 1, 0, 3), (4, 3, 2))
    fn, fc, lf = H.boundary_faces(mesh)
    assert fn.shape == (2 * (4 * 3 + 4 * 2 + 3 * 2), 4)
    assert (np.diff(fc) >= 0).all()                                   # cell-major
    same = np.diff(fc) == 0
    assert (np.diff(lf)[same] > 0).all()                              # local faces ascending inside a cell
    # the first cell (corner) owns local faces 1 (z=0), 3 (y=0), 5 (x=0) with the reference's local node order
    assert lf[:3].tolist() == [1, 3, 5]
    assert fn[0].tolist() == mesh.cell_nodes[0][[0, 1, 2, 3]].tolist()
    assert fn[1].tolist() == mesh.cell_nodes[0][[0, 1, 4, 5]].tolist()
    assert fn[2].tolist() == mesh.cell_nodes[0][[0, 2, 4, 6]].tolist()
    for sides, area in (([1], 2.0), ([6], 3.0), ([3, 4], 12.0), (None, 2 * (2 + 3 + 6))):
        for order in (1, 2):
            V = H.lagrange_space(mesh, order, None)
            fp = H.face_problem(V, sides, 2 * order)
            tab = dict(w=fp.tab.w, N=fp.tab.N, dN=fp.tab.dN, M=fp.tab.M, dM=fp.tab.dM)
            b = O.assemble_vector(O.SOURCE_CONST, mesh.node_coordinates, fp.face_nodes, fp.face_dofs, V.n_free, V.n_dirichlet, tab, f_const=[1.0])
            assert abs(b.sum() - area) < 1e-12 * area
    with pytest.raises(ValueError):
        H.boundary_faces(H.cartesian_mesh((0, 1, 0, 1), (3, 1)))     # cartesian_mesh.jl:98-100
    # the dofs of a face sit at the face's own lattice points (cubes: Q1 map of the face; simplices: barycentric), and a
    # simplexified mesh has the simplex sub-faces of every boundary cube face, cube-face order preserved
    for simp in (False, True):
        for order in (1, 2, 3):
            m = H.cartesian_mesh((0, 2, 0, 1, 0, 3), (4, 3, 2), simplexify=simp)
            V = H.lagrange_space(m, order, [1])
            fp = H.face_problem(V, [2, 3, 6], 2 * order)
            assert fp.face_nodes.shape[0] == (2 if simp else 1) * (4 * 3 + 4 * 2 + 3 * 2)
            X = m.node_coordinates[fp.face_nodes.astype(np.int64) - 1]
            lat = H.monomial_exponents(2, order, "P" if simp else "Q").astype(np.float64) / order
            if simp:
                xl = X[:, None, 0, :] + np.einsum("lm,fmd->fld", lat, X[:, 1:, :] - X[:, :1, :])
            else:
                xl = X[:, None, 0, :] + np.einsum("lm,fmd->fld", lat, np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]], axis=1))
            d = fp.face_dofs
            ii = np.abs(d) - 1
            got = np.where((d > 0)[..., None], V.free_dof_nodes[np.minimum(ii, V.n_free - 1)],
                           V.dirichlet_dof_nodes[np.minimum(ii, V.n_dirichlet - 1)])
            assert np.allclose(got, xl, atol=1e-13)


@pytest.mark.parametrize("cells,domain", [((3, 2), (0, 1, 0, 1)), ((4, 3), (0, 2, 0, 1)), ((2, 3, 2), (0, 1, 0, 1, 0, 1)), ((3, 3, 3), (0, 1, 0, 1, 0, 1)),
                                          ((1, 1, 1), (0, 1, 0, 1, 0, 1)), ((4, 2, 3), (0, 1, 0, 2, 0, 1))])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_high_order_numbering_matches_literal_face_complex(cells, domain, order):
    """hostprep (vectorised, first occurrence by np.unique) against the oracle's loop-for-loop restatement of complexify +
    generate_dof_ids (topology.jl:1034-1125, 1468-1704; space.jl:299-535) — SURVEY.md A.3-A.5.  Also: the numbering is
    conforming (a dof shared by two cells sits at one physical point) and Dirichlet dofs are exactly those on Γ."""
    import importlib
    R = importlib.import_module("galerkintoolkit_jl_b200.refnumbering")
    mesh = H.cartesian_mesh(domain, cells)
    D = len(cells)
    for bc in (None, "boundary", [1, 4]):
        for n_comp in (1, 2):
            lit = O.lagrange_space_literal(domain, cells, order, bc, n_comp=n_comp)
            cd, nf, nd, xf, xd = R.scalar_or_vector_dofs(mesh, order, n_comp, bc)
            assert np.array_equal(cd, lit["cell_dofs"]) and (nf, nd) == (lit["n_free"], lit["n_dirichlet"])
            if order == 1:      # consistent with the separate Q1 restatement
                assert np.array_equal(cd, O.q1_space(domain, cells, bc, n_comp=n_comp)["cell_dofs"])
            if order <= 3:      # simplexified mesh (cartesian_mesh.jl:265-461; simplexify(::UnitNCube) domain.jl:270-336)
                smesh = H.cartesian_mesh(domain, cells, simplexify=True)
                lit = O.lagrange_space_literal(domain, cells, order, bc, n_comp=n_comp, simplexify=True)
                cd, nf, nd, xf, xd = R.scalar_or_vector_dofs(smesh, order, n_comp, bc)
                assert np.array_equal(cd, lit["cell_dofs"]) and (nf, nd) == (lit["n_free"], lit["n_dirichlet"])
                if order == 1:
                    assert np.array_equal(cd, O.q1_space(domain, cells, bc, simplexify=True, n_comp=n_comp)["cell_dofs"])
    if order >= 2:      # simplexified: dof positions hostprep reports are the barycentric lattice points of the cells
        smesh = H.cartesian_mesh(domain, cells, simplexify=True)
        Vs = H.lagrange_space(smesh, order, [1, 4], 1)
        lat_s = np.array(O._simplex_lattice(D, order), dtype=np.float64) / order
        Xs = smesh.node_coordinates[smesh.cell_nodes.astype(np.int64) - 1]              # [nc, D+1, D]
        xs = Xs[:, None, 0, :] + np.einsum("lm,cmd->cld", lat_s, Xs[:, 1:, :] - Xs[:, :1, :])
        ds = Vs.cell_dofs
        ii = np.abs(ds) - 1
        got = np.where((ds > 0)[..., None], Vs.free_dof_nodes[np.minimum(ii, Vs.n_free - 1)],
                       Vs.dirichlet_dof_nodes[np.minimum(ii, max(Vs.n_dirichlet, 1) - 1)])
        assert np.allclose(got, xs, atol=1e-13)
        on = (np.abs(xs[..., D - 1] - domain[2 * (D - 1)]) < 1e-13) | (np.abs(xs[..., D - 2] - domain[2 * (D - 2) + 1]) < 1e-13)
        assert np.array_equal(ds < 0, on)
    V = H.lagrange_space(mesh, order, [1, 4])
    # conformity + geometry of the Dirichlet set through the dof coordinates hostprep reports
    lat = np.array(O._lattice(D, order), dtype=np.float64) / order
    X = mesh.node_coordinates[mesh.cell_nodes.astype(np.int64) - 1]                 # [nc, 2^D, D]
    lo, hi = X[:, 0, :], X[:, -1, :]
    xl = lo[:, None, :] + lat[None, :, :] * (hi - lo)[:, None, :]                    # dof positions per cell
    d = V.cell_dofs
    idx = np.abs(d) - 1
    got = np.where((d > 0)[..., None], V.free_dof_nodes[np.minimum(idx, V.n_free - 1)],
                   V.dirichlet_dof_nodes[np.minimum(idx, V.n_dirichlet - 1)])
    assert np.allclose(got, xl, atol=1e-13)
    on = (np.abs(xl[..., D - 1] - domain[2 * (D - 1)]) < 1e-13) | (np.abs(xl[..., D - 2] - domain[2 * (D - 2) + 1]) < 1e-13)   # sides 1 and 4
    assert np.array_equal(d < 0, on)


def test_every_entry_point_rejects_a_null_context(built_library):
    """No GPU needed: a NULL ctx is an argument error (or a defined sentinel), never a crash."""
    lib = built_library
    z = None      # NULL for every pointer parameter
    calls = {
        "gtk_destroy": (z,), "gtk_set_stream": (z, z), "gtk_set_mesh": (z, 3, 0, z, 0, 8, z), "gtk_set_manifold_dim": (z, 2),
        "gtk_set_active_cells": (z, 0, 0), "gtk_update_coordinates": (z, z), "gtk_set_space": (z, 8, 1, z, 0, 0),
        "gtk_set_tabulation": (z, 8, z, z, z, z, z), "gtk_matrix_symbolic": (z, 1, 1, z), "gtk_matrix_pattern": (z, z, z),
        "gtk_matrix_numeric": (z, 1, z, z), "gtk_matrix_numeric_device": (z, 1, z), "gtk_vector_symbolic": (z, 1),
        "gtk_set_vector": (z, z), "gtk_vector_assemble": (z, 101, z, z), "gtk_vector_assemble_device": (z, 101, z),
        "gtk_assemble_matrix_and_vector": (z, 1, z, 101, z, z, z), "gtk_assemble_matrix_and_vector_device": (z, 1, z, 101, z),
        "gtk_select_matrix": (z, 0), "gtk_matvec_add_device": (z, 1.0, z, 1.0), "gtk_matvec_add": (z, 1.0, z, 1.0, z),
        "gtk_device_pointer": (z, 0, z, z), "gtk_copy_nzval": (z, z), "gtk_copy_vector": (z, z), "gtk_set_profiling": (z, 1),
        "gtk_profile_get": (z, 0, z, z), "gtk_comm_init": (z, 0, 1, z), "gtk_comm_set_exchange": (z, 1, 0, z, 0, z, 0, z, 0, z),
        "gtk_comm_sum_ghost_rows": (z,), "gtk_assemble_and_sum_ghost_rows_device": (z, 1, z, 101, z),
        "gtk_comm_p2p_export": (z, 0, z), "gtk_comm_p2p_import": (z, 0, z),
        "gtk_field_set_values": (z, z, z), "gtk_field_set_values_device": (z, z, z), "gtk_field_get_values": (z, z, z),
        "gtk_field_axpy_free": (z, 1.0, z), "gtk_space_dof_coordinates": (z, z, z, z), "gtk_scalar_assemble": (z, 201, z, z),
        "gtk_comm_build_exchange": (z, 0, z), "gtk_comm_connect_peer_memory": (z,),
        "gtk_set_cartesian_q1_problem": (z, z, z, 0, 2, 0, z, z), "gtk_copy_device_array": (z, 0, z, 0), "gtk_matrix_pattern_i64": (z, z, z),
        "gtk_set_parts": (z, 4, z, z, z, 1, z, 1, 1, z), "gtk_matrix_numeric_blocks": (z, 0, z, z),
        "gtk_matrix_numeric_blocks_device": (z, 0, z), "gtk_vector_assemble_blocks": (z, 0, z, 0, z),
        "gtk_vector_assemble_blocks_device": (z, 0, z, 0),
        "gtk_matrix_sum_symbolic": (z, 1, z, z), "gtk_matrix_sum_numeric": (z, 1, z, z), "gtk_matrix_sum_numeric_device": (z, 1, z),
        "gtk_set_skeleton_cells": (z, 1, 4, z, z, z, z), "gtk_matrix_colptr_at": (z, 0, z, z),
        "gtk_vector_assemble_blocks_data": (z, 0, z, z, 0, z), "gtk_vector_assemble_blocks_data_device": (z, 0, z, z, 0),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc in (E.GTK_OK, E.GTK_ERR_INVALID) if name == "gtk_destroy" else rc == E.GTK_ERR_INVALID, (name, rc)
    assert lib.gtk_info(z, 0) < 0 and lib.gtk_comm_ghost_info(z, 0) < 0 and lib.gtk_profile_count(z) <= 0
    lib.gtk_last_error.restype = ctypes.c_char_p
    assert b"null" in lib.gtk_last_error(z).lower()
    covered = set(calls) | {"gtk_info", "gtk_comm_ghost_info", "gtk_profile_count", "gtk_last_error", "gtk_version", "gtk_create", "gtk_comm_unique_id"}
    assert covered == set(E.ABI_SYMBOLS), set(E.ABI_SYMBOLS) ^ covered
