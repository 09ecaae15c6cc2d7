"""Shared problem builders for the tests: inputs come from the package's hostprep (what the
product ships), the expected results from the oracle (test infrastructure)."""
import numpy as np

import gt_oracle as O
import gtk_b200

H = gtk_b200.hostprep


def problem(cells, order=1, bc="boundary", n_comp=1, degree=None, simplexify=False, warp=0.0, seed=0, domain=None):
    D = len(cells)
    domain = domain or tuple([0, 1] * D)
    mesh = H.cartesian_mesh(domain, cells, simplexify=simplexify)
    if warp:
        # interior nodes displaced by warp*h*U(-1,1): non-affine cells (SURVEY.md §8d robustness variant)
        rng = np.random.default_rng(seed)
        h = np.array([(domain[2 * d + 1] - domain[2 * d]) / cells[d] for d in range(D)])
        inner = ~H.boundary_node_mask(mesh)
        mesh.node_coordinates[inner] += warp * h * rng.uniform(-1, 1, size=(int(inner.sum()), D))
    V = H.lagrange_space(mesh, order, bc, n_comp)
    tab = H.measure_tabulation(V, 2 * order if degree is None else degree)
    return mesh, V, tab


def tab_dict(tab):
    return dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)


def oracle_matrix(form, mesh, V, tab, fd=(O.FREE, O.FREE), **params):
    return O.assemble_matrix(form, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet,
                             tab_dict(tab), n_comp=V.n_comp, free_or_dirichlet=fd, **params)


def oracle_vector(form, mesh, V, tab, fd=O.FREE, **params):
    return O.assemble_vector(form, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet,
                             tab_dict(tab), n_comp=V.n_comp, free_or_dirichlet=fd, **params)


def make_engine(mesh, V, tab, device=0):
    eng = gtk_b200.engine.Engine(device)
    eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
    eng.set_space(V.cell_dofs, V.n_free, V.n_dirichlet, V.n_comp)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    return eng


def assert_values_close(got, ref, tol=1e-12):
    """max|Δ| <= 1e-12 * max|ref| (norm-relative: structurally present entries can be
    analytically zero, e.g. 3D Q1 Laplacian edge neighbours)."""
    got = np.asarray(got); ref = np.asarray(ref)
    assert got.shape == ref.shape
    scale = np.abs(ref).max() if ref.size else 1.0
    err = np.abs(got - ref).max() if ref.size else 0.0
    assert err <= tol * max(scale, 1e-300), f"max|Δ|={err:.3e} > {tol:g}*max|ref|={scale:.3e}"
