"""ORACLE — test infrastructure only.  Pinned by the reference's own known answers (see below).

CPU restatement of GalerkinToolkit.jl v0.6.3's assembly hot path, following the
reference line by line.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this file.
The product (``galerkintoolkit.jl_b200``) never does.

Pinning status: the reference is pure Julia and cannot run in this container
(no ``julia``, no depot, no network), and its test-suite holds **no golden
matrices** for assembly.  What it does hold on this path are
numeric known answers, and this oracle reproduces them:
  * the discrete p-Laplacian L2 norm 0.09133166701839236
    (test/problems_ext_tests.jl:148-172, tolerance 1e-10) —
    tests/test_oracle_invariants.py::test_oracle_reproduces_the_reference_plaplacian_golden;
  * the interior-penalty example's ``@assert el2 < 1.0e-9``
    (docs/src/src_jl/example_hello_world_dg.jl: discontinuous space, skeleton +
    Nitsche terms, unit normals, face diameters) —
    tests/test_multifield.py::test_oracle_reproduces_the_reference_interior_penalty_example;
  * test/issue_224.jl's ``@test sqrt(sum(int)) < 1.0e-10`` (Nitsche terms on tetrahedra) —
    tests/test_multifield.py::test_oracle_reproduces_the_reference_issue_224_known_answer;
plus the reference's invariants checked in ``tests/test_oracle_invariants.py``:
sum(M)=|Ω|, sum(b)=∫f (test/problems_tests.jl:53-57), tabulator(nodes)=I
(test/space_tests.jl:218-220), quadrature weights sum (test/integration_tests.jl:28-34),
dof counts (test/assembly_tests.jl:72-73), manufactured Poisson solution
(test/problems_tests.jl:96-105), ∫_Λ jump(v) ≈ 0 (test/assembly_tests.jl:420-427).
NOT pinned by any reference-held number: dof numbering (norms are numbering-invariant),
order >= 2, simplices, entry-level colptr/rowval — those are re-derived from the
cited source lines and cross-checked between two independent restatements.

Third-party algorithms restated (not vendored in /root/reference; compat pins of
Project.toml:39-63): StaticArrays 1.x closed-form ``det``/``\\`` for 2x2 and 3x3
SMatrix, SparseArrays ``sparse(I,J,V,m,n)`` (column-major, rows sorted, duplicates
summed in input order, explicit zeros kept), PartitionedArrays 0.5.4
``sparse_matrix``/``dense_vector`` (skip indices < 1), FastGaussQuadrature
``gausslegendre``.

All citations are relative to /root/reference/src.  Arithmetic is vectorised
over *cells only*; per cell the operation order is exactly the scalar order of
the reference (numpy elementwise ops are IEEE-754 and are not FMA-contracted).
"""
from __future__ import annotations

import itertools
import numpy as np

FREE, DIRICHLET = 1, 2

# ---------------------------------------------------------------------------
# Literal mesh + numbering restatements (small sizes; python loops)
# ---------------------------------------------------------------------------
SIMPLEX_NODES = {2: [[1, 2, 3], [4, 3, 2]],
                 3: [[1, 2, 3, 7], [1, 2, 5, 7], [2, 3, 4, 7], [2, 4, 7, 8], [2, 5, 6, 7], [2, 6, 7, 8]]}  # domain.jl:322-336

# local faces of the reference cube / simplex: domain.jl:224, 252-255, 452-457
CUBE_FACES = {
    2: [[[1], [2], [3], [4]], [[1, 2], [3, 4], [1, 3], [2, 4]]],
    3: [[[i] for i in range(1, 9)],
        [[1, 2], [3, 4], [1, 3], [2, 4], [5, 6], [7, 8], [5, 7], [6, 8], [1, 5], [3, 7], [2, 6], [4, 8]],
        [[1, 2, 3, 4], [5, 6, 7, 8], [1, 2, 5, 6], [3, 4, 7, 8], [1, 3, 5, 7], [2, 4, 6, 8]]],
}


def cartesian_chain(domain, cells_per_dir, simplexify=False):
    """cartesian_mesh.jl:213-263 (hex) / :265-328 (simplices), loop for loop.
    Returns coords [n,D] and cell_nodes (list of lists, 1-based)."""
    D = len(cells_per_dir)
    pmin = [domain[2 * d] for d in range(D)]
    pmax = [domain[2 * d + 1] for d in range(D)]
    h = [(pmax[d] - pmin[d]) / cells_per_dir[d] for d in range(D)]
    npd = [c + 1 for c in cells_per_dir]

    def lin(ci, dims):  # LinearIndices, 1-based, first index fastest
        li, s = 0, 1
        for d in range(len(dims)):
            li += (ci[d] - 1) * s
            s *= dims[d]
        return li + 1

    def cis(dims):      # CartesianIndices iteration order (first fastest), 1-based tuples
        for t in itertools.product(*[range(1, n + 1) for n in reversed(dims)]):
            yield tuple(reversed(t))

    lnode_cis = [tuple(reversed(t)) for t in itertools.product(*[range(0, 2)] * D)]
    cell_nodes = []
    for cell_ci in cis(cells_per_dir):
        cl = [lin(tuple(cell_ci[d] + ln[d] for d in range(D)), npd) for ln in lnode_cis]
        if simplexify:
            for lnodes in SIMPLEX_NODES[D]:
                cell_nodes.append([cl[i - 1] for i in lnodes])
        else:
            cell_nodes.append(cl)
    coords = np.zeros((int(np.prod(npd)), D))
    for node_li, node_ci in enumerate(cis(npd)):
        for d in range(D):
            coords[node_li, d] = pmin[d] + h[d] * (node_ci[d] - 1)
    return coords, cell_nodes


def boundary_0faces(cell_nodes, nnodes, D):
    """cartesian_mesh.jl:107-166 for d=0: nodes touched by exactly one cell,
    enumerated cell-major then local-vertex order."""
    node_to_n = [0] * (nnodes + 1)
    for nodes in cell_nodes:
        for n in nodes:
            node_to_n[n] += 1
    out = []
    for nodes in cell_nodes:
        for n in nodes:              # local 0-faces are the local nodes in order
            if node_to_n[n] <= 1:    # nmax = 2^0
                out.append(n)
    return out


def vertex_ids(cell_nodes, nnodes, preexisting_0faces):
    """topology.jl:1034-1097 (create_vertices!)."""
    node_vertex = [0] * (nnodes + 1)
    for nodes in cell_nodes:
        for n in nodes:
            node_vertex[n] = -1
    v = 0
    for n in preexisting_0faces:
        if node_vertex[n] == -1:
            v += 1
            node_vertex[n] = v
    for n in range(1, nnodes + 1):
        if node_vertex[n] == -1:
            v += 1
            node_vertex[n] = v
    return node_vertex[1:]


def q1_space(domain, cells_per_dir, dirichlet_sides=None, simplexify=False, n_comp=1):
    """lagrange_space(Ω,1;dirichlet_boundary) on cartesian_mesh: space.jl:299-535.
    One own dof per vertex ⇒ dof = vertex id (:348-417); Dirichlet tagging of the
    dofs of boundary (D-1)-faces (:477-511) and stable partition (:512-524, :910-920).
    ``dirichlet_sides``: None (no BC) | "boundary" | list of 1-based box-side ids."""
    D = len(cells_per_dir)
    coords, cell_nodes = cartesian_chain(domain, cells_per_dir, simplexify)
    nn = coords.shape[0]
    if all(c == 1 for c in cells_per_dir) and not simplexify:
        pre = list(cell_nodes[0])
    else:
        pre = boundary_0faces(cell_nodes, nn, D) if not simplexify else _simplex_boundary_0faces(cells_per_dir)
    vert = vertex_ids(cell_nodes, nn, pre)
    scal = [[vert[n - 1] for n in nodes] for nodes in cell_nodes]
    tag = [0] * nn
    if dirichlet_sides is not None:
        npd = [c + 1 for c in cells_per_dir]
        sides = range(1, 2 * D + 1) if dirichlet_sides == "boundary" else dirichlet_sides
        for node in range(nn):
            idx, r = [], node
            for d in range(D):
                idx.append(r % npd[d]); r //= npd[d]
            for s in sides:
                axis = D - 1 - (s - 1) // 2
                if idx[axis] == ((npd[axis] - 1) if (s - 1) % 2 == 1 else 0):
                    tag[vert[node] - 1] = 1
    # vector-valued: dof = (node-1)*n_comp + c  (space.jl:1267-1271)
    ndofs = nn * n_comp
    dof_tag = [tag[(d // n_comp)] for d in range(ndofs)]
    free = [d for d in range(ndofs) if dof_tag[d] == 0]
    diri = [d for d in range(ndofs) if dof_tag[d] != 0]
    perm = [0] * ndofs
    for i, d in enumerate(free):
        perm[d] = i + 1
    for i, d in enumerate(diri):
        perm[d] = -(i + 1)
    cell_dofs = [[perm[(s - 1) * n_comp + c] for s in row for c in range(n_comp)] for row in scal]
    return dict(coords=coords, cell_nodes=np.array(cell_nodes, dtype=np.int32),
                cell_dofs=np.array(cell_dofs, dtype=np.int32), n_free=len(free), n_dirichlet=len(diri))


def _simplex_boundary_0faces(cells_per_dir):
    """structured_simplex_mesh_with_boundary keeps the same 2^D box corners as 0-faces
    (cartesian_mesh.jl:330-461: boundary built from the hex chain), lexicographic."""
    D = len(cells_per_dir)
    npd = [c + 1 for c in cells_per_dir]
    out = []
    for ln in range(2 ** D):
        li, s = 0, 1
        for d in range(D):
            li += ((ln >> d) & 1) * (npd[d] - 1) * s
            s *= npd[d]
        out.append(li + 1)
    return out


# ---------------------------------------------------------------------------
# Reference element / quadrature / tabulation
# ---------------------------------------------------------------------------
def monomial_exponents(D, order, kind):
    """space.jl:1127-1145."""
    out = []
    for t in itertools.product(*[range(order + 1)] * D):
        e = tuple(reversed(t))
        if kind == "P" and sum(e) > order:
            continue
        out.append(e)
    return out


def tabulate(D, order, kind, points):
    """space.jl:960-970 + accessors.jl:486-496: value and gradient tables [p][dof]."""
    exps = monomial_exponents(D, order, kind)
    nodes = [[e[d] / order for d in range(D)] for e in exps]

    def mono(e, x):
        v = 1.0
        for d in range(D):
            v *= x[d] ** e[d]
        return v

    def dmono(e, x, k):
        v = 1.0
        for d in range(D):
            if d == k:
                v *= (e[d] * x[d] ** (e[d] - 1)) if e[d] > 0 else 0.0
            else:
                v *= x[d] ** e[d]
        return v
    A = np.array([[mono(e, x) for e in exps] for x in nodes])
    B = np.linalg.solve(A, np.eye(len(exps)))
    C = np.array([[mono(e, x) for e in exps] for x in points])
    N = C @ B
    dN = np.zeros((len(points), len(exps), D))
    for k in range(D):
        Ck = np.array([[dmono(e, x, k) for e in exps] for x in points])
        dN[:, :, k] = Ck @ B
    return N, dN


def tensor_gauss(D, degree):
    """quadrature.jl:60-106."""
    n = int(np.ceil((degree + 1) / 2))
    x, w = np.polynomial.legendre.leggauss(n)
    x = 0.5 * x + 0.5
    w = 0.5 * w
    pts, wts = [], []
    for t in itertools.product(*[range(n)] * D):
        ci = tuple(reversed(t))
        pts.append([x[i] for i in ci])
        ww = 1.0
        for i in ci:
            ww *= w[i]
        wts.append(ww)
    return np.array(pts), np.array(wts)


def strang_tet(degree):
    """quadrature.jl:500-635, literally: strang_quadrature_<degree>(::UnitSimplex{3}) -> (points [n,3], weights [n]).
    Used by `quadrature(geo, degree)` for tetrahedra of degree 1..5 (quadrature.jl:41-48)."""
    if degree == 1:
        a = 1.0 / 4.0; b = 1.0 / 6.0
        x = [(a, a, a)]; w = [b]
    elif degree == 2:
        a = 0.5854101966249685; b = 0.1381966011250105; c = 1.0 / 24.0
        x = [(b, b, b), (a, b, b), (b, a, b), (b, b, a)]; w = [c, c, c, c]
    elif degree == 3:
        a = 1.0 / 4.0; b = 1.0 / 6.0; c = 1.0 / 2.0; d = -2.0 / 15.0; e = 1.5 / 20.0
        x = [(a, a, a), (b, b, b), (c, b, b), (b, c, b), (b, b, c)]; w = [d, e, e, e, e]
    elif degree == 4:
        a = 0.3994035761667992; b = 0.1005964238332008
        c = (343.0 / 7500.0) / 6.0; d = (56.0 / 375.0) / 6.0
        e = 1.0 / 4.0; f = 11.0 / 14.0; g = 1.0 / 14.0; h = (-148.0 / 1875.0) / 6.0
        x = [(e, e, e), (f, g, g), (g, f, g), (g, g, f), (g, g, g), (a, a, b), (a, b, a), (a, b, b), (b, a, a), (b, a, b), (b, b, a)]
        w = [h, c, c, c, c, d, d, d, d, d, d]
    elif degree == 5:
        a = 0.0673422422100983; b = 0.3108859192633005; c = 0.7217942490673264; d = 0.0927352503108912
        e = 0.4544962958743506; f = 0.0455037041256494
        p = 0.1126879257180162 / 6.0; q = 0.0734930431163619 / 6.0; r = 0.0425460207770812 / 6.0
        x = [(a, b, b), (b, a, b), (b, b, a), (b, b, b), (c, d, d), (d, c, d), (d, d, c), (d, d, d),
             (e, e, f), (e, f, e), (e, f, f), (f, e, e), (f, e, f), (f, f, e)]
        w = [p, p, p, p, q, q, q, q, r, r, r, r, r, r]
    else:
        raise ValueError(degree)
    return np.array(x, dtype=np.float64), np.array(w, dtype=np.float64)


# ---------------------------------------------------------------------------
# StaticArrays closed forms
# ---------------------------------------------------------------------------
def _det(J):
    """StaticArrays det: 2x2 a11*a22 - a12*a21 ; 3x3 x0 . (x1 × x2) by columns."""
    D = J.shape[-1]
    if D == 1:
        return J[..., 0, 0]
    if D == 2:
        return J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    x0, x1, x2 = J[..., :, 0], J[..., :, 1], J[..., :, 2]
    c0 = x1[..., 1] * x2[..., 2] - x1[..., 2] * x2[..., 1]
    c1 = x1[..., 2] * x2[..., 0] - x1[..., 0] * x2[..., 2]
    c2 = x1[..., 0] * x2[..., 1] - x1[..., 1] * x2[..., 0]
    return x0[..., 0] * c0 + x0[..., 1] * c1 + x0[..., 2] * c2


def _matmul_small(A, B):
    """SMatrix product, sequential k-sum."""
    n, m, k = A.shape[-2], B.shape[-1], A.shape[-1]
    out = np.zeros(A.shape[:-2] + (n, m))
    for i in range(n):
        for j in range(m):
            acc = A[..., i, 0] * B[..., 0, j]
            for kk in range(1, k):
                acc = acc + A[..., i, kk] * B[..., kk, j]
            out[..., i, j] = acc
    return out


def change_of_measure(J):
    """quadrature.jl:4-6: sqrt(det(transpose(J)*J))."""
    Jt = np.swapaxes(J, -1, -2)
    return np.sqrt(_det(_matmul_small(Jt, J)))


def _solve(a, b):
    """StaticArrays ``a \\ b`` for 2x2 / 3x3 (Cramer / adjugate closed form, one det)."""
    D = a.shape[-1]
    d = _det(a)
    if D == 1:
        return b / a[..., 0, :]
    if D == 2:
        return np.stack([(a[..., 1, 1] * b[..., 0] - a[..., 0, 1] * b[..., 1]) / d,
                         (a[..., 0, 0] * b[..., 1] - a[..., 1, 0] * b[..., 0]) / d], axis=-1)
    A = lambda i, j: a[..., i - 1, j - 1]
    B = lambda i: b[..., i - 1]
    r1 = ((A(2, 2) * A(3, 3) - A(2, 3) * A(3, 2)) * B(1) + (A(1, 3) * A(3, 2) - A(1, 2) * A(3, 3)) * B(2)
          + (A(1, 2) * A(2, 3) - A(1, 3) * A(2, 2)) * B(3)) / d
    r2 = ((A(2, 3) * A(3, 1) - A(2, 1) * A(3, 3)) * B(1) + (A(1, 1) * A(3, 3) - A(1, 3) * A(3, 1)) * B(2)
          + (A(1, 3) * A(2, 1) - A(1, 1) * A(2, 3)) * B(3)) / d
    r3 = ((A(2, 1) * A(3, 2) - A(2, 2) * A(3, 1)) * B(1) + (A(1, 2) * A(3, 1) - A(1, 1) * A(3, 2)) * B(2)
          + (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * B(3)) / d
    return np.stack([r1, r2, r3], axis=-1)


def _dot(a, b):
    acc = a[..., 0] * b[..., 0]
    for k in range(1, a.shape[-1]):
        acc = acc + a[..., k] * b[..., k]
    return acc


# ---------------------------------------------------------------------------
# Per-point geometry (accessors.jl:941-968, 1000-1007, 1365-1368)
# ---------------------------------------------------------------------------
def point_geometry(coords, cell_nodes, dM_q):
    """J = Σ_node x_node ⊗ ∇̂M_node (sequential in local-node order), for all cells."""
    nc, nln = cell_nodes.shape
    D = coords.shape[1]
    J = np.zeros((nc, D, dM_q.shape[1]))
    for n in range(nln):
        x = coords[cell_nodes[:, n] - 1]                 # [nc, D]
        J = J + x[:, :, None] * dM_q[n][None, None, :]   # outer(x, ref∇m)  (quadrature.jl:2)
    return J


# ---------------------------------------------------------------------------
# Forms: integrand(r, c, q) following compiler.jl:1879-1894 / A.8b
# ---------------------------------------------------------------------------
LAPLACE, MASS, ELASTICITY = 1, 2, 3
SOURCE_CONST, SOURCE_NODAL, SOURCE_QP = 101, 102, 103


def element_matrices(form, coords, cell_nodes, tab, n_comp=1, alpha=1.0, lam=1.0, mu=1.0, coef_nodal=None, coef_qp=None):
    """be[cell, r, c] = Σ_q ((α·integrand(u=φ_r, v=φ_c))·dV_q), q outermost, then c, then r
    (compiler.jl:1865-1900; problems.jl:55-63).
    coef_nodal / coef_qp: scalar coefficient κ(x_q) multiplying the integrand (∫ κ ∇u·∇v, ∫ κ u v): a nodal field
    interpolated with the geometry functions (DiscreteField parameter, accessors.jl:1489-1563: Σ_node κ_node M_node(ξ_q),
    sequential) or host-sampled values per (cell, point) (AnalyticalField, field.jl:17-58)."""
    w, N, dN, dM = tab["w"], tab["N"], tab["dN"], tab["dM"]
    nc = cell_nodes.shape[0]
    nq, nls = N.shape
    D = coords.shape[1]
    nld = nls * n_comp
    be = np.zeros((nc, nld, nld))
    for q in range(nq):
        J = point_geometry(coords, cell_nodes, dM[q])
        dV = change_of_measure(J) * w[q]                                        # accessors.jl:1000-1007
        kq = None
        if coef_qp is not None:
            kq = np.asarray(coef_qp, dtype=np.float64).reshape(nc, nq)[:, q]
        elif coef_nodal is not None:
            kq = np.zeros(nc)
            kn = np.asarray(coef_nodal, dtype=np.float64).reshape(-1)
            for n in range(cell_nodes.shape[1]):
                kq = kq + kn[cell_nodes[:, n] - 1] * tab["M"][q, n]
        Jt = np.swapaxes(J, -1, -2)
        if form in (LAPLACE, ELASTICITY):
            g = [_solve(Jt, np.broadcast_to(dN[q, a], (nc, D))) for a in range(nls)]   # accessors.jl:1365-1368
        for c in range(nld):
            for r in range(nld):
                a, i = divmod(r, n_comp)      # dof = (node-1)*n_comp + comp  (space.jl:1267-1271)
                b, j = divmod(c, n_comp)
                if form == LAPLACE:           # ∇u:∇v = δ_ij ∇s_a·∇s_b for vector-valued spaces
                    t = _dot(g[a], g[b]) if i == j else np.zeros(nc)
                    v = (alpha * (t if kq is None else kq * t)) * dV
                elif form == MASS:            # u·v = δ_ij s_a s_b
                    t = (N[q, a] * N[q, b]) if i == j else 0.0
                    v = (alpha * (t if kq is None else kq * t)) * dV
                elif form == ELASTICITY:
                    # σ(ε(u)):ε(v), u = s_a e_i, v = s_b e_j, isotropic (λ, μ)
                    t = lam * (g[a][:, i] * g[b][:, j]) + mu * (g[a][:, j] * g[b][:, i])
                    if i == j:
                        t = t + mu * _dot(g[a], g[b])
                    v = (alpha * t) * dV
                else:
                    raise ValueError(form)
                be[:, r, c] += v
    return be


def element_vectors(form, coords, cell_nodes, tab, n_comp=1, alpha=1.0, f_const=None, f_nodal=None, f_qp=None):
    """be[cell, i] = Σ_q ((α·(f(x_q)·φ_i))·dV_q)  (compiler.jl:1958-1979)."""
    w, N, M, dM = tab["w"], tab["N"], tab["M"], tab["dM"]
    nc = cell_nodes.shape[0]
    nq, nls = N.shape
    nld = nls * n_comp
    be = np.zeros((nc, nld))
    for q in range(nq):
        J = point_geometry(coords, cell_nodes, dM[q])
        dV = change_of_measure(J) * w[q]
        if form == SOURCE_CONST:
            f = np.broadcast_to(np.asarray(f_const, dtype=np.float64).reshape(1, -1), (nc, n_comp))
        elif form == SOURCE_NODAL:   # f_h(x_q) = Σ_node f_node M_node(ξ_q), sequential
            f = np.zeros((nc, n_comp))
            fn = np.asarray(f_nodal, dtype=np.float64).reshape(coords.shape[0], n_comp)
            for n in range(cell_nodes.shape[1]):
                f = f + fn[cell_nodes[:, n] - 1] * M[q, n]
        elif form == SOURCE_QP:
            f = np.asarray(f_qp, dtype=np.float64).reshape(nc, nq, n_comp)[:, q, :]
        else:
            raise ValueError(form)
        for i in range(nld):
            a, comp = divmod(i, n_comp)
            be[:, i] += (alpha * (f[:, comp] * N[q, a])) * dV
    return be


# ---------------------------------------------------------------------------
# Forms with a DiscreteField parameter u_h (nonlinear problems: residual / Jacobian re-assembly,
# problems.jl:276-285, 352-361, 465-497; field evaluation accessors.jl:1489-1563)
# ---------------------------------------------------------------------------
PLAPLACE_JACOBIAN = 4
PLAPLACE_RESIDUAL = 104
SCALAR_VOLUME, SCALAR_L2SQ, SCALAR_H1SQ = 200, 201, 202


def field_cell_values(cell_dofs, free_values, dirichlet_values):
    """values(::NewDiscreteFieldFace{AtInterior}) accessors.jl:1489-1510: per cell, dof > 0 reads free_values[dof],
    dof < 0 reads dirichlet_values[-dof].  -> [n_cells, n_ldofs]"""
    d = np.asarray(cell_dofs, dtype=np.int64)
    fv = np.asarray(free_values, dtype=np.float64)
    dv = np.asarray(dirichlet_values, dtype=np.float64)
    out = np.empty(d.shape)
    pos = d > 0
    out[pos] = fv[d[pos] - 1]
    out[~pos] = dv[-d[~pos] - 1]
    return out


def _physical_gradients(coords, cell_nodes, tab, q):
    """(g[a] = Jᵀ \\ ∇̂N_a for every local shape function, dV) at point q (accessors.jl:941-1007, 1365-1368)"""
    nc = cell_nodes.shape[0]
    D = coords.shape[1]
    J = point_geometry(coords, cell_nodes, tab["dM"][q])
    dV = change_of_measure(J) * tab["w"][q]
    Jt = np.swapaxes(J, -1, -2)
    g = [_solve(Jt, np.broadcast_to(tab["dN"][q, a], (nc, D))) for a in range(tab["N"].shape[1])]
    return g, dV


def _field_gradient(uc, g):
    """field(gradient, ::NewDiscreteFieldFace) accessors.jl:1549-1556: sum(i->x[i]*s[i], 1:n; init=zero) — sequential"""
    gu = np.zeros_like(g[0])
    for a in range(len(g)):
        gu = gu + uc[:, a, None] * g[a]
    return gu


def _norm(v):
    """LinearAlgebra.norm of an SVector: sqrt of the sequential sum of abs2"""
    return np.sqrt(_dot(v, v))


def plaplace_flux(gu, q):
    """flux(∇u) = norm(∇u)^(q-2) * ∇u  (test/problems_ext_tests.jl:160, test/assembly_tests.jl:678)"""
    return (_norm(gu) ** (q - 2))[:, None] * gu


def plaplace_dflux(gdu, gu, q):
    """dflux(∇du,∇u) = (q-2)*norm(∇u)^(q-4)*(∇u⋅∇du)*∇u + norm(∇u)^(q-2)*∇du  (test/problems_ext_tests.jl:161),
    evaluated left to right as Julia parses it"""
    n = _norm(gu)
    t1 = (((q - 2) * n ** (q - 4)) * _dot(gu, gdu))[:, None] * gu
    t2 = (n ** (q - 2))[:, None] * gdu
    return t1 + t2


def element_vectors_plaplace_residual(coords, cell_nodes, cell_dofs, tab, free_values, dirichlet_values, q,
                                      alpha=1.0, f_const=1.0, f_qp=None):
    """be[cell, i] = Σ_p ((α·(∇φ_i⋅flux(∇u_h) − f φ_i))·dV_p): the residual of test/problems_ext_tests.jl:162
    (`∇(v,x)⋅GT.call(flux,∇(u,x)) - v(x)`, f ≡ 1) and of test/assembly_tests.jl:685 (`- f(x)*v(x)`, f sampled)."""
    nc = cell_nodes.shape[0]
    nq, nls = tab["N"].shape
    uc = field_cell_values(cell_dofs, free_values, dirichlet_values)
    be = np.zeros((nc, nls))
    for p in range(nq):
        g, dV = _physical_gradients(coords, cell_nodes, tab, p)
        fl = plaplace_flux(_field_gradient(uc, g), q)
        f = f_const if f_qp is None else np.asarray(f_qp, dtype=np.float64).reshape(nc, nq)[:, p]
        for i in range(nls):
            be[:, i] += (alpha * (_dot(g[i], fl) - f * tab["N"][p, i])) * dV
    return be


def element_matrices_plaplace_jacobian(coords, cell_nodes, cell_dofs, tab, free_values, dirichlet_values, q, alpha=1.0):
    """be[cell, r, c] = Σ_p ((α·(∇φ_c⋅dflux(∇φ_r, ∇u_h)))·dV_p), du = φ_r, v = φ_c (the form is symmetric
    in (du, v) up to rounding) — the Jacobian of test/problems_ext_tests.jl:163."""
    nc = cell_nodes.shape[0]
    nq, nls = tab["N"].shape
    uc = field_cell_values(cell_dofs, free_values, dirichlet_values)
    be = np.zeros((nc, nls, nls))
    for p in range(nq):
        g, dV = _physical_gradients(coords, cell_nodes, tab, p)
        gu = _field_gradient(uc, g)
        for c in range(nls):
            for r in range(nls):
                be[:, r, c] += (alpha * _dot(g[c], plaplace_dflux(g[r], gu, q))) * dV
    return be


def assemble_scalar_field(kind, coords, cell_nodes, cell_dofs, tab, free_values, dirichlet_values, g_qp=None):
    """assemble_scalar (problems.jl:173-190; loop compiler.jl:1073-1083): s = init; for face, for point: s += integrand·dV.
    kind: SCALAR_VOLUME ∫1, SCALAR_L2SQ ∫abs2(u_h − g), SCALAR_H1SQ ∫(∇u_h − ∇g)⋅(∇u_h − ∇g); g sampled at the points
    (g_qp [n_cells][n_q] values or [n_cells][n_q][D] gradients), absent = 0.  The sum over (cell, point) is returned
    with numpy's pairwise summation — the order of this reduction is not something 1e-12 can see."""
    nc = cell_nodes.shape[0]
    nq, nls = tab["N"].shape
    uc = field_cell_values(cell_dofs, free_values, dirichlet_values)
    vals = np.zeros((nc, nq))
    for p in range(nq):
        g, dV = _physical_gradients(coords, cell_nodes, tab, p)
        if kind == SCALAR_VOLUME:
            t = np.ones(nc)
        elif kind == SCALAR_L2SQ:
            u = np.zeros(nc)
            for a in range(nls):
                u = u + uc[:, a] * tab["N"][p, a]
            if g_qp is not None:
                u = u - np.asarray(g_qp, dtype=np.float64).reshape(nc, nq)[:, p]
            t = u * u
        elif kind == SCALAR_H1SQ:
            gu = _field_gradient(uc, g)
            if g_qp is not None:
                gu = gu - np.asarray(g_qp, dtype=np.float64).reshape(nc, nq, -1)[:, p, :]
            t = _dot(gu, gu)
        else:
            raise ValueError(kind)
        vals[:, p] = t * dV
    return float(vals.sum())


def space_dof_coordinates(coords, cell_nodes, cell_dofs, n_free, n_dirichlet, M_at_nodes, n_comp=1):
    """node_coordinates(::LagrangeMeshSpace) (space.jl:1876-1897) seen through free_dof_node / dirichlet_dof_node
    (space.jl:1960-1998), loop for loop: for every cell in order, every local node: x = zero; for lmnode: x +=
    tab[lnode,lmnode]*x_mnode; the LAST cell holding the node wins.  -> (x_free [n_free,D], x_dirichlet [n_dirichlet,D])"""
    D = coords.shape[1]
    xf, xd = np.zeros((n_free, D)), np.zeros((n_dirichlet, D))
    for cell in range(cell_nodes.shape[0]):
        mnodes = cell_nodes[cell]
        for ldof in range(cell_dofs.shape[1]):
            lnode = ldof // n_comp
            x = np.zeros(D)
            for m in range(len(mnodes)):
                x = x + M_at_nodes[lnode, m] * coords[mnodes[m] - 1]
            dof = cell_dofs[cell, ldof]
            if dof > 0:
                xf[dof - 1] = x
            else:
                xd[-dof - 1] = x
    return xf, xd


def assemble_vector_plaplace_residual(coords, cell_nodes, cell_dofs, n_free, tab, free_values, dirichlet_values, q, **kw):
    be = element_vectors_plaplace_residual(coords, cell_nodes, cell_dofs, tab, free_values, dirichlet_values, q, **kw)
    I, V = coo_vector(be, cell_dofs, FREE)
    return dense_vector(I, V, n_free)


def assemble_matrix_plaplace_jacobian(coords, cell_nodes, cell_dofs, n_free, tab, free_values, dirichlet_values, q, **kw):
    be = element_matrices_plaplace_jacobian(coords, cell_nodes, cell_dofs, tab, free_values, dirichlet_values, q, **kw)
    I, J, V = coo_matrix(be, cell_dofs, cell_dofs, FREE, FREE)
    return sparse_csc(I, J, V, n_free, n_free)


# ---------------------------------------------------------------------------
# Scatter + compression
# ---------------------------------------------------------------------------
def _skip(d, fd):
    """assembly.jl:155-157."""
    return ((d > 0) & (fd == DIRICHLET)) | ((d < 0) & (fd == FREE))


def coo_matrix(be, cell_dofs_rows, cell_dofs_cols, fd_rows=FREE, fd_cols=FREE):
    """contribute!(::MatrixAllocation) assembly.jl:189-208 → COO push :545-556.
    Order: cell-major; within a cell column j outer, row i inner; skipped by sign."""
    nc, ni, nj = be.shape
    I = np.broadcast_to(cell_dofs_rows[:, :, None], (nc, ni, nj))
    Jd = np.broadcast_to(cell_dofs_cols[:, None, :], (nc, ni, nj))
    # memory order must be (cell, j, i): transpose to [cell, j, i]
    I = np.transpose(I, (0, 2, 1)).reshape(-1)
    Jd = np.transpose(Jd, (0, 2, 1)).reshape(-1)
    V = np.transpose(be, (0, 2, 1)).reshape(-1)
    keep = ~(_skip(I, fd_rows) | _skip(Jd, fd_cols))
    mi = 1 if fd_rows == FREE else -1
    mj = 1 if fd_cols == FREE else -1
    return (mi * I[keep]).astype(np.int32), (mj * Jd[keep]).astype(np.int32), V[keep]


def sparse_csc(I, J, V, m, n):
    """Julia sparse(I,J,V,m,n) via PartitionedArrays.sparse_matrix (assembly.jl:571-575):
    CSC, rows sorted inside each column, duplicates summed left-to-right in input order,
    explicit zeros kept.  Returns 1-based Int32 colptr/rowval and Float64 nzval."""
    I = np.asarray(I, dtype=np.int64)
    J = np.asarray(J, dtype=np.int64)
    key = (J - 1) * m + (I - 1)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    vs = np.asarray(V, dtype=np.float64)[order]
    head = np.ones(ks.size, dtype=bool)
    head[1:] = ks[1:] != ks[:-1]
    starts = np.flatnonzero(head)
    nnz = starts.size
    lens = np.diff(np.append(starts, ks.size))
    nzval = np.zeros(nnz)
    if nnz:
        nzval[:] = vs[starts]
        for k in range(1, int(lens.max())):
            sel = lens > k
            nzval[sel] = nzval[sel] + vs[starts[sel] + k]     # sequential, input order
    ukeys = ks[starts]
    rowval = (ukeys % m + 1).astype(np.int32)
    cols = ukeys // m
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, cols + 1, 1)
    colptr = (np.cumsum(colptr) + 1).astype(np.int32)
    return colptr, rowval, nzval


def coo_vector(be, cell_dofs, fd=FREE):
    """contribute!(::VectorAllocation) assembly.jl:175-187 → :535-543."""
    I = cell_dofs.reshape(-1)
    V = be.reshape(-1)
    keep = ~_skip(I, fd)
    m = 1 if fd == FREE else -1
    return (m * I[keep]).astype(np.int32), V[keep]


def dense_vector(I, V, n, b0=None):
    """PartitionedArrays.dense_vector (assembly.jl:558-569): b[i] += v sequentially in COO order.
    b0: sums already accumulated from the COO entries of EARLIER integrals of the same linear form (a sum of integrals
    pushes into one COO vector, contribution after contribution, problems.jl:258-266), so continuing from b0 is the
    same left-to-right sum."""
    I = np.asarray(I, dtype=np.int64) - 1
    order = np.argsort(I, kind="stable")
    Is, Vs = I[order], np.asarray(V, dtype=np.float64)[order]
    b = np.zeros(n) if b0 is None else np.array(b0, dtype=np.float64, copy=True)
    if Is.size == 0:
        return b
    head = np.ones(Is.size, dtype=bool)
    head[1:] = Is[1:] != Is[:-1]
    starts = np.flatnonzero(head)
    lens = np.diff(np.append(starts, Is.size))
    acc = b[Is[starts]] + Vs[starts]               # 0.0 + v is exact
    for k in range(1, int(lens.max())):
        sel = lens > k
        acc[sel] = acc[sel] + Vs[starts[sel] + k]
    b[Is[starts]] = acc
    return b


# ---------------------------------------------------------------------------
# Whole path: assemble_matrix / assemble_vector (problems.jl:244-274, 319-350)
# ---------------------------------------------------------------------------
def assemble_matrix(form, coords, cell_nodes, cell_dofs, n_free, n_dirichlet, tab, n_comp=1,
                    free_or_dirichlet=(FREE, FREE), **params):
    be = element_matrices(form, coords, cell_nodes, tab, n_comp=n_comp, **params)
    fr, fcn = free_or_dirichlet
    I, J, V = coo_matrix(be, cell_dofs, cell_dofs, fr, fcn)
    m = n_free if fr == FREE else n_dirichlet
    n = n_free if fcn == FREE else n_dirichlet
    return sparse_csc(I, J, V, m, n)


def assemble_vector(form, coords, cell_nodes, cell_dofs, n_free, n_dirichlet, tab, n_comp=1,
                    free_or_dirichlet=FREE, b0=None, **params):
    """`cell_nodes`/`cell_dofs`/`tab` may describe the faces of a boundary domain (cells of dimension D-1 embedded in D
    dimensions, tabulated on the reference face): the same loop then integrates ∫_Γ g v dΓ, dV = sqrt(det(JᵀJ)) w with
    the D x (D-1) Jacobian (quadrature.jl:4-6)."""
    be = element_vectors(form, coords, cell_nodes, tab, n_comp=n_comp, **params)
    I, V = coo_vector(be, cell_dofs, free_or_dirichlet)
    return dense_vector(I, V, n_free if free_or_dirichlet == FREE else n_dirichlet, b0)


# ---------------------------------------------------------------------------
# linear-problem right-hand side: mul!(b, Ad, xd, -1, 1)  (problems.jl:447)
# ---------------------------------------------------------------------------
def spmatmul_add(colptr, rowval, nzval, x, alpha, beta, b):
    """Julia's 5-argument mul!(C, A::SparseMatrixCSC, B, alpha, beta) for a vector B, as SparseArrays implements it
    (stdlib `_spmatmul!`, not vendored in /root/reference; semantics restated): C is scaled by beta first (left
    alone for beta == 1, zero-filled for beta == 0), then for each column k in increasing order axk = B[k]*alpha and
    every stored entry j of the column does C[rowval[j]] += nzval[j]*axk — multiply and add rounded separately.
    1-based colptr/rowval as in the reference.  Returns the updated copy of b."""
    c = np.array(b, dtype=np.float64, copy=True)
    if beta != 1:
        c = c * beta if beta != 0 else np.zeros_like(c)
    cp = np.asarray(colptr, dtype=np.int64) - 1
    rv = np.asarray(rowval, dtype=np.int64) - 1
    nz = np.asarray(nzval, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    for k in range(cp.size - 1):
        lo, hi = cp[k], cp[k + 1]
        if hi > lo:
            axk = x[k] * alpha
            c[rv[lo:hi]] = c[rv[lo:hi]] + nz[lo:hi] * axk     # rows inside one column are distinct
    return c


def linear_problem_rhs(form_a, form_l, coords, cell_nodes, cell_dofs, n_free, n_dirichlet, tab, xd, n_comp=1,
                       params_a=None, params_l=None):
    """A, Ad, b of assemble_matrix_and_vector_with_free_and_dirichlet_columns (problems.jl:413-430) followed by
    b <- b - Ad*xd (problems.jl:447)."""
    params_a, params_l = params_a or {}, params_l or {}
    A = assemble_matrix(form_a, coords, cell_nodes, cell_dofs, n_free, n_dirichlet, tab, n_comp=n_comp,
                        free_or_dirichlet=(FREE, FREE), **params_a)
    Ad = assemble_matrix(form_a, coords, cell_nodes, cell_dofs, n_free, n_dirichlet, tab, n_comp=n_comp,
                         free_or_dirichlet=(FREE, DIRICHLET), **params_a)
    b = assemble_vector(form_l, coords, cell_nodes, cell_dofs, n_free, n_dirichlet, tab, n_comp=n_comp, **params_l)
    return A, Ad, spmatmul_add(Ad[0], Ad[1], Ad[2], xd, -1.0, 1.0, b)


# ---------------------------------------------------------------------------
# Literal face complex + order-k Lagrange dof numbering on quad / hex meshes
# (python loops, small meshes; what `lagrange_space(Ω, k; dirichlet_boundary)` numbers for k >= 1)
# ---------------------------------------------------------------------------
def _cube_lfaces(n, d):
    """local d-faces of the reference n-cube as lists of 1-based local vertices (domain.jl:188-255)"""
    if d == n:
        return [list(range(1, 2 ** n + 1))]
    if n == 1:
        return [[1], [2]]
    return CUBE_FACES[n][d]


def parent_boundary_faces(cell_nodes, nnodes, D, d):
    """cartesian_mesh.jl:117-168 for one d: boundary d-faces cell-major / local-face order; -> (nodes, group = ldface)"""
    node_to_n = [0] * (nnodes + 1)
    for nodes in cell_nodes:
        for n in nodes:
            node_to_n[n] += 1
    nmax = 2 ** d
    faces, groups = [], []
    for nodes in cell_nodes:
        for ldface, lnodes in enumerate(_cube_lfaces(D, d), start=1):
            if all(node_to_n[nodes[ln - 1]] <= nmax for ln in lnodes):
                faces.append([nodes[ln - 1] for ln in lnodes])
                groups.append(ldface)
    return faces, groups


SIMPLEX_LFACES = {   # domain.jl:389-470 (face_nodes of mesh(::UnitSimplex{n}))
    1: {0: [[1], [2]]},
    2: {0: [[1], [2], [3]], 1: [[1, 2], [1, 3], [2, 3]]},
    3: {0: [[1], [2], [3], [4]], 1: [[1, 2], [1, 3], [2, 3], [1, 4], [2, 4], [3, 4]], 2: [[1, 2, 3], [1, 2, 4], [1, 3, 4], [2, 3, 4]]},
}


def _simplex_lfaces(n, d):
    return [list(range(1, n + 2))] if d == n else SIMPLEX_LFACES[n][d]


def _complex(cell_vertices, D, lfaces, parents, n_vertices):
    """complexify's face generation (generate_face_boundary topology.jl:1594-1704 + generate_face_vertices :1468-1540):
    d-faces from (d+1)-faces, highest d first; pre-existing faces `parents[d]` (vertex lists) keep ids 1.., every other
    face gets the next id at its first encounter looping over the (d+1)-faces in id order and their local faces in
    reference order, with the vertex order seen from that first parent.  -> vertices[d], cell_faces[d]."""
    vertices = {D: [list(v) for v in cell_vertices]}
    nface_dfaces = {}
    for d in range(D - 1, 0, -1):
        n = d + 1
        ids = {frozenset(v): i + 1 for i, v in enumerate(parents.get(d, []))}            # same_valid_ids: equal vertex sets
        assert len(ids) == len(parents.get(d, [])), "pre-existing faces are not pairwise distinct"
        verts = [list(v) for v in parents.get(d, [])]
        inc = []
        for nv in vertices[n]:
            row = []
            for lv in lfaces(n, d):
                fv = [nv[i - 1] for i in lv]
                key = frozenset(fv)
                if key not in ids:
                    ids[key] = len(verts) + 1
                    verts.append(fv)
                row.append(ids[key])
            inc.append(row)
        vertices[d] = verts
        nface_dfaces[(n, d)] = inc
    vertices[0] = [[v] for v in range(1, n_vertices + 1)]
    cell_faces = {}
    for d in range(0, D + 1):                                                # face_incidence(topology, D, d)
        if d == D:
            cell_faces[d] = [[c + 1] for c in range(len(cell_vertices))]
        elif d == 0:
            cell_faces[d] = [list(v) for v in vertices[D]]
        elif (D, d) in nface_dfaces:
            cell_faces[d] = nface_dfaces[(D, d)]
        else:                                                                # non-adjacent dimensions: match vertex sets
            ids = {frozenset(v): i + 1 for i, v in enumerate(vertices[d])}
            cell_faces[d] = [[ids[frozenset(cv[i - 1] for i in lv)] for lv in lfaces(D, d)] for cv in vertices[D]]
    return vertices, cell_faces


def face_complex(cell_nodes, nnodes, D):
    """complexify (topology.jl:1034-1125, 1468-1540, 1594-1704) of `cartesian_mesh` output (not simplexified).
    -> dict with, per dimension d: vertices[d][face] (vertex lists), cell_faces[d][cell] (global ids in the reference
    cell's local order), n_parent[d] / group[d] (pre-existing boundary faces keep ids 1..n_parent, group = local face id)."""
    if len(cell_nodes) == 1:
        pre0 = list(cell_nodes[0])
    else:
        pre0 = [f[0] for f in parent_boundary_faces(cell_nodes, nnodes, D, 0)[0]]
    node_vertex = vertex_ids(cell_nodes, nnodes, pre0)
    parents, n_parent, group = {}, {}, {}
    for d in range(1, D):
        pfaces, pgroups = parent_boundary_faces(cell_nodes, nnodes, D, d)
        parents[d] = [[node_vertex[x - 1] for x in f] for f in pfaces]
        n_parent[d], group[d] = len(pfaces), pgroups
    vertices, cell_faces = _complex([[node_vertex[n - 1] for n in nodes] for nodes in cell_nodes], D, _cube_lfaces, parents,
                                    max(node_vertex))
    return dict(node_vertex=node_vertex, vertices=vertices, cell_faces=cell_faces, n_parent=n_parent, group=group)


def simplexified_unit_cube(D):
    """simplexify(::UnitNCube) (domain.jl:270-320): the 2 / 6 simplices of the cube (:322-336) complexified WITHOUT
    pre-existing faces (vertex id = node id), and per cube-local d-face the simplex d-faces lying in it, ascending ids.
    -> face_nodes[d] (1-based cube-local nodes), groups[d][cube ldface-1] = [simplex d-face ids]"""
    cells = SIMPLEX_NODES[D]
    vertices, _ = _complex(cells, D, _simplex_lfaces, {}, 2 ** D)
    groups = {}
    for d in range(0, D):
        groups[d] = []
        for cnodes in _cube_lfaces(D, d):
            groups[d].append([i + 1 for i, sn in enumerate(vertices[d]) if all(n in cnodes for n in sn)])
    return vertices, groups


def simplex_parent_boundary_faces(hex_cell_nodes, nnodes, D, d):
    """structured_simplex_mesh_with_boundary (cartesian_mesh.jl:330-461) for one d: for every boundary local d-face of
    every HEX cell (same test as the hex mesh, on the hex chain), its simplex sub-faces in ascending id."""
    ref_nodes, ref_groups = simplexified_unit_cube(D)
    node_to_n = [0] * (nnodes + 1)
    for nodes in hex_cell_nodes:
        for n in nodes:
            node_to_n[n] += 1
    nmax = 2 ** d
    faces, groups = [], []
    for nodes in hex_cell_nodes:
        for ldface, lnodes in enumerate(_cube_lfaces(D, d), start=1):
            if all(node_to_n[nodes[ln - 1]] <= nmax for ln in lnodes):
                for sface in ref_groups[d][ldface - 1]:
                    faces.append([nodes[ln - 1] for ln in ref_nodes[d][sface - 1]])
                    groups.append(ldface)
    return faces, groups


def simplex_face_complex(domain, cells_per_dir):
    """complexify of cartesian_mesh(domain, cells; simplexify=true)"""
    D = len(cells_per_dir)
    coords, hex_cells = cartesian_chain(domain, cells_per_dir, False)
    _, cell_nodes = cartesian_chain(domain, cells_per_dir, True)
    nn = coords.shape[0]
    pre0 = [f[0] for f in simplex_parent_boundary_faces(hex_cells, nn, D, 0)[0]]
    node_vertex = vertex_ids(cell_nodes, nn, pre0)
    parents, n_parent, group = {}, {}, {}
    for d in range(1, D):
        pfaces, pgroups = simplex_parent_boundary_faces(hex_cells, nn, D, d)
        parents[d] = [[node_vertex[x - 1] for x in f] for f in pfaces]
        n_parent[d], group[d] = len(pfaces), pgroups
    vertices, cell_faces = _complex([[node_vertex[n - 1] for n in nodes] for nodes in cell_nodes], D, _simplex_lfaces, parents,
                                    max(node_vertex))
    return coords, cell_nodes, dict(node_vertex=node_vertex, vertices=vertices, cell_faces=cell_faces, n_parent=n_parent, group=group)


def _vertex_permutations(d):
    """domain.jl:53-100: admissible vertex permutations of the unit d-cube, Combinatorics.permutations order;
    only the identity for d > 2 and d == 0."""
    if d == 0:
        return [[1]]
    if d > 2:
        return [list(range(1, 2 ** d + 1))]
    if d == 1:
        return [[1, 2], [2, 1]]
    # unit square, vertices (0,0),(1,0),(0,1),(1,1): |det J| at the centre must equal the reference area
    X = [(0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0, 1.0)]
    out = []
    for p in itertools.permutations(range(1, 5)):
        Y = [X[i - 1] for i in p]
        # J at (1/2,1/2) of the bilinear map: dx/da = ((Y2-Y1)+(Y4-Y3))/2, dx/db = ((Y3-Y1)+(Y4-Y2))/2
        ja = [((Y[1][k] - Y[0][k]) + (Y[3][k] - Y[2][k])) / 2 for k in range(2)]
        jb = [((Y[2][k] - Y[0][k]) + (Y[3][k] - Y[1][k])) / 2 for k in range(2)]
        if abs(abs(ja[0] * jb[1] - ja[1] * jb[0]) - 1.0) < 1e-12:
            out.append(list(p))
    return out


def _lattice(d, k, interior=False):
    """multi-indices of the order-k d-cube element's nodes, first index fastest (space.jl:1127-1177)"""
    rng = range(1, k) if interior else range(0, k + 1)
    return [tuple(reversed(t)) for t in itertools.product(*[rng] * d)]


def _q1_map(t, k, corners):
    """Σ_v M_v(t/k) X_v for the d-cube with vertex coordinates `corners` (2^d tuples), returned scaled by k (integers)"""
    d = len(t)
    out = [0.0] * len(corners[0])
    for v, X in enumerate(corners):
        w = 1.0
        for m in range(d):
            w *= (t[m] / k) if (v >> m) & 1 else (1.0 - t[m] / k)
        for c in range(len(X)):
            out[c] += w * X[c]
    return tuple(int(round(k * x)) for x in out)


def _simplex_lattice(d, k, interior=False):
    """exponents of the order-k d-simplex element's nodes (sum <= k, first index fastest; space.jl:1127-1177);
    interior: every barycentric coordinate >= 1"""
    out = []
    for t in itertools.product(*[range(k + 1)] * d):
        e = tuple(reversed(t))
        if sum(e) > k:
            continue
        if interior and (any(x < 1 for x in e) or sum(e) > k - 1):
            continue
        out.append(e)
    return out


def _p1_map(t, k, corners):
    """k * Σ_v M_v(t/k) X_v for the d-simplex with vertex coordinates `corners` (d+1 tuples): barycentric map, integers"""
    X0 = corners[0]
    out = [k * x for x in X0]
    for m, tm in enumerate(t):
        for c in range(len(X0)):
            out[c] += tm * (corners[m + 1][c] - X0[c])
    return tuple(int(round(x)) for x in out)


def _simplex_vertex_permutations(d):
    """domain.jl:53-77: every permutation is admissible for a simplex (Combinatorics.permutations order), identity for d > 2"""
    if d == 0:
        return [[1]]
    if d > 2:
        return [list(range(1, d + 2))]
    return [list(p) for p in itertools.permutations(range(1, d + 2))]


def reference_face_tables(D, k, n_comp=1, simplex=False):
    """Per dimension d: for every local d-face of the order-k reference element (D-cube, or D-simplex)
    dofs[d][ldface]      all local dofs on the face (face_dofs, space.jl:1343-1360, 1488-1510)
    own[d][ldface]       its own (interior) local dofs (face_own_dofs, :1374-1392)
    perms[d][ldface]     own-dof permutation per vertex permutation id (face_own_dof_permutations, :1439-1487, 1512-1575)
    local dof = (node-1)*n_comp + c, node-major / component-minor."""
    if simplex:
        lattice, fmap, lfaces, vpf = _simplex_lattice, _p1_map, _simplex_lfaces, _simplex_vertex_permutations
        corner = lambda v: tuple(1.0 if v - 2 == m else 0.0 for m in range(D))          # v1 = 0, v_{m+2} = e_m
        unit_of = lambda d: [tuple(1.0 if v - 1 == m else 0.0 for m in range(d)) for v in range(d + 1)]
    else:
        lattice, fmap, lfaces, vpf = _lattice, _q1_map, _cube_lfaces, _vertex_permutations
        corner = lambda v: tuple(float((v - 1) >> m & 1) for m in range(D))
        unit_of = lambda d: [tuple(float((v >> m) & 1) for m in range(d)) for v in range(2 ** d)]
    cell_nodes = {t: i + 1 for i, t in enumerate(lattice(D, k))}
    dofs, own, perms, vperms = {}, {}, {}, {}
    for d in range(D + 1):
        dofs[d], own[d], perms[d] = [], [], []
        unit = unit_of(d)
        inter = lattice(d, k, interior=True) if d > 0 else [()]
        vperms[d] = vpf(d)
        node_perms = []
        for P in vperms[d]:                                 # interior node iq -> interior node located at the permuted map
            pc = [unit[p - 1] for p in P]
            node_perms.append([inter.index(fmap(t, k, pc)) + 1 for t in inter] if d > 0 else [1])
        for lv in lfaces(D, d):
            X = [corner(v) for v in lv]
            allnodes = [cell_nodes[fmap(t, k, X)] for t in (lattice(d, k) if d > 0 else [()])]
            inodes = [cell_nodes[fmap(t, k, X)] for t in inter]
            expand = lambda nodes: [(n - 1) * n_comp + c + 1 for n in nodes for c in range(n_comp)]
            dofs[d].append(expand(allnodes))
            own[d].append(expand(inodes))
            perms[d].append([[(j - 1) * n_comp + c + 1 for j in npm for c in range(n_comp)] for npm in node_perms])
    return dofs, own, perms, vperms


def vperms_by_dim(D):
    return {d: _vertex_permutations(d) for d in range(D + 1)}


def lagrange_space_literal(domain, cells_per_dir, order, dirichlet_sides=None, n_comp=1, simplexify=False):
    """generate_dof_ids (space.jl:299-535) on cartesian_mesh(domain, cells) with the order-k Lagrange cube element:
    dof offsets dimension-major then by global face id (:348-370), own dofs placed through the face's permutation id
    relative to the cell (:380-417, topology.jl:593-666), Dirichlet tagging of ALL dofs of the cell-local (D-1)-faces in Γ
    (:477-511), stable free/Dirichlet partition (:512-524, 910-920)."""
    D = len(cells_per_dir)
    if simplexify:
        coords, cell_nodes, fc = simplex_face_complex(domain, cells_per_dir)
        lfaces_of = _simplex_lfaces
    else:
        coords, cell_nodes = cartesian_chain(domain, cells_per_dir, False)
        fc = face_complex(cell_nodes, coords.shape[0], D)
        lfaces_of = _cube_lfaces
    ldofs, own, perms, vperms = reference_face_tables(D, order, n_comp, simplex=simplexify)
    nld = len((_simplex_lattice if simplexify else _lattice)(D, order)) * n_comp
    # offsets
    offset, ndofs = {}, 0
    for d in range(D + 1):
        nown = len(own[d][0])
        offset[d] = []
        for _ in fc["vertices"][d]:
            offset[d].append(ndofs)
            ndofs += nown
    cell_dofs = [[0] * nld for _ in cell_nodes]
    for d in range(D + 1):
        lfaces = lfaces_of(D, d)
        for cell, cv in enumerate(fc["vertices"][D]):
            for lface, cvertices in enumerate(lfaces):
                face = fc["cell_faces"][d][cell][lface]
                pindex = 0
                if 0 < d < D:                                         # fill_face_permutation_ids! (topology.jl:593-634)
                    fv = fc["vertices"][d][face - 1]
                    for pi, P in enumerate(vperms[d]):
                        if all(fv[P[c] - 1] == cv[cvertices[c] - 1] for c in range(len(cvertices))):
                            pindex = pi
                            break
                    else:
                        raise AssertionError("Valid pindex not found")
                perm = perms[d][lface][pindex]
                for i, own_dof in enumerate(own[d][lface]):
                    cell_dofs[cell][own_dof - 1] = perm[i] + offset[d][face - 1]
    # Dirichlet
    tag = [0] * ndofs
    if dirichlet_sides is not None:
        N = D - 1
        sides = set(range(1, 2 * D + 1)) if dirichlet_sides == "boundary" else set(dirichlet_sides)
        if D == 1:
            raise NotImplementedError
        face_tag = [0] * len(fc["vertices"][N])
        for f in range(fc["n_parent"][N]):                            # parent faces keep ids 1..n_parent; group = ldface
            if fc["group"][N][f] in sides:
                face_tag[f] = 1
        for cell in range(len(cell_nodes)):
            for lface in range(len(lfaces_of(D, N))):
                if face_tag[fc["cell_faces"][N][cell][lface] - 1]:
                    for ld in ldofs[N][lface]:
                        tag[cell_dofs[cell][ld - 1] - 1] = 1
    free = [i for i in range(ndofs) if tag[i] == 0]
    diri = [i for i in range(ndofs) if tag[i] != 0]
    newid = [0] * ndofs
    for i, dof in enumerate(free):
        newid[dof] = i + 1
    for i, dof in enumerate(diri):
        newid[dof] = -(i + 1)
    out = [[newid[x - 1] for x in row] for row in cell_dofs]
    return dict(coords=coords, cell_nodes=np.array(cell_nodes, dtype=np.int32), cell_dofs=np.array(out, dtype=np.int32),
                n_free=len(free), n_dirichlet=len(diri), n_dofs=ndofs, face_complex=fc)


# ---------------------------------------------------------------------------
# Multi-field spaces (CartesianProductSpace) and skeleton integrals — SURVEY.md §8 f4
# Literal restatement, python loops, small meshes only.
# ---------------------------------------------------------------------------
def monolithic_offsets(lengths):
    """assembly.jl:321-333: offsets = blocklasts(dofs) .- map(length, blocks(dofs))"""
    out, s = [], 0
    for n in lengths:
        out.append(s)
        s += n
    return out


def skeleton_faces(cell_nodes, nnodes, D, fc=None, simplex=False):
    """GT.skeleton(mesh) on cartesian_mesh output (quads / hexahedra; simplexified meshes with `fc` = the face complex of
    simplex_face_complex and simplex=True): the (D-1)-faces with two cells around, in face-id order of the complexified mesh
    (domain.jl: skeleton = faces with 2 cells in face_incidence(topo, D-1, D)); cells around in increasing cell id
    (face_incidence(topo, d, D) is the transpose of the cell -> face incidence, filled looping over the cells in order,
    topology.jl:313-334).  Per face: its nodes in the face's own vertex order, the two (cell, local face id, permutation id)
    triples (face_permutation_ids, topology.jl:593-634)."""
    if fc is None:
        fc = face_complex(cell_nodes, nnodes, D)
    d = D - 1
    vertex_node = {v: n + 1 for n, v in enumerate(fc["node_vertex"])}
    around = {}
    for cell, row in enumerate(fc["cell_faces"][d]):
        for lface, face in enumerate(row):
            around.setdefault(face, []).append((cell + 1, lface + 1))
    vperms = _simplex_vertex_permutations(d) if simplex else _vertex_permutations(d)
    lfaces = _simplex_lfaces(D, d) if simplex else _cube_lfaces(D, d)
    out = []
    for face in range(1, len(fc["vertices"][d]) + 1):
        ar = around.get(face, [])
        if len(ar) != 2:
            continue
        fv = fc["vertices"][d][face - 1]
        sides = []
        for cell, lface in ar:
            cv = fc["vertices"][D][cell - 1]
            cvertices = lfaces[lface - 1]
            for pi, P in enumerate(vperms):
                if all(fv[P[c] - 1] == cv[cvertices[c] - 1] for c in range(len(cvertices))):
                    sides.append((cell, lface, pi + 1))
                    break
            else:
                raise AssertionError("Valid pindex not found")
        out.append(dict(face=face, nodes=[vertex_node[v] for v in fv], sides=sides))
    return out


def reference_map_tables(D, point_to_x, simplex=False):
    """accessors.jl:1914-1943 reference_map(refdface, refDface) evaluated at the face quadrature points, for the unit
    D-cube (or D-simplex) and its (D-1)-faces: per local face, per node permutation `ids` of the face,
    φ(x) = Σ_dof coeff[dof]·M_dof(x) with coeff[ids] = node_coordinates(boundary)[lface_nodes]  ->  tables[ldface][perm] = [n_points][D]."""
    d = D - 1
    if simplex:
        Xref = [[0.0] * D] + [[1.0 if m == k else 0.0 for m in range(D)] for k in range(D)]      # v1 = 0, v_{k+2} = e_k
        lfaces, vperms = _simplex_lfaces(D, d), _simplex_vertex_permutations(d)
    else:
        Xref = [[float((v >> m) & 1) for m in range(D)] for v in range(2 ** D)]                 # first index fastest
        lfaces, vperms = _cube_lfaces(D, d), _vertex_permutations(d)
    M, _ = tabulate(d, 1, "P" if simplex else "Q", point_to_x)                                   # shape functions of the reference face
    out = []
    for lnodes in lfaces:
        per_perm = []
        for ids in vperms:
            coeff = [None] * len(lnodes)
            for k, node in enumerate(lnodes):
                coeff[ids[k] - 1] = Xref[node - 1]
            pts = []
            for p in range(len(point_to_x)):
                x = [0.0] * D
                for dof in range(len(coeff)):
                    for c in range(D):
                        x[c] += coeff[dof][c] * M[p, dof]
                pts.append(x)
            per_perm.append(pts)
        out.append(per_perm)
    return out


class _LoopPoint:
    """What the generated integrand sees at one (face, point) for one combination of loop indices: shape functions of the
    fields masked by `ifelse(face_around == the_face_around && field == the_field, sfun, zero)` (compiler.jl:728-739).
    v = argument 1 (loop indices field_1 / face_around_1 / dof_1), u = argument 2."""

    def __init__(self, D, field_ncomp):
        self.D, self.field_ncomp = D, field_ncomp

    def _zero(self, field, gradient):
        """zero(eltype(shape_functions(f, … the_field …))): the zero of THAT field's shape-function type"""
        nc = self.field_ncomp[field]
        if gradient:
            return np.zeros(self.D) if nc == 1 else np.zeros((nc, self.D))
        return 0.0 if nc == 1 else np.zeros(nc)

    def _val(self, which, field, side):
        f, a, val, _ = which
        return val if (f == field and a == side) else self._zero(field, False)

    def _grad(self, which, field, side):
        f, a, _, g = which
        return g if (f == field and a == side) else self._zero(field, True)

    def n(self, side): return self.normals[side - 1]       # unit_normal(mesh, D-1)[side](x)  (accessors.jl:1009-1035)

    def v(self, field, side=1): return self._val(self.V, field, side)
    def u(self, field, side=1): return self._val(self.U, field, side)
    def grad_v(self, field, side=1): return self._grad(self.V, field, side)     # ForwardDiff.gradient / jacobian
    def grad_u(self, field, side=1): return self._grad(self.U, field, side)
    def div_v(self, field, side=1): return float(np.trace(self._grad(self.V, field, side)))
    def div_u(self, field, side=1): return float(np.trace(self._grad(self.U, field, side)))


def _shape(N_a, g_a, comp, n_comp, D):
    """value and gradient of the shape function (node a, component comp): scalar -> (N_a, ∇N_a); vector-valued ->
    (N_a e_comp, e_comp ⊗ ∇N_a) (space.jl:1267-1271; the Jacobian's row `comp` is ∇N_a)"""
    if n_comp == 1:
        return N_a, g_a
    val = np.zeros(n_comp)
    val[comp] = N_a
    if g_a is None:
        return val, None
    jac = np.zeros((n_comp, D))
    jac[comp, :] = g_a
    return val, jac


def map_unit_normal(J, n):
    """accessors.jl:1026-1035, literally: pinvJt = transpose(inv(Jt*J)*Jt); v = pinvJt*n; v / sqrt(v⋅v) (zero below eps)"""
    J = np.asarray(J, dtype=np.float64)
    Jt = J.T
    pinvJt = (np.linalg.inv(Jt @ J) @ Jt).T
    v = pinvJt @ np.asarray(n, dtype=np.float64)
    m = np.sqrt(frobenius(v, v))
    if m < np.finfo(np.float64).eps:
        return np.zeros_like(v)
    return v / m


def face_diameter(coords, nodes):
    """diameter(::MeshFace) (accessors.jl:907-921): the largest distance between two nodes of the face"""
    diam = 0.0
    for i in nodes:
        for j in nodes:
            dx = np.asarray(coords[i - 1]) - np.asarray(coords[j - 1])
            diam = max(diam, float(np.sqrt(frobenius(dx, dx))))
    return diam


def frobenius(a, b):
    """`a ⋅ b` of two SVectors / SMatrices: Σ of the elementwise products in memory (column-major) order"""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.ndim == 0:
        return float(a * b)
    fa, fb = a.reshape(-1, order="F"), b.reshape(-1, order="F")
    s = fa[0] * fb[0]
    for k in range(1, fa.size):
        s = s + fa[k] * fb[k]
    return float(s)


def assemble_matrix_multifield(D, coords, face_nodes, face_tab, sides, fields, integrand, alpha=1.0,
                               free_or_dirichlet=(FREE, FREE), cell_geometry=None, return_coo=False, skeleton_geometry=None):
    """generate_matrix_assembly_template (compiler.jl:1826-1923) + MonolithicAssemblyAllocation (assembly.jl:386-416)
    + contribute!(::MatrixAllocation) (:189-208) + compress (:571-575), loop for loop.

    coords, face_nodes        geometry of the integration faces (cells of the mesh for volume integrals)
    face_tab                  dict(w [nq], dM [nq][nlnf][d]): quadrature and geometry tabulation on the reference face
    sides[face]               list over the cells around of (cell (1-based), tabulation variant (0-based))
    fields                    list of dict(cell_dofs [n_cells][nld] signed, n_free, n_dirichlet, n_comp,
                                           N [n_var][nq][nls], dN [n_var][nq][nls][D] or None)
    integrand(pt)             user integrand on a _LoopPoint (masked shape functions), e.g.
                              lambda p: frobenius(p.grad_v(0), p.grad_u(0)) - p.div_v(0) * p.u(1) + p.v(1) * p.div_u(0)
    cell_geometry             (cell_nodes, dM_cell [nq][nln][D]) for physical gradients (volume integrals only)
    skeleton_geometry         (cell_nodes, dM_cell [n_var][nq][nln][D], ref_normals [n_var][D]): gradients and unit normals of
                              the cells around a skeleton face (pt.grad_u(f, side), pt.n(side)); pt.h = diameter of the face
    -> colptr, rowval, nzval of the monolithic matrix (rows: u's fields, columns: v's fields — SURVEY A.8b)."""
    fr, fcn = free_or_dirichlet
    nf = len(fields)
    n_rows_f = [f["n_free"] if fr == FREE else f["n_dirichlet"] for f in fields]
    n_cols_f = [f["n_free"] if fcn == FREE else f["n_dirichlet"] for f in fields]
    off_i, off_j = monolithic_offsets(n_rows_f), monolithic_offsets(n_cols_f)
    w, dMf = face_tab["w"], face_tab["dM"]
    nq = len(w)
    fn = np.asarray(face_nodes)
    I, J, V = [], [], []
    pt = _LoopPoint(D, [f["n_comp"] for f in fields])
    for face in range(fn.shape[0]):
        n_around = len(sides[face])
        nld = [len(f["cell_dofs"][0]) for f in fields]
        be = {}
        for f1 in range(nf):
            for f2 in range(nf):
                for a1 in range(n_around):
                    for a2 in range(n_around):
                        be[a2, a1, f2, f1] = np.zeros((nld[f2], nld[f1]))
        for q in range(nq):
            Jf = point_geometry(coords, fn[face:face + 1], np.asarray(dMf[q]))
            dV = float(change_of_measure(Jf)[0] * w[q])                          # the FACE's own geometry (accessors.jl:1000-1007)
            # shape functions of every field on every cell around at this point
            sh = {}
            if skeleton_geometry is not None:
                cnS, dMS, nrefS = skeleton_geometry
                pt.h = face_diameter(coords, fn[face])
                pt.normals, JS = [], []
                for (cell, var) in sides[face]:
                    Jc = point_geometry(coords, np.asarray(cnS)[cell - 1:cell], np.asarray(dMS[var][q]))
                    JS.append(Jc)
                    pt.normals.append(map_unit_normal(Jc[0], nrefS[var]))
            for f, fld in enumerate(fields):
                nc = fld["n_comp"]
                for a, (cell, var) in enumerate(sides[face]):
                    g = None
                    if fld.get("dN") is not None and skeleton_geometry is not None:
                        Jt = np.swapaxes(JS[a], -1, -2)
                        g = [_solve(Jt, np.asarray(fld["dN"][var][q][s]).reshape(1, D))[0] for s in range(len(fld["N"][var][q]))]
                    elif fld.get("dN") is not None and cell_geometry is not None:
                        cn, dMc = cell_geometry
                        Jc = point_geometry(coords, np.asarray(cn)[cell - 1:cell], np.asarray(dMc[q]))
                        Jt = np.swapaxes(Jc, -1, -2)
                        g = [_solve(Jt, np.asarray(fld["dN"][var][q][s]).reshape(1, D))[0] for s in range(len(fld["N"][var][q]))]
                    vals = []
                    for ld in range(nld[f]):
                        s, comp = divmod(ld, nc)
                        vals.append(_shape(fld["N"][var][q][s], None if g is None else g[s], comp, nc, D))
                    sh[f, a] = vals
            for f1 in range(nf):
                for f2 in range(nf):
                    for a1 in range(n_around):
                        for a2 in range(n_around):
                            for j in range(nld[f1]):
                                pt.V = (f1, a1 + 1) + sh[f1, a1][j]
                                for i in range(nld[f2]):
                                    pt.U = (f2, a2 + 1) + sh[f2, a2][i]
                                    be[a2, a1, f2, f1][i, j] += (alpha * integrand(pt)) * dV
        for f1 in range(nf):
            for f2 in range(nf):
                for a1 in range(n_around):
                    dofs_j = fields[f1]["cell_dofs"][sides[face][a1][0] - 1]
                    for a2 in range(n_around):
                        dofs_i = fields[f2]["cell_dofs"][sides[face][a2][0] - 1]
                        blk = be[a2, a1, f2, f1]
                        for j, dj in enumerate(dofs_j):
                            if _skip(dj, fcn):
                                continue
                            for i, di in enumerate(dofs_i):
                                if _skip(di, fr):
                                    continue
                                I.append(off_i[f2] + (di if fr == FREE else -di))
                                J.append(off_j[f1] + (dj if fcn == FREE else -dj))
                                V.append(blk[i, j])
    if return_coo:      # the triplets as pushed: a sum of integrals concatenates them before ONE compress (problems.jl:319-350)
        return np.array(I, dtype=np.int64), np.array(J, dtype=np.int64), np.array(V)
    return sparse_csc(np.array(I, dtype=np.int64), np.array(J, dtype=np.int64), np.array(V), sum(n_rows_f), sum(n_cols_f))


def assemble_matrix_sum(coo_list, m, n):
    """assemble_matrix over a sum of integrals (problems.jl:319-350): every contribution pushes its triplets into the same COO
    allocation, in the order of the sum; one compress."""
    I = np.concatenate([np.asarray(c[0], dtype=np.int64) for c in coo_list])
    J = np.concatenate([np.asarray(c[1], dtype=np.int64) for c in coo_list])
    V = np.concatenate([np.asarray(c[2], dtype=np.float64) for c in coo_list])
    return sparse_csc(I, J, V, m, n)


def assemble_vector_multifield(D, coords, face_nodes, face_tab, sides, fields, integrand, alpha=1.0, free_or_dirichlet=FREE,
                               skeleton_geometry=None, point_data=None):
    """generate_vector_assembly_template (compiler.jl:1933-2000) + monolithic contribute! (assembly.jl:392-399), loop for
    loop; integrand(pt) sees the masked test function v only.  skeleton_geometry as in assemble_matrix_multifield (gradients of
    v, pt.n(side), pt.h); point_data [n_faces][nq]: an analytical field sampled at the face points (pt.g)."""
    nf = len(fields)
    n_rows_f = [f["n_free"] if free_or_dirichlet == FREE else f["n_dirichlet"] for f in fields]
    off = monolithic_offsets(n_rows_f)
    w, dMf = face_tab["w"], face_tab["dM"]
    fn = np.asarray(face_nodes)
    I, V = [], []
    pt = _LoopPoint(D, [f["n_comp"] for f in fields])
    for face in range(fn.shape[0]):
        n_around = len(sides[face])
        nld = [len(f["cell_dofs"][0]) for f in fields]
        be = {(a, f): np.zeros(nld[f]) for f in range(nf) for a in range(n_around)}
        for q in range(len(w)):
            Jf = point_geometry(coords, fn[face:face + 1], np.asarray(dMf[q]))
            dV = float(change_of_measure(Jf)[0] * w[q])
            if point_data is not None:
                pt.g = float(point_data[face][q])
            JS = None
            if skeleton_geometry is not None:
                cnS, dMS, nrefS = skeleton_geometry
                pt.h = face_diameter(coords, fn[face])
                pt.normals, JS = [], []
                for (cell, var) in sides[face]:
                    Jc = point_geometry(coords, np.asarray(cnS)[cell - 1:cell], np.asarray(dMS[var][q]))
                    JS.append(Jc)
                    pt.normals.append(map_unit_normal(Jc[0], nrefS[var]))
            for f in range(nf):
                nc = fields[f]["n_comp"]
                for a in range(n_around):
                    var = sides[face][a][1]
                    g = None
                    if JS is not None and fields[f].get("dN") is not None:
                        Jt = np.swapaxes(JS[a], -1, -2)
                        g = [_solve(Jt, np.asarray(fields[f]["dN"][var][q][s]).reshape(1, D))[0] for s in range(len(fields[f]["N"][var][q]))]
                    for ld in range(nld[f]):
                        s, comp = divmod(ld, nc)
                        pt.V = (f, a + 1) + _shape(fields[f]["N"][var][q][s], None if g is None else g[s], comp, nc, D)
                        be[a, f][ld] += (alpha * integrand(pt)) * dV
        for f in range(nf):
            for a in range(n_around):
                dofs = fields[f]["cell_dofs"][sides[face][a][0] - 1]
                for i, di in enumerate(dofs):
                    if _skip(di, free_or_dirichlet):
                        continue
                    I.append(off[f] + (di if free_or_dirichlet == FREE else -di))
                    V.append(be[a, f][i])
    return dense_vector(np.array(I, dtype=np.int64), np.array(V), sum(n_rows_f))
