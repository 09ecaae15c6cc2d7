"""Host-side input preparation for the assembly engine (vectorised numpy).

The engine's C ABI takes flat arrays: node coordinates,
cell->node and cell->dof maps, quadrature weights and tabulated shape
functions.  In production the Julia host computes them with GalerkinToolkit
itself; this module is the Python stand-in that produces *the same arrays the
reference would produce* for the BASELINE.json workloads, at full size, fast.

Everything here is input preparation, not the hot path.  The literal,
loop-by-loop restatement used to check these closed forms lives in
``oracle/gt_oracle.py`` (test infrastructure; never imported from here).

Reference citations are relative to /root/reference/src.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

FREE = 1        # field.jl:136-142 (FREE / DIRICHLET enums)
DIRICHLET = 2


# ----------------------------------------------------------------------------
# Mesh  (cartesian_mesh.jl:16-54, 213-263, 265-328)
# ----------------------------------------------------------------------------
@dataclass
class Mesh:
    D: int
    node_coordinates: np.ndarray          # [n_nodes, D] float64 (= Vector{SVector{D,Float64}})
    cell_nodes: np.ndarray                # [n_cells, n_lnodes] int32, 1-based
    cells_per_dir: tuple
    simplex: bool
    domain: tuple

    @property
    def n_nodes(self) -> int:
        return self.node_coordinates.shape[0]

    @property
    def n_cells(self) -> int:
        return self.cell_nodes.shape[0]

    @property
    def n_lnodes(self) -> int:
        return self.cell_nodes.shape[1]


# local hex nodes -> simplices, domain.jl:322-336 (simplex_nodes)
_SIMPLEX_NODES = {
    1: np.array([[1, 2]]),
    2: np.array([[1, 2, 3], [4, 3, 2]]),
    3: np.array([[1, 2, 3, 7], [1, 2, 5, 7], [2, 3, 4, 7],
                 [2, 4, 7, 8], [2, 5, 6, 7], [2, 6, 7, 8]]),
}


def cartesian_mesh(domain: Sequence[float], cells_per_dir: Sequence[int],
                   simplexify: bool = False, z_cell_range: Optional[tuple] = None) -> Mesh:
    """GT.cartesian_mesh(domain, cells_per_dir; simplexify) — nodes, coordinates, cells.

    Node id (1-based) = 1 + i + (n1+1) j + (n1+1)(n2+1) k, coordinates
    ``pmin + h*(i,j,k)`` with ``h = (pmax-pmin)/cells`` (cartesian_mesh.jl:213-247);
    cells x-fastest with local nodes in tensor order (:229-241); simplices
    6*(hex-1)+s with the local-node table of domain.jl:331-332 (:265-328).

    ``z_cell_range=(k0,k1)`` (0-based, half open) returns only the cells of that
    slab of the *last* direction, keeping global node ids — used by the
    multi-GPU partition; nodes/coordinates stay global.
    """
    cells = tuple(int(c) for c in cells_per_dir)
    D = len(cells)
    assert len(domain) == 2 * D
    pmin = np.array([domain[2 * d] for d in range(D)], dtype=np.float64)
    pmax = np.array([domain[2 * d + 1] for d in range(D)], dtype=np.float64)
    h = (pmax - pmin) / np.array(cells, dtype=np.float64)
    nodes_per_dir = tuple(c + 1 for c in cells)
    # node coordinates, first index fastest
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in nodes_per_dir], indexing="ij")
    coords = np.empty((int(np.prod(nodes_per_dir)), D), dtype=np.float64)
    for d in range(D):
        coords[:, d] = (pmin[d] + h[d] * grids[d]).reshape(-1, order="F")
    # cells
    lo = [0] * D
    hi = list(cells)
    if z_cell_range is not None:
        lo[D - 1], hi[D - 1] = z_cell_range
    cidx = np.meshgrid(*[np.arange(lo[d], hi[d], dtype=np.int64) for d in range(D)], indexing="ij")
    cidx = [c.reshape(-1, order="F") for c in cidx]
    strides = np.cumprod((1,) + nodes_per_dir[:-1]).astype(np.int64)
    base = sum(cidx[d] * strides[d] for d in range(D))           # 0-based node of local node 1
    hexn = np.empty((base.shape[0], 2 ** D), dtype=np.int64)
    for ln in range(2 ** D):
        off = sum(((ln >> d) & 1) * strides[d] for d in range(D))
        hexn[:, ln] = base + off + 1
    if simplexify:
        tab = _SIMPLEX_NODES[D] - 1
        cn = hexn[:, tab]                                         # [nhex, ns, D+1]
        cn = cn.reshape(-1, D + 1)
    else:
        cn = hexn
    return Mesh(D, coords, np.ascontiguousarray(cn, dtype=np.int32), cells, bool(simplexify), tuple(domain))


def node_to_vertex(mesh: Mesh) -> np.ndarray:
    """Vertex id of every node (1-based), topology.jl:1034-1097.

    The 0-faces that pre-exist in ``cartesian_mesh`` are the 2^D box corners
    (cartesian_mesh.jl:117-135, nmax = 2^0), found in cell order, i.e. in
    lexicographic order; they get vertex ids 1..2^D (:1072-1078).  Every other
    node follows in node order (:1079-1084).
    """
    D = mesh.D
    npd = tuple(c + 1 for c in mesh.cells_per_dir)
    strides = np.cumprod((1,) + npd[:-1]).astype(np.int64)
    if all(c == 1 for c in mesh.cells_per_dir) and not mesh.simplex:
        return np.arange(1, mesh.n_nodes + 1, dtype=np.int32)
    corners = []
    for ln in range(2 ** D):
        corners.append(int(sum(((ln >> d) & 1) * (npd[d] - 1) * strides[d] for d in range(D))))
    is_corner = np.zeros(mesh.n_nodes, dtype=bool)
    is_corner[corners] = True
    vert = np.empty(mesh.n_nodes, dtype=np.int32)
    vert[corners] = np.arange(1, 2 ** D + 1, dtype=np.int32)      # corners list is lexicographic
    rest = np.flatnonzero(~is_corner)
    vert[rest] = np.arange(2 ** D + 1, mesh.n_nodes + 1, dtype=np.int32)
    return vert


def boundary_node_mask(mesh: Mesh, sides: Optional[Sequence[int]] = None) -> np.ndarray:
    """Nodes lying on the selected sides of the box.  ``sides`` are the 1-based
    local (D-1)-face ids of the reference cube (domain.jl:224, 252-255):
    2D: 1:y=0 2:y=1 3:x=0 4:x=1 ; 3D: 1:z=0 2:z=1 3:y=0 4:y=1 5:x=0 6:x=1.
    ``None`` = the whole boundary (group "boundary", cartesian_mesh.jl:176-178)."""
    D = mesh.D
    npd = tuple(c + 1 for c in mesh.cells_per_dir)
    idx = np.unravel_index(np.arange(mesh.n_nodes), npd, order="F")
    if sides is None:
        sides = range(1, 2 * D + 1)
    mask = np.zeros(mesh.n_nodes, dtype=bool)
    for s in sides:
        axis = D - 1 - (s - 1) // 2
        upper = (s - 1) % 2 == 1
        mask |= idx[axis] == (npd[axis] - 1 if upper else 0)
    return mask


def boundary_faces(mesh: Mesh, sides: Optional[Sequence[int]] = None):
    """The (D-1)-faces `cartesian_mesh` creates on the boundary (cartesian_mesh.jl:117-168; simplexified meshes :344-409,
    sub-faces from simplexify(::UnitNCube) domain.jl:270-320): hex-cell-major, cube-local faces in increasing order,
    simplex sub-faces in increasing id; a cube-local face is a boundary face iff each of its nodes belongs to at most
    2^(D-1) hex cells; it goes to group "<D-1>-face-<ldface>".  `sides` selects groups (None = all, group "boundary");
    faces of a domain come in increasing face id (domain.jl:705-753).
    -> face_nodes [nf, n_face_nodes] int32 1-based (the face's own local order), face_cell [nf] 0-based HEX cell,
       face_ldface [nf] 1-based cube-local face."""
    from . import refnumbering
    nodes, group, cell = refnumbering.boundary_face_nodes(mesh, mesh.D - 1)
    if sides is not None:
        keep = np.isin(group, np.asarray(list(sides), dtype=np.int64))
        nodes, group, cell = nodes[keep], group[keep], cell[keep]
    return np.ascontiguousarray(nodes + 1, dtype=np.int32), cell.astype(np.int64), group.astype(np.int32)


def lattice_dof_map(space: "LagrangeSpace"):
    """signed dof id of every point of the order-times refined node lattice (x fastest) and component:
    -> (lat2dof [n_lattice_points, n_comp] int32, strides [D], k)"""
    mesh, k, nc = space.mesh, space.order, space.n_comp
    D = mesh.D
    npd = np.array([c + 1 for c in mesh.cells_per_dir], dtype=np.int64)
    strides_n = np.cumprod(np.concatenate(([1], npd[:-1])))
    cnodes = mesh.cell_nodes.astype(np.int64) - 1
    vidx = np.stack([(cnodes // strides_n[d]) % npd[d] for d in range(D)], axis=2)      # [nc, n_lnodes, D]
    lat = monomial_exponents(D, k, space.kind)                                           # [nls, D] local lattice
    if mesh.simplex:
        glat = k * vidx[:, None, 0, :] + np.einsum("lm,cmd->cld", lat, vidx[:, 1:, :] - vidx[:, :1, :])
    else:
        glat = k * vidx[:, None, 0, :] + lat[None, :, :]
    ext = k * (npd - 1) + 1
    strides = np.cumprod(np.concatenate(([1], ext[:-1])))
    lin = (glat * strides[None, None, :]).sum(axis=2)                                    # [nc, nls]
    lat2dof = np.zeros((int(np.prod(ext)), nc), dtype=np.int32)
    lat2dof[lin.reshape(-1)] = space.cell_dofs.reshape(-1, nc)
    return lat2dof, strides, k


@dataclass
class FaceProblem:
    """A boundary domain handed to the engine as a mesh of (D-1)-cells embedded in D dimensions."""
    face_nodes: np.ndarray     # [nf, n_face_nodes] int32 1-based mesh nodes
    face_dofs: np.ndarray      # [nf, n_lfdofs] int32 signed dofs of the space on each face
    face_cell: np.ndarray
    face_ldface: np.ndarray
    tab: "Tabulation"          # tabulation on the reference face (gradients with D-1 components)


def face_problem(space: "LagrangeSpace", sides: Optional[Sequence[int]], degree: int) -> FaceProblem:
    """Inputs of a boundary integral ∫_Γ g v dΓ (Neumann term): measure(Γ, degree) + the space's dofs on Γ's faces.
    The Lagrange functions of a cell restricted to one of its faces are the face's own Lagrange functions (all others
    vanish there), so a face's dofs are the space's dofs at the lattice points of the face, in the face's own node order
    (its vertices as `face_nodes` lists them; Q1 / barycentric map of the reference face)."""
    mesh = space.mesh
    D, k = mesh.D, space.order
    fn, fc, lf = boundary_faces(mesh, sides)
    lat2dof, strides, _ = lattice_dof_map(space)
    npd = np.array([c + 1 for c in mesh.cells_per_dir], dtype=np.int64)
    strides_n = np.cumprod(np.concatenate(([1], npd[:-1])))
    f0 = fn.astype(np.int64) - 1
    vidx = np.stack([(f0 // strides_n[d]) % npd[d] for d in range(D)], axis=2)           # [nf, n_face_nodes, D]
    latf = monomial_exponents(D - 1, k, space.kind)                                       # [nlf, D-1]
    if mesh.simplex:
        g = k * vidx[:, None, 0, :] + np.einsum("lm,fmd->fld", latf, vidx[:, 1:, :] - vidx[:, :1, :])
    else:       # face vertices in tensor order: v0, v0 + e_a, v0 + e_b, ... -> axes from vertices 2^m
        axes = np.stack([vidx[:, 2 ** m, :] - vidx[:, 0, :] for m in range(D - 1)], axis=1)   # [nf, D-1, D]
        g = k * vidx[:, None, 0, :] + np.einsum("lm,fmd->fld", latf, axes)
    lin = (g * strides[None, None, :]).sum(axis=2)                                        # [nf, nlf]
    fd = lat2dof[lin].reshape(fn.shape[0], -1)                                            # node-major, component-minor
    if (fd == 0).any():
        raise AssertionError("a boundary-face lattice point carries no dof")
    q = quadrature(D - 1, mesh.simplex, degree)
    N, dN = tabulate(D - 1, k, space.kind, q.coordinates)
    M, dM = tabulate(D - 1, 1, space.kind, q.coordinates)
    return FaceProblem(fn, np.ascontiguousarray(fd, dtype=np.int32), fc, lf,
                       Tabulation(np.ascontiguousarray(q.weights), N, dN, M, dM, q.coordinates))


# ----------------------------------------------------------------------------
# Reference elements, quadrature, tabulation
# ----------------------------------------------------------------------------
def monomial_exponents(D: int, order: int, kind: str) -> np.ndarray:
    """space.jl:1127-1145: all of {0..k}^D, first index fastest; P keeps sum<=k."""
    rng = [np.arange(order + 1)] * D
    g = np.meshgrid(*rng, indexing="ij")
    e = np.stack([x.reshape(-1, order="F") for x in g], axis=1) if D > 0 else np.zeros((1, 0), dtype=int)
    if kind == "P":
        e = e[e.sum(axis=1) <= order]
    return e.astype(np.int64)


def reference_nodes(D: int, order: int, kind: str) -> np.ndarray:
    """space.jl:1149-1177: node = exponent / order."""
    e = monomial_exponents(D, order, kind)
    return e.astype(np.float64) / order if order != 0 else e.astype(np.float64)


def _monomials(e: np.ndarray, x: np.ndarray):
    """values m_j(x_p) [p,j] and gradients [p,j,d] of x->prod(x.^e_j)
    (space.jl:1204-1211; derivative = what ForwardDiff returns, exact)."""
    P, D = x.shape
    nj = e.shape[0]
    pw = np.ones((P, nj, D))
    dpw = np.zeros((P, nj, D))
    for d in range(D):
        for j in range(nj):
            k = e[j, d]
            pw[:, j, d] = x[:, d] ** k
            dpw[:, j, d] = k * x[:, d] ** (k - 1) if k > 0 else 0.0
    val = np.prod(pw, axis=2)
    grad = np.empty((P, nj, D))
    for d in range(D):
        others = np.ones((P, nj))
        for dd in range(D):
            others = others * (dpw[:, :, dd] if dd == d else pw[:, :, dd])
        grad[:, :, d] = others
    return val, grad


def tabulate(D: int, order: int, kind: str, points: np.ndarray):
    """tabulator(fe)(f, x) = C*B with A[i,j] = m_j(x_i), B = A\\I (space.jl:960-970).
    Returns N[p,dof] and dN[p,dof,d]; the reference stores the transpose
    [dof,point] (accessors.jl:486-496), which in Julia column-major memory is the
    same byte order as this row-major [point][dof] array."""
    e = monomial_exponents(D, order, kind)
    nodes = reference_nodes(D, order, kind)
    A, _ = _monomials(e, nodes)
    B = np.linalg.solve(A, np.eye(A.shape[0]))
    C, Cg = _monomials(e, np.asarray(points, dtype=np.float64).reshape(-1, D))
    N = C @ B
    dN = np.einsum("pjd,ji->pid", Cg, B)
    return np.ascontiguousarray(N), np.ascontiguousarray(dN)


# Strang tet rules used by the reference for D==3 (quadrature.jl:500-635).  Only
# the degree the BASELINE workloads use are tabulated here as *inputs*; the
# numbers are the published Strang-Fix / Keast constants.
def _tet_perm4(a, b):
    return [(a, b, b), (b, a, b), (b, b, a), (b, b, b)]


def gauss_legendre_01(n: int):
    """quadrature.jl:60-66 with limits (0,1): x = 0.5x+0.5, w *= 0.5."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * x + 0.5, 0.5 * w


@dataclass
class Quadrature:
    coordinates: np.ndarray   # [nq, D]
    weights: np.ndarray       # [nq]


def quadrature(D: int, simplex: bool, degree: int) -> Quadrature:
    """quadrature.jl:36-52.  n-cube: tensor Gauss-Legendre, n = ceil((d+1)/2)
    per direction, first index fastest, weight = prod (:78-106).  Simplex:
    Duffy (:124-157) for triangles always and for tets of degree > 5; Strang-Fix tables for tets of degree 1-5
    (:41-48, 500-635)."""
    if not simplex:
        n = int(np.ceil((degree + 1) / 2))
        x1, w1 = gauss_legendre_01(n)
        g = np.meshgrid(*([np.arange(n)] * D), indexing="ij")
        idx = [a.reshape(-1, order="F") for a in g]
        x = np.stack([x1[i] for i in idx], axis=1)
        w = np.ones(n ** D)
        for i in idx:
            w = w * w1[i]
        return Quadrature(np.ascontiguousarray(x), w)
    from . import _simplex_rules
    if D == 3 and degree in _simplex_rules.STRANG_TET_DEGREES:      # quadrature.jl:41-48: Strang only in 3D, else Duffy
        return Quadrature(*_simplex_rules.strang_tet_quadrature(degree))
    return Quadrature(*_simplex_rules.simplex_quadrature(D, degree))


# ----------------------------------------------------------------------------
# Spaces (space.jl:299-535, 1702-1735)
# ----------------------------------------------------------------------------
@dataclass
class LagrangeSpace:
    mesh: Mesh
    order: int
    n_comp: int
    kind: str                          # "Q" | "P"
    cell_dofs: np.ndarray              # [n_cells, n_ldofs] int32, 1-based, <0 = Dirichlet id
    n_free: int
    n_dirichlet: int
    # for interpolation / tests: coordinates of the scalar node behind every free / dirichlet dof
    free_dof_nodes: Optional[np.ndarray] = None
    dirichlet_dof_nodes: Optional[np.ndarray] = None

    @property
    def n_ldofs(self) -> int:
        return self.cell_dofs.shape[1]


def _apply_dirichlet(cell_dofs_all: np.ndarray, ndofs: int, tag: np.ndarray):
    """space.jl:512-524 + partition_from_mask :910-920: free dofs renumbered
    1..nfree in increasing old id; Dirichlet ones become -(1..ndiri)."""
    free = np.flatnonzero(~tag)
    diri = np.flatnonzero(tag)
    perm = np.empty(ndofs, dtype=np.int64)
    perm[free] = np.arange(1, free.size + 1)
    perm[diri] = -np.arange(1, diri.size + 1)
    return perm[cell_dofs_all - 1].astype(np.int32), int(free.size), int(diri.size), free, diri


def discontinuous_lagrange_space(mesh: Mesh, order: int = 1, n_comp: int = 1) -> LagrangeSpace:
    """GT.lagrange_space(Ω, order; continuous=false): every dof is an own dof of its cell (reference_face_own_dofs,
    space.jl:860-882: all local dofs belong to the D-face), so dof = (cell-1) * n_ldofs + local dof; no Dirichlet dofs
    (boundary conditions are imposed weakly, docs/src/src_jl/example_hello_world_dg.jl)."""
    kind = "P" if mesh.simplex else "Q"
    lat = reference_nodes(mesh.D, order, kind)                                  # [nls, D] reference nodes
    nls = lat.shape[0]
    nld = nls * n_comp
    cell_dofs = (np.arange(mesh.n_cells * nld, dtype=np.int64).reshape(mesh.n_cells, nld) + 1).astype(np.int32)
    M, _ = tabulate(mesh.D, 1, kind, lat)                                       # geometry functions at the dof nodes
    X = np.einsum("sk,ckd->csd", M, mesh.node_coordinates[mesh.cell_nodes.astype(np.int64) - 1])   # [nc, nls, D]
    xyz = np.repeat(X.reshape(-1, mesh.D), n_comp, axis=0)
    return LagrangeSpace(mesh, order, n_comp, kind, np.ascontiguousarray(cell_dofs), int(mesh.n_cells * nld), 0,
                         free_dof_nodes=xyz, dirichlet_dof_nodes=np.zeros((0, mesh.D)))


def lagrange_space(mesh: Mesh, order: int = 1, dirichlet_boundary=None,
                   n_comp: int = 1, node_dof_override=None) -> LagrangeSpace:
    """GT.lagrange_space(Ω, order; dirichlet_boundary, tensor_size=Val((n_comp,))).

    order 1: dof = vertex id (space.jl:327-417 with one own dof per 0-face),
    vertex ids from :func:`node_to_vertex`.  Order >= 2 (quad / hex and simplexified
    meshes): the reference's face-complex numbering, :mod:`refnumbering` (checked
    against the oracle's loop-for-loop restatement).
    ``dirichlet_boundary``: None | "boundary" | list of box-side ids.
    Vector-valued: dof = (node-1)*n_comp + c (space.jl:1267-1271, 1506-1510).
    """
    kind = "P" if mesh.simplex else "Q"
    if order >= 2 and node_dof_override is None:
        # the reference's own numbering (face complex + dimension-major offsets + face permutations), see refnumbering.py
        from . import refnumbering
        cell_dofs, nfree, ndiri, xf, xd = refnumbering.scalar_or_vector_dofs(mesh, order, n_comp, dirichlet_boundary)
        return LagrangeSpace(mesh, order, n_comp, kind, cell_dofs, nfree, ndiri, free_dof_nodes=xf, dirichlet_dof_nodes=xd)
    if order == 1:
        vert = node_to_vertex(mesh).astype(np.int64)
        scal = vert[mesh.cell_nodes.astype(np.int64) - 1]          # [nc, nln] scalar dof ids
        n_scal = mesh.n_nodes
        # scalar dof -> node (for tagging and for tests)
        dof_node = np.empty(n_scal, dtype=np.int64)
        dof_node[vert - 1] = np.arange(mesh.n_nodes)
        if dirichlet_boundary is None:
            tag_scal = np.zeros(n_scal, dtype=bool)
        else:
            sides = None if dirichlet_boundary == "boundary" else dirichlet_boundary
            tag_scal = boundary_node_mask(mesh, sides)[dof_node]
        dof_xyz = mesh.node_coordinates[dof_node]
    else:
        raise ValueError("order must be >= 1 (order >= 2 is numbered by refnumbering.py above)")
    if n_comp == 1:
        all_dofs, ndofs, tag = scal, n_scal, tag_scal
        dof_scalar = np.arange(n_scal)
    else:
        c = np.arange(n_comp, dtype=np.int64)
        all_dofs = ((scal[:, :, None] - 1) * n_comp + c[None, None, :] + 1).reshape(scal.shape[0], -1)
        ndofs = n_scal * n_comp
        tag = np.repeat(tag_scal, n_comp)
        dof_scalar = np.repeat(np.arange(n_scal), n_comp)
    cell_dofs, nfree, ndiri, free, diri = _apply_dirichlet(all_dofs, ndofs, tag)
    return LagrangeSpace(mesh, order, n_comp, kind, np.ascontiguousarray(cell_dofs), nfree, ndiri,
                         free_dof_nodes=dof_xyz[dof_scalar[free]], dirichlet_dof_nodes=dof_xyz[dof_scalar[diri]])


@dataclass
class Tabulation:
    """What gtk_set_tabulation takes."""
    w: np.ndarray      # [nq]
    N: np.ndarray      # [nq, n_lscalar]     space shape values
    dN: np.ndarray     # [nq, n_lscalar, D]  space reference gradients
    M: np.ndarray      # [nq, n_lnodes]      geometry shape values
    dM: np.ndarray     # [nq, n_lnodes, D]
    xq: np.ndarray     # [nq, D] reference points (host-side only)


def measure_tabulation(space: LagrangeSpace, degree: int) -> Tabulation:
    """GT.measure(Ω, degree) + space_face(V, dΩ; tabulate=(value, ∇)) (problems.jl:12-25,
    accessors.jl:1135-1191, 486-496).  Geometry is the order-1 Lagrange map of the mesh."""
    mesh = space.mesh
    q = quadrature(mesh.D, mesh.simplex, degree)
    N, dN = tabulate(mesh.D, space.order, space.kind, q.coordinates)
    M, dM = tabulate(mesh.D, 1, space.kind, q.coordinates)
    return Tabulation(np.ascontiguousarray(q.weights), N, dN, M, dM, q.coordinates)
