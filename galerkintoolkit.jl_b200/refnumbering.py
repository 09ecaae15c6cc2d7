"""The reference's dof numbering of order-k Lagrange spaces on quad / hex `cartesian_mesh`es, vectorised (host-side
input preparation; `cell_dofs` is an input of the C ABI).

What the reference does:
  * `complexify` builds the face complex: vertex ids (topology.jl:1034-1097); then, highest dimension first, the d-faces
    from the (d+1)-faces (generate_face_boundary, topology.jl:1594-1704): the boundary d-faces `cartesian_mesh` created
    (cartesian_mesh.jl:117-168) keep ids 1.., every other face gets the next id when it is FIRST met looping over the
    (d+1)-faces in id order and their local faces in reference order; its vertex list is the one seen from that first
    parent (generate_face_vertices, topology.jl:1468-1540);
  * `generate_dof_ids` (space.jl:299-535) numbers dofs dimension-major, then by face id; the own dofs of a face are placed
    in a cell through the permutation that maps the face's vertex order to the cell's (topology.jl:593-666,
    space.jl:1439-1487); Dirichlet dofs = all dofs of the cell-local (D-1)-faces lying in Γ, stable partition (:477-535).

Here "first met" is computed with `np.unique(..., return_index=True)` over all (parent, local face) candidates in loop
order instead of a loop with a dictionary — a different implementation of the same rule; `tests/test_host_side.py` checks
it against the loop-for-loop restatement in the oracle on small meshes, orders 1-4, 2D and 3D.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import hostprep as _hp

# local faces of the unit n-cube as 0-based local vertices (domain.jl:188-255)
_LFACES = {
    (1, 0): [[0], [1]],
    (2, 0): [[0], [1], [2], [3]], (2, 1): [[0, 1], [2, 3], [0, 2], [1, 3]],
    (3, 0): [[i] for i in range(8)],
    (3, 1): [[0, 1], [2, 3], [0, 2], [1, 3], [4, 5], [6, 7], [4, 6], [5, 7], [0, 4], [2, 6], [1, 5], [3, 7]],
    (3, 2): [[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 4, 5], [2, 3, 6, 7], [0, 2, 4, 6], [1, 3, 5, 7]],
}
# admissible vertex permutations (domain.jl:53-100): both orders of a segment; the 8 symmetries of the square in the
# lexicographic order of Combinatorics.permutations; identity only for d = 0 and d > 2
_VPERMS = {0: [[0]], 1: [[0, 1], [1, 0]],
           2: [[0, 1, 2, 3], [0, 2, 1, 3], [1, 0, 3, 2], [1, 3, 0, 2], [2, 0, 3, 1], [2, 3, 0, 1], [3, 1, 2, 0], [3, 2, 1, 0]],
           3: [list(range(8))]}


# the same for the unit n-simplex (domain.jl:389-470) — every permutation of a simplex is admissible
_SLFACES = {
    (1, 0): [[0], [1]],
    (2, 0): [[0], [1], [2]], (2, 1): [[0, 1], [0, 2], [1, 2]],
    (3, 0): [[0], [1], [2], [3]], (3, 1): [[0, 1], [0, 2], [1, 2], [0, 3], [1, 3], [2, 3]],
    (3, 2): [[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]],
}
_SVPERMS = {0: [[0]], 1: [[0, 1], [1, 0]], 2: [list(p) for p in itertools.permutations(range(3))], 3: [list(range(4))]}
# simplices of the unit cube, 0-based cube-local nodes (domain.jl:322-336)
_CUBE_SIMPLICES = {2: [[0, 1, 2], [3, 2, 1]],
                   3: [[0, 1, 2, 6], [0, 1, 4, 6], [1, 2, 3, 6], [1, 3, 6, 7], [1, 4, 5, 6], [1, 5, 6, 7]]}


def _lfaces(n, d, simplex=False):
    if simplex:
        return np.array([list(range(n + 1))] if d == n else _SLFACES[(n, d)], dtype=np.int64)
    return np.array([list(range(2 ** n))] if d == n else _LFACES[(n, d)], dtype=np.int64)


def _simplexified_cube_subfaces(D, d):
    """simplexify(::UnitNCube) (domain.jl:270-320): d-faces of the complexified 2 / 6 simplices of the cube (no pre-existing
    faces: first-encounter ids, vertex = node), grouped by the cube-local d-face they lie in, ascending id.
    -> per cube-local d-face: array [n_sub, d+1] of 0-based cube-local nodes"""
    verts = {D: np.array(_CUBE_SIMPLICES[D], dtype=np.int64)}
    for dd in range(D - 1, d - 1, -1):
        if dd == 0:
            verts[0] = np.arange(2 ** D, dtype=np.int64)[:, None]
            break
        cand = verts[dd + 1][:, _lfaces(dd + 1, dd, True)].reshape(-1, dd + 1)
        _, first = _first_occurrence_ids(np.sort(cand, axis=1))
        verts[dd] = cand[first]
    out = []
    for cnodes in _lfaces(D, d):
        inside = np.isin(verts[d], cnodes).all(axis=1)
        out.append(verts[d][inside])
    return out


def _first_occurrence_ids(rows_sorted: np.ndarray):
    """ids 1.. in order of first occurrence of each distinct row; -> (id per row, first row index of every id)"""
    _, first, inv = np.unique(rows_sorted, axis=0, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")                 # unique keys by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv.reshape(-1)] + 1, first[order]


def boundary_face_nodes(mesh: _hp.Mesh, d: int):
    """The d-faces `cartesian_mesh` puts on the boundary (cartesian_mesh.jl:117-168; simplexified: :344-409), in face-id
    order: cell-major over the HEX cells, local cube face ascending, (simplices: sub-face id ascending).
    -> (nodes [nf, n_face_nodes] 0-based mesh nodes in the face's own local order, group [nf] = 1-based cube-local face,
        hex cell [nf] 0-based)"""
    D = mesh.D
    sx = bool(mesh.simplex)
    if any(c != 1 for c in mesh.cells_per_dir) and any(c < 2 for c in mesh.cells_per_dir):
        raise ValueError("At least 2 cells in any direction (or 1 cell in all directions)")   # cartesian_mesh.jl:98-100, 331-333
    cn = mesh.cell_nodes.astype(np.int64) - 1
    hexn = (_hp.cartesian_mesh(mesh.domain, mesh.cells_per_dir).cell_nodes.astype(np.int64) - 1) if sx else cn
    node_to_n = np.bincount(hexn.reshape(-1), minlength=mesh.n_nodes)
    tabD = _lfaces(D, d)
    isb = (node_to_n[hexn[:, tabD]] <= 2 ** d).all(axis=2)
    pc, pl = np.nonzero(isb)                                                        # cell-major, local face ascending
    if sx:
        sub = _simplexified_cube_subfaces(D, d)
        nsub = sub[0].shape[0]
        assert all(x.shape[0] == nsub for x in sub)
        subtab = np.stack(sub)                                                       # [n_lfaces, nsub, d+1]
        nodes = hexn[pc[:, None, None], subtab[pl]].reshape(-1, d + 1)
        return nodes, np.repeat(pl, nsub) + 1, np.repeat(pc, nsub)
    return hexn[pc[:, None], tabD[pl]], pl + 1, pc


def face_complex(mesh: _hp.Mesh):
    """vertex lists per face dimension, cell -> faces incidence, number / group of the pre-existing boundary faces"""
    D = mesh.D
    sx = bool(mesh.simplex)
    if any(c != 1 for c in mesh.cells_per_dir) and any(c < 2 for c in mesh.cells_per_dir):
        raise ValueError("At least 2 cells in any direction (or 1 cell in all directions)")   # cartesian_mesh.jl:98-100, 331-333
    cn = mesh.cell_nodes.astype(np.int64) - 1
    # boundary detection always looks at the HEX chain (cartesian_mesh.jl:117-168; for simplexified meshes :344-409)
    hexn = (_hp.cartesian_mesh(mesh.domain, mesh.cells_per_dir).cell_nodes.astype(np.int64) - 1) if sx else cn
    vert = _hp.node_to_vertex(mesh).astype(np.int64)
    node_to_n = np.bincount(hexn.reshape(-1), minlength=mesh.n_nodes)
    verts = {D: vert[cn]}
    inc, n_parent, group = {}, {}, {}
    for d in range(D - 1, 0, -1):
        n = d + 1
        pnodes, pgroup, _ = boundary_face_nodes(mesh, d)
        pv, pl = vert[pnodes], pgroup - 1                                            # parents, own vertex order
        tab = _lfaces(n, d, sx)
        cand = verts[n][:, tab].reshape(-1, tab.shape[1])                           # (parent, local face) in loop order
        rows = np.vstack([pv, cand])
        ids, first = _first_occurrence_ids(np.sort(rows, axis=1))
        verts[d] = rows[first]                                                      # vertex order of the first incident parent
        inc[(n, d)] = ids[pv.shape[0]:].reshape(-1, tab.shape[0])
        n_parent[d], group[d] = pv.shape[0], pl + 1
        if not np.array_equal(ids[: pv.shape[0]], np.arange(1, pv.shape[0] + 1)):
            raise ValueError("boundary faces of the parent mesh are not pairwise distinct")
    nv = int(vert.max())
    verts[0] = np.arange(1, nv + 1, dtype=np.int64)[:, None]
    cell_faces = {D: np.arange(1, cn.shape[0] + 1, dtype=np.int64)[:, None], 0: verts[D]}
    for d in range(1, D):
        if (D, d) in inc:
            cell_faces[d] = inc[(D, d)]
        else:                                                                       # non-adjacent dimensions: match vertex sets
            tab = _lfaces(D, d, sx)
            cand = np.sort(verts[D][:, tab].reshape(-1, tab.shape[1]), axis=1)
            rows = np.vstack([np.sort(verts[d], axis=1), cand])
            _, inv = np.unique(rows, axis=0, return_inverse=True)
            inv = inv.reshape(-1)
            lut = np.zeros(inv.max() + 1, dtype=np.int64)
            lut[inv[: verts[d].shape[0]]] = np.arange(1, verts[d].shape[0] + 1)
            cell_faces[d] = lut[inv[verts[d].shape[0]:]].reshape(-1, tab.shape[0])
            if (cell_faces[d] == 0).any():
                raise ValueError("a cell-local face is missing from the face complex")
    return dict(vert=vert, verts=verts, cell_faces=cell_faces, n_parent=n_parent, group=group)


def _lattice(d, k, interior=False, simplex=False):
    if simplex:
        out = []
        for t in itertools.product(*[range(k + 1)] * d):
            e = tuple(reversed(t))
            if sum(e) > k or (interior and (min(e, default=1) < 1 or sum(e) > k - 1)):
                continue
            out.append(e)
        return out
    rng = range(1, k) if interior else range(0, k + 1)
    return [tuple(reversed(t)) for t in itertools.product(*[rng] * d)]


def _map_lattice(t, k, corners, simplex=False):
    """k * Σ_v M_v(t/k) X_v as integers; corners: vertex coordinates of the (sub)element (Q1 map of a cube face, or the
    barycentric map of a simplex face)"""
    d = len(t)
    if simplex:
        X0 = np.asarray(corners[0], dtype=np.float64)
        acc = k * X0
        for m in range(d):
            acc = acc + t[m] * (np.asarray(corners[m + 1], dtype=np.float64) - X0)
        return tuple(int(round(x)) for x in acc)
    acc = np.zeros(len(corners[0]))
    for v, X in enumerate(corners):
        w = 1.0
        for m in range(d):
            w *= (t[m] / k) if (v >> m) & 1 else (1.0 - t[m] / k)
        acc += w * np.asarray(X, dtype=np.float64)
    return tuple(int(round(k * x)) for x in acc)


def element_tables(D, k, n_comp, simplex=False):
    """per d: dofs[d] [nlf][n_face_dofs], own[d] [nlf][n_own], perms[d] [n_perms][n_own] (0-based local dofs / positions)"""
    node_id = {t: i for i, t in enumerate(_lattice(D, k, simplex=simplex))}
    if simplex:
        corner = lambda v: [1 if v - 1 == m else 0 for m in range(D)]               # v0 = origin, v_{m+1} = e_m
        unit_of = lambda d: [[1 if v - 1 == m else 0 for m in range(d)] for v in range(d + 1)]
        vperms = _SVPERMS
    else:
        corner = lambda v: [(v >> m) & 1 for m in range(D)]
        unit_of = lambda d: [[(v >> m) & 1 for m in range(d)] for v in range(2 ** d)]
        vperms = _VPERMS
    dofs, own, perms = {}, {}, {}
    for d in range(D + 1):
        unit = unit_of(d)
        inter = _lattice(d, k, interior=True, simplex=simplex) if d > 0 else [()]
        allp = _lattice(d, k, simplex=simplex) if d > 0 else [()]
        expand = lambda nodes: [n * n_comp + c for n in nodes for c in range(n_comp)]
        rows_all, rows_own = [], []
        for lv in _lfaces(D, d, simplex):
            X = [corner(v) for v in lv]
            rows_all.append(expand([node_id[_map_lattice(t, k, X, simplex)] for t in allp]))
            rows_own.append(expand([node_id[_map_lattice(t, k, X, simplex)] for t in inter]))
        dofs[d] = np.array(rows_all, dtype=np.int64)
        own[d] = np.array(rows_own, dtype=np.int64).reshape(len(rows_own), -1)
        pp = []
        for P in vperms[d]:
            npm = [inter.index(_map_lattice(t, k, [unit[p] for p in P], simplex)) for t in inter] if d > 0 else [0]
            pp.append(expand(npm))
        perms[d] = np.array(pp, dtype=np.int64).reshape(len(vperms[d]), -1)
    return dofs, own, perms


def scalar_or_vector_dofs(mesh: _hp.Mesh, order: int, n_comp: int = 1, dirichlet_boundary=None):
    """-> (cell_dofs [nc, nld] signed Int32 as the reference numbers them, n_free, n_dirichlet,
           free_dof_xyz [n_free, D], dirichlet_dof_xyz [n_dirichlet, D])"""
    D, k = mesh.D, int(order)
    sx = bool(mesh.simplex)
    fc = face_complex(mesh)
    dofs, own, perms = element_tables(D, k, n_comp, sx)
    nc = mesh.n_cells
    lat = np.array(_lattice(D, k, simplex=sx), dtype=np.int64).reshape(-1, D)      # [nls, D]
    nld = lat.shape[0] * n_comp
    vperms = _SVPERMS if sx else _VPERMS
    cell_dofs = np.zeros((nc, nld), dtype=np.int64)
    base = 0
    cv = fc["verts"][D]
    for d in range(D + 1):
        nown = own[d].shape[1]
        nfaces = fc["verts"][d].shape[0]
        if nown:
            tab = _lfaces(D, d, sx)
            for lf in range(tab.shape[0]):
                face = fc["cell_faces"][d][:, lf]                                   # [nc] 1-based
                if 0 < d < D:                                                       # permutation id (topology.jl:593-634)
                    fv = fc["verts"][d][face - 1]                                   # [nc, n_face_vertices]
                    want = cv[:, tab[lf]]
                    P = np.array(vperms[d], dtype=np.int64)                         # [nP, n_face_vertices]
                    ok = (fv[:, P] == want[:, None, :]).all(axis=2)                 # [nc, nP]
                    if not ok.any(axis=1).all():
                        raise ValueError("Valid pindex not found")
                    pindex = np.argmax(ok, axis=1)
                else:
                    pindex = np.zeros(nc, dtype=np.int64)
                cell_dofs[:, own[d][lf]] = base + (face[:, None] - 1) * nown + perms[d][pindex] + 1
        base += nfaces * nown
    ndofs = base
    if (cell_dofs == 0).any():
        raise AssertionError("some local dof was not numbered")
    tag = np.zeros(ndofs, dtype=bool)
    if dirichlet_boundary is not None:
        N = D - 1
        sides = range(1, 2 * D + 1) if dirichlet_boundary == "boundary" else dirichlet_boundary
        face_tag = np.zeros(fc["verts"][N].shape[0], dtype=bool)
        face_tag[: fc["n_parent"][N]] = np.isin(fc["group"][N], np.asarray(list(sides), dtype=np.int64))
        for lf in range(_lfaces(D, N, sx).shape[0]):
            sel = face_tag[fc["cell_faces"][N][:, lf] - 1]
            if sel.any():
                tag[cell_dofs[sel][:, dofs[N][lf]].reshape(-1) - 1] = True
    # physical position of every dof (lattice of the order-times refined mesh)
    npd = np.array([c + 1 for c in mesh.cells_per_dir], dtype=np.int64)
    strides = np.cumprod(np.concatenate(([1], npd[:-1])))
    cnodes = mesh.cell_nodes.astype(np.int64) - 1
    vidx = np.stack([(cnodes // strides[d]) % npd[d] for d in range(D)], axis=2)    # [nc, n_lnodes, D] node lattice index
    if sx:      # barycentric: k V0 + Σ_m t_m (V_m - V0)
        glat = k * vidx[:, None, 0, :] + np.einsum("lm,cmd->cld", lat, vidx[:, 1:, :] - vidx[:, :1, :])
    else:       # tensor lattice from the lowest corner
        glat = k * vidx[:, None, 0, :] + lat[None, :, :]                            # [nc, nls, D]
    pmin = np.array([mesh.domain[2 * d] for d in range(D)])
    pmax = np.array([mesh.domain[2 * d + 1] for d in range(D)])
    ext = k * (npd - 1)
    xyz_l = pmin + (pmax - pmin) * glat / ext                                        # [nc, nls, D]
    dof_xyz = np.zeros((ndofs, D))
    dof_xyz[cell_dofs.reshape(-1) - 1] = np.repeat(xyz_l.reshape(-1, D), n_comp, axis=0)
    free = np.flatnonzero(~tag)
    diri = np.flatnonzero(tag)
    newid = np.empty(ndofs, dtype=np.int64)
    newid[free] = np.arange(1, free.size + 1)
    newid[diri] = -np.arange(1, diri.size + 1)
    return (np.ascontiguousarray(newid[cell_dofs - 1], dtype=np.int32), int(free.size), int(diri.size), dof_xyz[free], dof_xyz[diri])
