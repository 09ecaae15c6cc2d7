"""Host-side inputs of multi-field spaces and skeleton integrals, vectorised numpy.

What the reference derives per face inside its generated loops — the dofs of every field on every cell around the face with
the field's block offset (assembly.jl:321-333, 386-416), and the cell's shape functions at the face's quadrature points for
the (local face, permutation) the face has in that cell (accessors.jl:456-473, 498-522, 1914-1943) — is handed to the engine
as flat tables (gtk_set_space's super dof table, gtk_set_parts).  Input preparation, not the hot path; the literal
loop-for-loop restatement these are checked against is oracle/gt_oracle.py (never imported from here).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import hostprep as _hp


def block_offsets(lengths: Sequence[int]) -> np.ndarray:
    """offsets = blocklasts(dofs) .- map(length, blocks(dofs))  (assembly.jl:321-333)"""
    return np.concatenate(([0], np.cumsum(np.asarray(lengths, dtype=np.int64))[:-1]))


def offset_dofs(spaces: Sequence[_hp.LagrangeSpace]):
    """per field: cell_dofs with free ids shifted by the free offset and Dirichlet ids by the Dirichlet offset (kept negative)
    -> (list of [n_cells, nld_f] int32, n_free_total, n_dirichlet_total, free offsets, Dirichlet offsets)"""
    fo = block_offsets([s.n_free for s in spaces])
    do = block_offsets([s.n_dirichlet for s in spaces])
    out = []
    for s, a, b in zip(spaces, fo, do):
        cd = s.cell_dofs.astype(np.int64)
        out.append(np.where(cd > 0, cd + a, cd - b).astype(np.int32))
    return out, int(sum(s.n_free for s in spaces)), int(sum(s.n_dirichlet for s in spaces)), fo, do


@dataclass
class BlockProblem:
    """What gtk_set_mesh / gtk_set_space / gtk_set_parts take for one integral over a product space."""
    node_coordinates: np.ndarray
    face_nodes: np.ndarray       # integration faces: cells (volume) or interior (D-1)-faces (skeleton)
    manifold_dim: int
    super_dofs: np.ndarray       # [n_faces, L]: field-major, then cell-around-major
    n_free: int
    n_dirichlet: int
    w: np.ndarray
    M: np.ndarray
    dM: np.ndarray
    parts: list                  # dicts(field, side, n_comp, N [n_var, nq, nls], dN or None)
    n_sides: int
    face_var: np.ndarray | None  # [n_faces, n_sides] 0-based tabulation variant
    side_cells: np.ndarray | None = None   # [n_faces, n_sides] 1-based cells around (skeleton)
    cell_nodes: np.ndarray | None = None   # skeleton with gradients: nodes of the D-cells,
    dM_cell: np.ndarray | None = None      #   their geometry gradients at the mapped face points [n_var, nq, nln, D],
    ref_normals: np.ndarray | None = None  #   the reference normal of every variant's local face [n_var, D]

    def part_index(self, field: int, side: int = 0) -> int:
        for k, p in enumerate(self.parts):
            if p["field"] == field and p["side"] == side:
                return k
        raise KeyError((field, side))


def volume_problem(spaces: Sequence[_hp.LagrangeSpace], degree: int) -> BlockProblem:
    """∫(…, measure(interior(mesh), degree)) over V1 × V2 × …: one part per field."""
    mesh = spaces[0].mesh
    if any(s.mesh is not mesh for s in spaces):
        raise ValueError("all fields of a product space must live on the same mesh")
    dofs, nfree, ndiri, _, _ = offset_dofs(spaces)
    q = _hp.quadrature(mesh.D, mesh.simplex, degree)
    kind = spaces[0].kind
    M, dM = _hp.tabulate(mesh.D, 1, kind, q.coordinates)
    parts = []
    for f, s in enumerate(spaces):
        N, dN = _hp.tabulate(mesh.D, s.order, s.kind, q.coordinates)
        parts.append(dict(field=f, side=0, n_comp=s.n_comp, N=N[None], dN=dN[None]))
    return BlockProblem(mesh.node_coordinates, mesh.cell_nodes, mesh.D, np.ascontiguousarray(np.concatenate(dofs, axis=1)),
                        nfree, ndiri, np.ascontiguousarray(q.weights), M, dM, parts, 1, None)


def interior_faces(mesh: _hp.Mesh):
    """GT.skeleton(mesh): the (D-1)-faces with two cells around, in face-id order of the complexified mesh; per face its
    nodes in the face's own vertex order and the two cells around in increasing cell id.
    -> (face_nodes [nf, nv] int32 1-based, side_cells [nf, 2] int64 1-based)"""
    from . import refnumbering as _rn
    fc = _rn.face_complex(mesh)
    d = mesh.D - 1
    cf = fc["cell_faces"][d]                                   # [nc, nlf] 1-based face ids
    nfaces = fc["verts"][d].shape[0]
    count = np.bincount(cf.reshape(-1) - 1, minlength=nfaces)
    inner = np.flatnonzero(count == 2)
    cell_of = np.repeat(np.arange(1, cf.shape[0] + 1), cf.shape[1])
    order = np.argsort(cf.reshape(-1), kind="stable")          # cells ascending inside one face
    sorted_faces = cf.reshape(-1)[order] - 1
    first = np.searchsorted(sorted_faces, inner)
    side_cells = np.stack([cell_of[order][first], cell_of[order][first + 1]], axis=1)
    vertex_node = np.empty(int(fc["vert"].max()) + 1, dtype=np.int64)
    vertex_node[fc["vert"]] = np.arange(1, fc["vert"].shape[0] + 1)
    face_nodes = vertex_node[fc["verts"][d][inner]]
    return np.ascontiguousarray(face_nodes, dtype=np.int32), side_cells


def _reference_normal(Xface: np.ndarray, simplex: bool) -> np.ndarray:
    """normals(mesh(domain(refface)))[ldface] (domain.jl:226, 258 cubes; :428, 460 simplices) of the local face whose reference
    nodes are Xface: the outward unit normal of that face of the reference cell"""
    D = Xface.shape[1]
    for k in range(D):
        if np.all(Xface[:, k] == Xface[0, k]) and (not simplex or Xface[0, k] == 0.0):
            n = np.zeros(D)
            n[k] = 1.0 if Xface[0, k] == 1.0 else -1.0
            return n
    if simplex:                                         # the face opposite the origin
        return np.full(D, 1.0 / np.sqrt(D))
    raise AssertionError("not a face of the reference cell")


def skeleton_problem(spaces: Sequence[_hp.LagrangeSpace], degree: int, gradients: bool = False) -> BlockProblem:
    """∫(…, measure(skeleton(mesh), degree)): per field two parts (the cells around), values only.
    The tabulation variant of (face, side) is named by where the face's nodes sit in the cell: the face point ξ maps to
    Σ_k X̂[loc_k] M_k(ξ) in the cell's reference coordinates (reference_map: the coefficient of the face's k-th shape function
    is the reference coordinate of the cell node that IS the face's k-th node)."""
    mesh = spaces[0].mesh
    if any(s.mesh is not mesh for s in spaces):
        raise ValueError("all fields of a product space must live on the same mesh")
    D = mesh.D
    kind = spaces[0].kind
    fn, sc = interior_faces(mesh)
    cn = mesh.cell_nodes.astype(np.int64)
    loc = np.empty((fn.shape[0], 2, fn.shape[1]), dtype=np.int64)
    for a in range(2):
        eq = cn[sc[:, a] - 1][:, None, :] == fn.astype(np.int64)[:, :, None]       # [nf, nv_face, nln]
        if not eq.any(axis=2).all():
            raise AssertionError("a face node is missing from a cell around the face")
        loc[:, a, :] = eq.argmax(axis=2)
    variants, inv = np.unique(loc.reshape(-1, fn.shape[1]), axis=0, return_inverse=True)
    face_var = inv.reshape(-1, 2).astype(np.int32)
    q = _hp.quadrature(D - 1, mesh.simplex, degree)
    M, dM = _hp.tabulate(D - 1, 1, kind, q.coordinates)                            # geometry functions of the face
    Xref = _hp.reference_nodes(D, 1, kind)                                         # [nln, D] reference nodes of the cell
    cell_pts = np.einsum("qk,vkd->vqd", M, Xref[variants])                         # [n_var, nq, D]
    dofs, nfree, ndiri, _, _ = offset_dofs(spaces)
    parts, cols = [], []
    for f, s in enumerate(spaces):
        tabs = [_hp.tabulate(D, s.order, s.kind, cell_pts[v]) for v in range(variants.shape[0])]
        N = np.stack([t[0] for t in tabs])
        dN = np.stack([t[1] for t in tabs]) if gradients else None
        for a in range(2):
            parts.append(dict(field=f, side=a, n_comp=s.n_comp, N=N, dN=dN))
            cols.append(dofs[f][sc[:, a] - 1])
    bp = BlockProblem(mesh.node_coordinates, fn, D - 1, np.ascontiguousarray(np.concatenate(cols, axis=1)), nfree, ndiri,
                      np.ascontiguousarray(q.weights), M, dM, parts, 2, face_var, sc)
    if gradients:       # geometry of the cells around at the mapped points + reference normals (unit_normal, accessors.jl:1009-1035)
        bp.cell_nodes = mesh.cell_nodes
        bp.dM_cell = np.stack([_hp.tabulate(D, 1, kind, cell_pts[v])[1] for v in range(variants.shape[0])])
        bp.ref_normals = np.stack([_reference_normal(Xref[variants[v]], mesh.simplex) for v in range(variants.shape[0])])
    return bp


def boundary_problem(spaces: Sequence[_hp.LagrangeSpace], sides, degree: int) -> BlockProblem:
    """∫(…, measure(boundary(mesh; group_names), degree)) with the cell around every face (`faces_around = Fill(1)`, mesh.jl:
    249-271): ALL dofs of that cell are pushed, its shape functions, gradients and unit normal are evaluated at the face points
    mapped into the cell — what Nitsche terms need (docs/src/src_jl/example_hello_world_dg.jl:70-76)."""
    mesh = spaces[0].mesh
    D = mesh.D
    kind = spaces[0].kind
    fn, fc, _ = _hp.boundary_faces(mesh, sides)
    fn = np.ascontiguousarray(fn, dtype=np.int32)
    if mesh.simplex:
        # boundary_faces names the parent HEXAHEDRON; the simplex around a boundary face is the one cell of the face complex that
        # holds it (the pre-existing boundary faces keep the ids 1..n_parent in the order boundary_faces lists them)
        from . import refnumbering as _rn
        cplx = _rn.face_complex(mesh)
        cf = cplx["cell_faces"][D - 1]
        n_parent = cplx["n_parent"][D - 1]
        owner = np.zeros(n_parent + 1, dtype=np.int64)
        cell_of = np.repeat(np.arange(1, cf.shape[0] + 1), cf.shape[1])
        is_b = cf.reshape(-1) <= n_parent
        owner[cf.reshape(-1)[is_b]] = cell_of[is_b]
        all_nodes, all_group, _ = _rn.boundary_face_nodes(mesh, D - 1)
        keep = np.ones(all_group.shape[0], dtype=bool) if sides is None else np.isin(all_group, np.asarray(list(sides), dtype=np.int64))
        sc = owner[1:][keep][:, None]
        if (sc == 0).any() or sc.shape[0] != fn.shape[0]:
            raise AssertionError("a boundary face has no cell around in the face complex")
    else:
        sc = (np.asarray(fc, dtype=np.int64) + 1)[:, None]                      # the one cell around, 1-based
    cn = mesh.cell_nodes.astype(np.int64)
    eq = cn[sc[:, 0] - 1][:, None, :] == fn.astype(np.int64)[:, :, None]
    if not eq.any(axis=2).all():
        raise AssertionError("a boundary-face node is missing from the cell around the face")
    loc = eq.argmax(axis=2)
    variants, inv = np.unique(loc, axis=0, return_inverse=True)
    face_var = inv.reshape(-1, 1).astype(np.int32)
    q = _hp.quadrature(D - 1, mesh.simplex, degree)
    M, dM = _hp.tabulate(D - 1, 1, kind, q.coordinates)
    Xref = _hp.reference_nodes(D, 1, kind)
    cell_pts = np.einsum("qk,vkd->vqd", M, Xref[variants])
    dofs, nfree, ndiri, _, _ = offset_dofs(spaces)
    parts, cols = [], []
    for f, s in enumerate(spaces):
        tabs = [_hp.tabulate(D, s.order, s.kind, cell_pts[v]) for v in range(variants.shape[0])]
        parts.append(dict(field=f, side=0, n_comp=s.n_comp, N=np.stack([t[0] for t in tabs]), dN=np.stack([t[1] for t in tabs])))
        cols.append(dofs[f][sc[:, 0] - 1])
    bp = BlockProblem(mesh.node_coordinates, fn, D - 1, np.ascontiguousarray(np.concatenate(cols, axis=1)), nfree, ndiri,
                      np.ascontiguousarray(q.weights), M, dM, parts, 1, face_var, sc)
    bp.cell_nodes = mesh.cell_nodes
    bp.dM_cell = np.stack([_hp.tabulate(D, 1, kind, cell_pts[v])[1] for v in range(variants.shape[0])])
    bp.ref_normals = np.stack([_reference_normal(Xref[variants[v]], mesh.simplex) for v in range(variants.shape[0])])
    return bp


def face_point_coordinates(bp: BlockProblem) -> np.ndarray:
    """physical coordinates of the quadrature points of every integration face: [n_faces, n_q, D]"""
    return np.einsum("qk,fkd->fqd", bp.M, bp.node_coordinates[bp.face_nodes.astype(np.int64) - 1])
