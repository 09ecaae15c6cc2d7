"""Order >= 2 Lagrange dof maps by lattice rank (host-side input preparation) — used for SIMPLEXIFIED meshes only; quad /
hex meshes get the reference's own numbering from refnumbering.py.

``cell_dofs`` is an INPUT of the C ABI: in production the Julia host passes ``face_dofs(V)`` as the reference
numbers it (face-complex construction, space.jl:299-535, topology.jl:1594-1704).  That
numbering is NOT restated here.  This module builds a *valid* conforming numbering of the same space instead —
global dof = rank of the node on the order-times-refined node lattice, x fastest — which is all the engine and
the oracle need to be compared on identical inputs (parity of the cell loop, scatter and compression for
high-order elements; the matrix differs from the reference's by a symmetric permutation of rows/columns).
Reference-element node order is the reference's (exponents / order, first index fastest; space.jl:1127-1177).
"""
from __future__ import annotations

import numpy as np

from . import hostprep as _hp


def _cell_lattice(mesh, order: int) -> np.ndarray:
    """[n_cells, n_lscalar, D] integer lattice coordinates (unit = h/order) of every local node of every cell."""
    D = mesh.D
    kind = "P" if mesh.simplex else "Q"
    e = _hp.monomial_exponents(D, order, kind).astype(np.int64)           # [nls, D] reference node * order
    npd = np.array([c + 1 for c in mesh.cells_per_dir], dtype=np.int64)
    strides = np.cumprod(np.concatenate(([1], npd[:-1])))
    cn = mesh.cell_nodes.astype(np.int64) - 1                              # [nc, nln] 0-based mesh nodes
    # mesh-node lattice index per direction
    def node_idx(n):
        return np.stack([(n // strides[d]) % npd[d] for d in range(D)], axis=-1)
    v = node_idx(cn)                                                       # [nc, nln, D]
    if not mesh.simplex:
        base = v[:, 0, :]                                                  # local node 1 = lowest corner (tensor order)
        return order * base[:, None, :] + e[None, :, :]
    # simplex: x = v0 + sum_i (e_i/order) (v_i - v0)
    v0 = v[:, 0, :]
    edges = v[:, 1:, :] - v0[:, None, :]                                   # [nc, D, D]
    return order * v0[:, None, :] + np.einsum("li,cid->cld", e, edges)


def scalar_dofs(mesh, order: int, dirichlet_boundary=None):
    """-> (cell_dofs [nc, nls] 1-based int64, n_dofs, dirichlet_tag [n_dofs] bool, dof_xyz [n_dofs, D])."""
    D = mesh.D
    lat = _cell_lattice(mesh, order)
    ext = np.array([order * c + 1 for c in mesh.cells_per_dir], dtype=np.int64)
    strides = np.cumprod(np.concatenate(([1], ext[:-1])))
    key = (lat * strides[None, None, :]).sum(axis=2)                       # lexicographic lattice id, x fastest
    used, inv = np.unique(key.reshape(-1), return_inverse=True)
    cell_dofs = inv.reshape(key.shape) + 1
    n = used.size
    idx = np.stack([(used // strides[d]) % ext[d] for d in range(D)], axis=1)
    pmin = np.array([mesh.domain[2 * d] for d in range(D)])
    pmax = np.array([mesh.domain[2 * d + 1] for d in range(D)])
    xyz = pmin + (pmax - pmin) * idx / (ext - 1)
    tag = np.zeros(n, dtype=bool)
    if dirichlet_boundary is not None:
        sides = range(1, 2 * D + 1) if dirichlet_boundary == "boundary" else dirichlet_boundary
        for s in sides:                                                    # same side ids as hostprep.boundary_node_mask
            axis = D - 1 - (s - 1) // 2
            upper = (s - 1) % 2 == 1
            tag |= idx[:, axis] == (ext[axis] - 1 if upper else 0)
    return cell_dofs, n, tag, xyz
