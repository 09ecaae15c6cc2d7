"""Simplex quadrature as an engine INPUT: the Duffy rule of the reference (quadrature.jl:108-157): Gauss-Jacobi
(alpha = D-d, beta = 0) points in the collapsed directions, Gauss-Legendre in the last one, n = ceil((degree+1)/2)
per direction, tensor product with the first index fastest, collapsed by duffy_map.  For tetrahedra of degree 1-5 the
reference switches to the symmetric Strang-Fix rules (quadrature.jl:36-52, 500-635): `strang_tet_quadrature` builds them
from their orbits (centroid, vertex-type orbit (s,t,t,t), edge-type orbit (a,a,b,b)) in the reference's point order —
the order is part of the parity contract (the cell loop sums over the points in this order)."""
from __future__ import annotations

import numpy as np
from scipy.special import roots_jacobi


def simplex_quadrature(D: int, degree: int):
    n = int(np.ceil((degree + 1.0) / 2.0))
    xs, ws = [], []
    for d in range(1, D):
        alpha = (D - 1) - (d - 1)
        x, w = roots_jacobi(n, alpha, 0)
        xs.append(0.5 * x + 0.5); ws.append(0.5 * w)
    x, w = np.polynomial.legendre.leggauss(n)
    xs.append(0.5 * x + 0.5); ws.append(0.5 * w)
    a = 0.5
    for d in range(D - 2, -1, -1):
        ws[d] = ws[d] * a
        a *= 0.5
    g = np.meshgrid(*([np.arange(n)] * D), indexing="ij")
    idx = [i.reshape(-1, order="F") for i in g]
    q = np.stack([xs[d][idx[d]] for d in range(D)], axis=1)
    wt = np.ones(n ** D)
    for d in range(D):
        wt = wt * ws[d][idx[d]]
    # duffy_map: m_1 = q_1, m_i = q_i * prod_{j<i} (1 - q_j)
    m = np.empty_like(q)
    acc = np.ones(q.shape[0])
    for i in range(D):
        if i == 0:
            m[:, 0] = q[:, 0]
        else:
            acc = acc * (1.0 - q[:, i - 1])
            m[:, i] = acc * q[:, i]
    return np.ascontiguousarray(m), wt


def _orbit4(s, t):
    """(s,t,t),(t,s,t),(t,t,s),(t,t,t): the 4 points with barycentric coordinates a permutation of (s,t,t,t), with the
    special coordinate moving through x, y, z and finally the implicit fourth one"""
    return [(s, t, t), (t, s, t), (t, t, s), (t, t, t)]


def _orbit6(a, b):
    """the 6 points with barycentric coordinates a permutation of (a,a,b,b), lexicographic in (a before b)"""
    return [(a, a, b), (a, b, a), (a, b, b), (b, a, a), (b, a, b), (b, b, a)]


STRANG_TET_DEGREES = (1, 2, 3, 4, 5)


def strang_tet_quadrature(degree: int):
    """Strang-Fix rules on the unit tetrahedron (weights sum to 1/6), points in the order of quadrature.jl:500-635.
    Degree 3 and 4 have a negative centroid weight."""
    if degree == 1:
        pts, wts = [(0.25, 0.25, 0.25)], [1.0 / 6.0]
    elif degree == 2:
        a, b = 0.5854101966249685, 0.1381966011250105
        o = _orbit4(a, b)
        pts, wts = [o[3], o[0], o[1], o[2]], [1.0 / 24.0] * 4          # (b,b,b) first
    elif degree == 3:
        o = _orbit4(0.5, 1.0 / 6.0)
        pts = [(0.25, 0.25, 0.25), o[3], o[0], o[1], o[2]]
        wts = [-2.0 / 15.0] + [1.5 / 20.0] * 4
    elif degree == 4:
        a, b = 0.3994035761667992, 0.1005964238332008
        pts = [(0.25, 0.25, 0.25)] + _orbit4(11.0 / 14.0, 1.0 / 14.0) + _orbit6(a, b)
        wts = [(-148.0 / 1875.0) / 6.0] + [(343.0 / 7500.0) / 6.0] * 4 + [(56.0 / 375.0) / 6.0] * 6
    elif degree == 5:
        a, b = 0.0673422422100983, 0.3108859192633005
        c, d = 0.7217942490673264, 0.0927352503108912
        e, f = 0.4544962958743506, 0.0455037041256494
        pts = _orbit4(a, b) + _orbit4(c, d) + _orbit6(e, f)
        wts = [0.1126879257180162 / 6.0] * 4 + [0.0734930431163619 / 6.0] * 4 + [0.0425460207770812 / 6.0] * 6
    else:
        raise ValueError("Strang rules exist for degree 1..5")
    return np.ascontiguousarray(np.array(pts, dtype=np.float64)), np.array(wts, dtype=np.float64)
