"""Simplex quadrature as an engine INPUT: the Duffy rule of the reference (quadrature.jl:108-157): Gauss-Jacobi
(alpha = D-d, beta = 0) points in the collapsed directions, Gauss-Legendre in the last one, n = ceil((degree+1)/2)
per direction, tensor product with the first index fastest, collapsed by duffy_map.  The reference switches to
tabulated Strang rules for tetrahedra of degree 1-5 (quadrature.jl:36-52, 500-635); those tables are data of the
reference and are not restated — the Julia host passes its own points/weights through gtk_set_tabulation."""
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
