"""ctypes wrapper of oracle/libgtoracle.so (C restatement; test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libgtoracle.so")


def build(force=False):
    src = os.path.join(HERE, "gt_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "libgtoracle.so"])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO)
        _lib.gto_assemble.restype = C.c_int
    return _lib


def assemble(form, coords, cell_nodes, cell_dofs, n_free, tab, alpha=1.0, f_const=1.0, with_vector=True,
             nthreads=1, nnz_cap=None):
    """→ (colptr, rowval, nzval, b, phase_seconds[count, loop, compress, vector])."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    cn = np.ascontiguousarray(cell_nodes, dtype=np.int32)
    cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
    w, N, dN, dM = (np.ascontiguousarray(tab[k], dtype=np.float64) for k in ("w", "N", "dN", "dM"))
    D = coords.shape[1]
    nc, nln = cn.shape
    nld = cd.shape[1]
    cap = int(nnz_cap if nnz_cap is not None else min(nc * nld * nld, max(1, n_free) * 27 if D == 3 else max(1, n_free) * 9) + 16)
    colptr = np.zeros(n_free + 1, dtype=np.int32)
    rowval = np.zeros(cap, dtype=np.int32)
    nzval = np.zeros(cap, dtype=np.float64)
    b = np.zeros(max(n_free, 1), dtype=np.float64)
    nnz = C.c_int64(0)
    tph = np.zeros(4)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib().gto_assemble(C.c_int(D), C.c_int64(coords.shape[0]), p(coords), C.c_int64(nc), C.c_int(nln), p(cn),
                            C.c_int(nld), p(cd), C.c_int64(n_free), C.c_int(w.size), p(w), p(N), p(dN), p(dM),
                            C.c_int(form), C.c_double(alpha), C.c_double(f_const), p(colptr), p(rowval), p(nzval),
                            C.c_int64(cap), C.byref(nnz), p(b) if with_vector else None, C.c_int(nthreads), p(tph))
    if rc != 0:
        return assemble(form, coords, cell_nodes, cell_dofs, n_free, tab, alpha, f_const, with_vector, nthreads, nnz.value)
    return colptr, rowval[: nnz.value], nzval[: nnz.value], b[:n_free], tph
