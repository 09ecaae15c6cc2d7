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


def set_vector_space(n_comp=1, lam=0.0, mu=0.0):
    """state of the C oracle for vector-valued spaces / isotropic elasticity (form 3); n_comp = 1 restores the scalar path"""
    lib().gto_set_vector_space(C.c_int(int(n_comp)), C.c_double(float(lam)), C.c_double(float(mu)))


def assemble(form, coords, cell_nodes, cell_dofs, n_free, tab, alpha=1.0, f_const=1.0, with_vector=True,
             nthreads=1, nnz_cap=None, K_out=None):
    """→ (colptr, rowval, nzval, b, phase_seconds[count, loop, compress, vector]).  K_out: int64 [N_coo] array that
    receives the nz index cache (0-based) for reassemble()."""
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
                            C.c_int64(cap), C.byref(nnz), p(b) if with_vector else None, C.c_int(nthreads), p(tph),
                            p(K_out) if K_out is not None else None)
    if rc != 0:
        return assemble(form, coords, cell_nodes, cell_dofs, n_free, tab, alpha, f_const, with_vector, nthreads, nnz.value, K_out)
    return colptr, rowval[: nnz.value], nzval[: nnz.value], b[:n_free], tph


class Reassembly:
    """update_matrix! / update_vector! on a cached pattern (problems.jl:276-285, 352-361; assembly.jl:577-588): the first
    assembly keeps the COO arrays and the nz index, every step() reruns the cell loops into them and calls
    sparse_matrix!(A, V, cache) / dense_vector!(b, I, V) — no counting loop, no sort."""

    def __init__(self, form, coords, cell_nodes, cell_dofs, n_free, tab, alpha=1.0, f_const=1.0, nthreads=1):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cn = np.ascontiguousarray(cell_nodes, dtype=np.int32)
        self.cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        self.tab = {k: np.ascontiguousarray(tab[k], dtype=np.float64) for k in ("w", "N", "dN", "dM")}
        self.form, self.alpha, self.f_const, self.n_free, self.nthreads = form, alpha, f_const, n_free, nthreads
        nm, nv = C.c_int64(0), C.c_int64(0)
        l = lib()
        l.gto_count(C.c_int64(self.cn.shape[0]), C.c_int(self.cd.shape[1]), self.cd.ctypes.data_as(C.c_void_p), C.byref(nm), C.byref(nv))
        self.K = np.zeros(max(nm.value, 1), dtype=np.int64)
        self.colptr, self.rowval, self.nzval, self.b, self.first_phases = assemble(
            form, self.coords, self.cn, self.cd, n_free, self.tab, alpha, f_const, True, nthreads,
            nnz_cap=None if self.cd.shape[1] <= 8 else int(nm.value), K_out=self.K)
        self.nzval = np.ascontiguousarray(self.nzval)
        self.b = np.ascontiguousarray(self.b)
        self.I = np.zeros(max(nm.value, 1), dtype=np.int32)
        self.J = np.zeros(max(nm.value, 1), dtype=np.int32)
        self.V = np.zeros(max(nm.value, 1), dtype=np.float64)
        self.VI = np.zeros(max(nv.value, 1), dtype=np.int32)
        self.VV = np.zeros(max(nv.value, 1), dtype=np.float64)
        self.phases = np.zeros(3)

    def step(self):
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        t = self.tab
        l = lib()
        l.gto_reassemble.restype = C.c_int
        rc = l.gto_reassemble(C.c_int(self.coords.shape[1]), p(self.coords), C.c_int64(self.cn.shape[0]), C.c_int(self.cn.shape[1]),
                              p(self.cn), C.c_int(self.cd.shape[1]), p(self.cd), C.c_int64(self.n_free), C.c_int(t["w"].size),
                              p(t["w"]), p(t["N"]), p(t["dN"]), p(t["dM"]), C.c_int(self.form), C.c_double(self.alpha),
                              C.c_double(self.f_const), p(self.I), p(self.J), p(self.V), p(self.VI), p(self.VV), p(self.K),
                              C.c_int64(self.nzval.size), p(self.nzval), p(self.b), C.c_int(self.nthreads), p(self.phases))
        assert rc == 0
        return self.nzval, self.b
