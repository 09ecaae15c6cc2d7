"""ctypes binding of libgtkasm's C ABI (include/gtk_assembly.h).

This is the stand-in for the Julia `ccall` shim: the same symbols,
the same argument order, plain pointers and sizes.  There is no CPU fallback: if the
shared library is missing, or no GPU is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgtkasm.so")

GTK_OK = 0
GTK_ERR_INVALID, GTK_ERR_CUDA, GTK_ERR_UNSUPPORTED_FORM, GTK_ERR_STATE, GTK_ERR_TOO_LARGE, GTK_ERR_NCCL = -1, -2, -3, -4, -5, -6
FREE, DIRICHLET = 1, 2
FORM_LAPLACE, FORM_MASS, FORM_ELASTICITY_ISO, FORM_PLAPLACE_JACOBIAN = 1, 2, 3, 4
FORM_SOURCE_CONST, FORM_SOURCE_NODAL, FORM_SOURCE_QP, FORM_PLAPLACE_RESIDUAL = 101, 102, 103, 104
SCALAR_VOLUME, SCALAR_L2SQ, SCALAR_H1SQ = 200, 201, 202

# every symbol include/gtk_assembly.h declares (tests check the .so exports all of them)
ABI_SYMBOLS = [
    "gtk_version", "gtk_create", "gtk_destroy", "gtk_last_error", "gtk_set_stream",
    "gtk_set_mesh", "gtk_set_active_cells", "gtk_update_coordinates", "gtk_set_space", "gtk_set_tabulation",
    "gtk_matrix_symbolic", "gtk_matrix_pattern", "gtk_matrix_numeric", "gtk_matrix_numeric_device",
    "gtk_vector_symbolic", "gtk_vector_assemble", "gtk_vector_assemble_device",
    "gtk_assemble_matrix_and_vector", "gtk_assemble_matrix_and_vector_device",
    "gtk_device_pointer", "gtk_copy_nzval", "gtk_copy_vector", "gtk_info",
    "gtk_comm_unique_id", "gtk_comm_init", "gtk_comm_set_exchange", "gtk_comm_sum_ghost_rows",
    "gtk_assemble_and_sum_ghost_rows_device", "gtk_select_matrix", "gtk_matvec_add_device", "gtk_matvec_add", "gtk_set_manifold_dim", "gtk_set_vector", "gtk_comm_p2p_export", "gtk_comm_p2p_import",
    "gtk_comm_ghost_info", "gtk_set_profiling", "gtk_profile_count", "gtk_profile_get",
    "gtk_field_set_values", "gtk_field_set_values_device", "gtk_field_get_values", "gtk_field_axpy_free",
    "gtk_space_dof_coordinates", "gtk_scalar_assemble", "gtk_comm_build_exchange", "gtk_comm_connect_peer_memory",
    "gtk_set_cartesian_q1_problem", "gtk_copy_device_array", "gtk_matrix_pattern_i64",
    "gtk_set_parts", "gtk_matrix_numeric_blocks", "gtk_matrix_numeric_blocks_device",
    "gtk_vector_assemble_blocks", "gtk_vector_assemble_blocks_device",
    "gtk_matrix_sum_symbolic", "gtk_matrix_sum_numeric", "gtk_matrix_sum_numeric_device", "gtk_set_skeleton_cells",
    "gtk_matrix_colptr_at", "gtk_vector_assemble_blocks_data", "gtk_vector_assemble_blocks_data_device",
]
BLOCK_ZERO, BLOCK_MASS, BLOCK_LAPLACE, BLOCK_VALU_DIVV, BLOCK_DIVU_VALV, BLOCK_IP, BLOCK_IP_NOH = 0, 1, 2, 3, 4, 5, 6
MAX_PARTS = 8


class GtkError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgtkasm error {code}: {msg}")
        self.code = code


class UnsupportedFormError(GtkError):
    """The engine does not recognise the form; it never falls back to the CPU."""


class FormParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("lam", C.c_double), ("mu", C.c_double),
                ("f_const", C.c_double * 3), ("f_nodal", C.c_void_p), ("f_qp", C.c_void_p),
                ("coef_nodal", C.c_void_p), ("coef_qp", C.c_void_p), ("accumulate", C.c_int32),
                ("exponent", C.c_double)]


class Part(C.Structure):       # gtk_part
    _fields_ = [("n_lshape", C.c_int32), ("n_comp", C.c_int32), ("side", C.c_int32), ("N", C.c_void_p), ("dN", C.c_void_p)]


class Block(C.Structure):      # gtk_block
    _fields_ = [("part_u", C.c_int32), ("part_v", C.c_int32), ("form", C.c_int32), ("alpha", C.c_double), ("c", C.c_double * 3)]


class VBlock(C.Structure):     # gtk_vblock
    _fields_ = [("part", C.c_int32), ("alpha", C.c_double), ("f_const", C.c_double * 3), ("c", C.c_double * 3)]


_lib = None


def load_library() -> C.CDLL:
    """dlopen libgtkasm.so (built in-tree by build.py).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    sig = {
        "gtk_version": (i32, []),
        "gtk_create": (i32, [i32, C.POINTER(vp)]),
        "gtk_destroy": (i32, [vp]),
        "gtk_last_error": (C.c_char_p, [vp]),
        "gtk_set_stream": (i32, [vp, vp]),
        "gtk_set_mesh": (i32, [vp, i32, i64, vp, i64, i32, vp]),
        "gtk_update_coordinates": (i32, [vp, vp]),
        "gtk_set_space": (i32, [vp, i32, i32, vp, i64, i64]),
        "gtk_set_tabulation": (i32, [vp, i32, vp, vp, vp, vp, vp]),
        "gtk_matrix_symbolic": (i32, [vp, i32, i32, C.POINTER(i64)]),
        "gtk_matrix_pattern": (i32, [vp, vp, vp]),
        "gtk_matrix_numeric": (i32, [vp, i32, C.POINTER(FormParams), vp]),
        "gtk_matrix_numeric_device": (i32, [vp, i32, C.POINTER(FormParams)]),
        "gtk_vector_symbolic": (i32, [vp, i32]),
        "gtk_vector_assemble": (i32, [vp, i32, C.POINTER(FormParams), vp]),
        "gtk_vector_assemble_device": (i32, [vp, i32, C.POINTER(FormParams)]),
        "gtk_assemble_matrix_and_vector": (i32, [vp, i32, C.POINTER(FormParams), i32, C.POINTER(FormParams), vp, vp]),
        "gtk_assemble_matrix_and_vector_device": (i32, [vp, i32, C.POINTER(FormParams), i32, C.POINTER(FormParams)]),
        "gtk_device_pointer": (i32, [vp, i32, C.POINTER(vp), C.POINTER(i64)]),
        "gtk_copy_nzval": (i32, [vp, vp]),
        "gtk_copy_vector": (i32, [vp, vp]),
        "gtk_info": (i64, [vp, i32]),
        "gtk_comm_unique_id": (i32, [vp]),
        "gtk_comm_init": (i32, [vp, i32, i32, vp]),
        "gtk_set_active_cells": (i32, [vp, i64, i64]),
        "gtk_comm_set_exchange": (i32, [vp, i32, i64, vp, i64, vp, i64, vp, i64, vp]),
        "gtk_comm_sum_ghost_rows": (i32, [vp]),
        "gtk_select_matrix": (i32, [vp, i32]),
        "gtk_comm_p2p_export": (i32, [vp, i32, C.c_void_p]),
        "gtk_comm_p2p_import": (i32, [vp, i32, C.c_void_p]),
        "gtk_set_manifold_dim": (i32, [vp, i32]),
        "gtk_set_vector": (i32, [vp, C.c_void_p]),
        "gtk_matvec_add_device": (i32, [vp, C.c_double, C.c_void_p, C.c_double]),
        "gtk_matvec_add": (i32, [vp, C.c_double, C.c_void_p, C.c_double, C.c_void_p]),
        "gtk_assemble_and_sum_ghost_rows_device": (i32, [vp, i32, C.POINTER(FormParams), i32, C.POINTER(FormParams)]),
        "gtk_comm_ghost_info": (i64, [vp, i32]),
        "gtk_set_profiling": (i32, [vp, i32]),
        "gtk_profile_count": (i32, [vp]),
        "gtk_profile_get": (i32, [vp, i32, C.c_char_p, C.POINTER(C.c_double)]),
        "gtk_field_set_values": (i32, [vp, vp, vp]),
        "gtk_field_set_values_device": (i32, [vp, vp, vp]),
        "gtk_field_get_values": (i32, [vp, vp, vp]),
        "gtk_field_axpy_free": (i32, [vp, C.c_double, vp]),
        "gtk_space_dof_coordinates": (i32, [vp, vp, vp, vp]),
        "gtk_scalar_assemble": (i32, [vp, i32, C.POINTER(FormParams), C.POINTER(C.c_double)]),
        "gtk_comm_build_exchange": (i32, [vp, i64, vp]),
        "gtk_comm_connect_peer_memory": (i32, [vp]),
        "gtk_copy_device_array": (i32, [vp, i32, vp, i64]),
        "gtk_matrix_pattern_i64": (i32, [vp, vp, vp]),
        "gtk_set_cartesian_q1_problem": (i32, [vp, vp, vp, i64, i64, i32, C.POINTER(i64), C.POINTER(i64)]),
        "gtk_set_parts": (i32, [vp, i32, vp, vp, vp, i32, C.POINTER(Part), i32, i32, vp]),
        "gtk_matrix_numeric_blocks": (i32, [vp, i32, C.POINTER(Block), vp]),
        "gtk_matrix_numeric_blocks_device": (i32, [vp, i32, C.POINTER(Block)]),
        "gtk_vector_assemble_blocks": (i32, [vp, i32, C.POINTER(VBlock), i32, vp]),
        "gtk_vector_assemble_blocks_device": (i32, [vp, i32, C.POINTER(VBlock), i32]),
        "gtk_set_skeleton_cells": (i32, [vp, i64, i32, vp, vp, vp, vp]),
        "gtk_matrix_colptr_at": (i32, [vp, i32, vp, vp]),
        "gtk_vector_assemble_blocks_data": (i32, [vp, i32, C.POINTER(VBlock), vp, i32, vp]),
        "gtk_vector_assemble_blocks_data_device": (i32, [vp, i32, C.POINTER(VBlock), vp, i32]),
        "gtk_matrix_sum_symbolic": (i32, [vp, i32, C.POINTER(vp), C.POINTER(i64)]),
        "gtk_matrix_sum_numeric": (i32, [vp, i32, C.POINTER(vp), vp]),
        "gtk_matrix_sum_numeric_device": (i32, [vp, i32, C.POINTER(vp)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def make_params(alpha=1.0, lam=0.0, mu=0.0, f_const=None, f_nodal=None, f_qp=None, accumulate=False,
                coef_nodal=None, coef_qp=None, exponent=0.0):
    """Returns (FormParams, keepalive) — keepalive holds the numpy buffers the struct points to."""
    p = FormParams()
    p.alpha, p.lam, p.mu = float(alpha), float(lam), float(mu)
    p.exponent = float(exponent)
    p.accumulate = 1 if accumulate else 0
    fc = np.zeros(3)
    if f_const is not None:
        v = np.atleast_1d(np.asarray(f_const, dtype=np.float64)).reshape(-1)
        fc[: v.size] = v
    for k in range(3):
        p.f_const[k] = fc[k]
    keep = []
    if f_nodal is not None:
        a = _f64(f_nodal); keep.append(a); p.f_nodal = a.ctypes.data
    if f_qp is not None:
        a = _f64(f_qp); keep.append(a); p.f_qp = a.ctypes.data
    if coef_nodal is not None:
        a = _f64(coef_nodal); keep.append(a); p.coef_nodal = a.ctypes.data
    if coef_qp is not None:
        a = _f64(coef_qp); keep.append(a); p.coef_qp = a.ctypes.data
    return p, keep


class Engine:
    """One gtk_ctx on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.gtk_create(int(device), C.byref(h))
        if rc != GTK_OK or not h.value:
            raise GtkError(rc, f"gtk_create(device={device}) failed — a CUDA GPU is required (no CPU fallback)")
        self.h = h
        self.device = device
        self.nnz = 0
        self.n_rows = 0
        self.n_cols = 0
        self.n_vec_rows = 0

    # -- plumbing ---------------------------------------------------------------
    def _ck(self, rc: int):
        if rc != GTK_OK:
            msg = self.lib.gtk_last_error(self.h)
            msg = msg.decode() if msg else ""
            if rc == GTK_ERR_UNSUPPORTED_FORM:
                raise UnsupportedFormError(rc, msg)
            raise GtkError(rc, msg)

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.gtk_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_handle: int):
        self._ck(self.lib.gtk_set_stream(self.h, C.c_void_p(cuda_stream_handle)))

    # -- inputs -------------------------------------------------------------------
    def set_mesh(self, node_coordinates, cell_nodes):
        xyz = _f64(node_coordinates)
        cn = _i32(cell_nodes)
        self._D = xyz.shape[1]
        self._n_cells = cn.shape[0]
        self._ck(self.lib.gtk_set_mesh(self.h, xyz.shape[1], xyz.shape[0], _ptr(xyz), cn.shape[0], cn.shape[1], _ptr(cn)))

    def set_cartesian_q1_problem(self, domain, cells, kz0: int = 0, kz1: Optional[int] = None, slab_local: bool = False):
        """mesh + Q1 space (whole boundary Dirichlet) of GT.cartesian_mesh(domain, cells) generated in HBM; node layers
        [kz0, kz1] only (multi-GPU slabs).  Replaces set_mesh + set_space.  -> (n_free, n_dirichlet)"""
        dom = _f64(domain)
        cl = np.ascontiguousarray(cells, dtype=np.int64)
        kz1 = int(cl[2]) if kz1 is None else int(kz1)
        nf, nd = C.c_int64(0), C.c_int64(0)
        self._ck(self.lib.gtk_set_cartesian_q1_problem(self.h, _ptr(dom), _ptr(cl), int(kz0), kz1, 1 if slab_local else 0,
                                                       C.byref(nf), C.byref(nd)))
        self._D = 3
        self._n_cells = int(cl[0] * cl[1] * (kz1 - kz0))
        self._n_free, self._n_diri = nf.value, nd.value
        return nf.value, nd.value

    def copy_device_array(self, which: int, dtype) -> np.ndarray:
        """host copy of a device array named by gtk_device_pointer's selector (tests / diagnostics)"""
        _, n = self.device_pointer(which)
        out = np.empty(n, dtype=dtype)
        self._ck(self.lib.gtk_copy_device_array(self.h, which, _ptr(out), out.nbytes))
        return out

    def set_manifold_dim(self, d: int):
        """Cells of reference dimension d < D (boundary faces as a mesh of their own); call between set_mesh and set_tabulation."""
        self._ck(self.lib.gtk_set_manifold_dim(self.h, int(d)))

    def set_vector(self, b: np.ndarray):
        """Upload the vector an `accumulate=True` linear-form assembly adds to."""
        b = np.ascontiguousarray(b, dtype=np.float64)
        if b.size != self.n_vec_rows:
            raise ValueError(f"b has {b.size} entries, the vector selection has {self.n_vec_rows} rows")
        self._ck(self.lib.gtk_set_vector(self.h, _ptr(b)))

    def set_active_cells(self, first: int, count: int):
        self._ck(self.lib.gtk_set_active_cells(self.h, int(first), int(count)))

    def update_coordinates(self, node_coordinates):
        xyz = _f64(node_coordinates)
        self._ck(self.lib.gtk_update_coordinates(self.h, _ptr(xyz)))

    def set_space(self, cell_dofs, n_free: int, n_dirichlet: int, n_comp: int = 1):
        cd = _i32(cell_dofs)
        self._n_free, self._n_diri = int(n_free), int(n_dirichlet)
        self._ck(self.lib.gtk_set_space(self.h, cd.shape[1], n_comp, _ptr(cd), n_free, n_dirichlet))

    def set_tabulation(self, w, N, dN, M, dM):
        w, N, dN, M, dM = map(_f64, (w, N, dN, M, dM))
        self._ck(self.lib.gtk_set_tabulation(self.h, w.shape[0], _ptr(w), _ptr(N), _ptr(dN), _ptr(M), _ptr(dM)))

    def set_parts(self, w, M, dM, parts, n_sides: int = 1, face_var=None):
        """parts: list of dicts(n_comp, side, N [n_var, n_q, n_lshape], dN [n_var, n_q, n_lshape, D] or None) in the order of
        the super dof table handed to set_space (product spaces / skeleton integrals)"""
        w, M, dM = map(_f64, (w, M, dM))
        arr = (Part * len(parts))()
        keep = []
        n_var = None
        for k, pd in enumerate(parts):
            N = _f64(pd["N"])
            if N.ndim == 2:
                N = N[None]
            dN = None if pd.get("dN") is None else _f64(pd["dN"])
            if dN is not None and dN.ndim == 3:
                dN = dN[None]
            keep += [N, dN]
            if n_var is None:
                n_var = N.shape[0]
            if N.shape[0] != n_var or N.shape[1] != w.shape[0] or (dN is not None and dN.shape[:3] != N.shape):
                raise ValueError("part tabulations must share n_var and n_q")
            arr[k].n_lshape, arr[k].n_comp, arr[k].side = N.shape[2], int(pd.get("n_comp", 1)), int(pd.get("side", 0))
            arr[k].N = N.ctypes.data
            arr[k].dN = None if dN is None else dN.ctypes.data
        fv = None if face_var is None else _i32(face_var)
        self._ck(self.lib.gtk_set_parts(self.h, w.shape[0], _ptr(w), _ptr(M), _ptr(dM), len(parts), arr, int(n_sides), int(n_var), _ptr(fv)))
        del keep

    @staticmethod
    def _block_array(blocks):
        blocks = list(blocks)
        arr = (Block * max(len(blocks), 1))()
        for k, blk in enumerate(blocks):
            pu, pv, form, alpha = blk[:4]
            arr[k].part_u, arr[k].part_v, arr[k].form, arr[k].alpha = int(pu), int(pv), int(form), float(alpha)
            c = blk[4] if len(blk) > 4 else (0.0, 0.0, 0.0)
            for i in range(3):
                arr[k].c[i] = float(c[i])
        return arr, len(blocks)

    def set_skeleton_cells(self, cell_nodes, side_cells, dM_cell, ref_normals):
        """cells around the faces of a skeleton measure (gradients / normals on the faces: BLOCK_IP); after set_parts"""
        cn, sc = _i32(cell_nodes), _i32(side_cells)
        dm, nr = _f64(dM_cell), _f64(ref_normals)
        self._ck(self.lib.gtk_set_skeleton_cells(self.h, cn.shape[0], cn.shape[1], _ptr(cn), _ptr(sc), _ptr(dm), _ptr(nr)))

    def matrix_numeric_blocks(self, blocks, out: Optional[np.ndarray] = None) -> np.ndarray:
        """blocks: iterable of (part_u, part_v, form, alpha[, (c0, c1, c2)])"""
        arr, n = self._block_array(blocks)
        nz = np.empty(self.nnz, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_matrix_numeric_blocks(self.h, n, arr, _ptr(nz)))
        return nz

    def matrix_numeric_blocks_device(self, blocks):
        arr, n = self._block_array(blocks)
        self._ck(self.lib.gtk_matrix_numeric_blocks_device(self.h, n, arr))

    # -- sums of integrals over different domains: this context holds the merged matrix ------------
    def matrix_sum_symbolic(self, sources) -> int:
        """union pattern of the matrices assembled by the `sources` engines (same row / column selection)"""
        arr = (C.c_void_p * len(sources))(*[e.h.value for e in sources])
        nnz = C.c_int64(0)
        self._ck(self.lib.gtk_matrix_sum_symbolic(self.h, len(sources), arr, C.byref(nnz)))
        self.nnz, self.n_rows, self.n_cols = nnz.value, sources[0].n_rows, sources[0].n_cols
        self._n_free, self._n_diri = sources[0]._n_free, sources[0]._n_diri
        return self.nnz

    def matrix_sum_numeric(self, sources, out: Optional[np.ndarray] = None) -> np.ndarray:
        arr = (C.c_void_p * len(sources))(*[e.h.value for e in sources])
        nz = np.empty(self.nnz, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_matrix_sum_numeric(self.h, len(sources), arr, _ptr(nz)))
        return nz

    def vector_assemble_blocks(self, vblocks, accumulate: bool = False, out: Optional[np.ndarray] = None, g_qp=None) -> np.ndarray:
        """vblocks: iterable of (part, alpha, f_const) — or, with data g_qp [n_faces, n_q], (part, alpha, (c0, c1, c2)):
        ∫ alpha g (c0 v + (c1/h) v + c2 n⋅∇v)"""
        if self.n_vec_rows == 0:
            self.vector_symbolic(FREE)
        vblocks = list(vblocks)
        arr = (VBlock * max(len(vblocks), 1))()
        for k, (part, alpha, f) in enumerate(vblocks):
            arr[k].part, arr[k].alpha = int(part), float(alpha)
            fv = np.zeros(3)
            f = np.atleast_1d(np.asarray(f, dtype=np.float64)).reshape(-1)
            fv[: f.size] = f
            for c in range(3):
                arr[k].f_const[c] = 0.0 if g_qp is not None else fv[c]
                arr[k].c[c] = fv[c] if g_qp is not None else 0.0
        b = np.empty(self.n_vec_rows, dtype=np.float64) if out is None else out
        if g_qp is None:
            self._ck(self.lib.gtk_vector_assemble_blocks(self.h, len(vblocks), arr, 1 if accumulate else 0, _ptr(b)))
        else:
            g = _f64(g_qp)
            self._ck(self.lib.gtk_vector_assemble_blocks_data(self.h, len(vblocks), arr, _ptr(g), 1 if accumulate else 0, _ptr(b)))
        return b

    # -- matrix ---------------------------------------------------------------------
    def select_matrix(self, slot: int):
        """Make matrix `slot` (0..3) the target of the matrix_* calls; every slot keeps pattern, plans and values."""
        self._ck(self.lib.gtk_select_matrix(self.h, int(slot)))
        if not hasattr(self, "_slot_dims"):
            self._slot_dims, self._slot = {}, 0
        self._slot_dims[self._slot] = (self.nnz, self.n_rows, self.n_cols)
        self._slot = int(slot)
        self.nnz, self.n_rows, self.n_cols = self._slot_dims.get(self._slot, (0, 0, 0))

    def matvec_add(self, alpha: float, x: np.ndarray, beta: float, out: Optional[np.ndarray] = None) -> np.ndarray:
        """b = beta*b + alpha*M*x on device (Julia's 5-argument mul! for SparseMatrixCSC, bitwise); returns b."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.n_cols:
            raise ValueError(f"x has {x.size} entries, the selected matrix has {self.n_cols} columns")
        b = np.empty(self.n_vec_rows, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_matvec_add(self.h, float(alpha), _ptr(x), float(beta), _ptr(b)))
        return b

    def matvec_add_device(self, alpha: float, x: np.ndarray, beta: float):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.n_cols:
            raise ValueError(f"x has {x.size} entries, the selected matrix has {self.n_cols} columns")
        self._ck(self.lib.gtk_matvec_add_device(self.h, float(alpha), _ptr(x), float(beta)))

    def matrix_symbolic(self, rows=FREE, cols=FREE) -> int:
        nnz = C.c_int64(0)
        self._ck(self.lib.gtk_matrix_symbolic(self.h, rows, cols, C.byref(nnz)))
        self.nnz = nnz.value
        self.n_rows = self._n_free if rows == FREE else self._n_diri
        self.n_cols = self._n_free if cols == FREE else self._n_diri
        return self.nnz

    def matrix_pattern(self, want_rowval: bool = True):
        colptr = np.empty(self.n_cols + 1, dtype=np.int32)
        rowval = np.empty(self.nnz, dtype=np.int32) if want_rowval else None
        self._ck(self.lib.gtk_matrix_pattern(self.h, _ptr(colptr), _ptr(rowval)))
        return colptr, rowval

    def matrix_colptr_at(self, cols) -> np.ndarray:
        """0-based colptr entries of a few (0-based) columns, straight from the device"""
        c = np.ascontiguousarray(cols, dtype=np.int64)
        out = np.empty(c.size, dtype=np.int64)
        self._ck(self.lib.gtk_matrix_colptr_at(self.h, c.size, _ptr(c), _ptr(out)))
        return out

    def matrix_pattern_i64(self, want_rowval: bool = True):
        """colptr / rowval as Int64 (assembly_options index_type = Int64; no 2^31 limit on nnz)"""
        colptr = np.empty(self.n_cols + 1, dtype=np.int64)
        rowval = np.empty(self.nnz, dtype=np.int64) if want_rowval else None
        self._ck(self.lib.gtk_matrix_pattern_i64(self.h, _ptr(colptr), _ptr(rowval)))
        return colptr, rowval

    def matrix_numeric(self, form: int, out: Optional[np.ndarray] = None, **params) -> np.ndarray:
        p, keep = make_params(**params)
        nz = np.empty(self.nnz, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_matrix_numeric(self.h, form, C.byref(p), _ptr(nz)))
        del keep
        return nz

    def matrix_numeric_device(self, form: int, **params):
        p, keep = make_params(**params)
        self._ck(self.lib.gtk_matrix_numeric_device(self.h, form, C.byref(p)))
        del keep

    # -- vector ---------------------------------------------------------------------
    def vector_symbolic(self, fd=FREE):
        self._ck(self.lib.gtk_vector_symbolic(self.h, fd))
        self.n_vec_rows = self._n_free if fd == FREE else self._n_diri

    def vector_assemble(self, form: int, out: Optional[np.ndarray] = None, **params) -> np.ndarray:
        if self.n_vec_rows == 0 and self.lib.gtk_info(self.h, 3) >= 0:
            self.vector_symbolic(FREE)
        p, keep = make_params(**params)
        b = np.empty(self.n_vec_rows, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_vector_assemble(self.h, form, C.byref(p), _ptr(b)))
        del keep
        return b

    def vector_assemble_device(self, form: int, **params):
        p, keep = make_params(**params)
        self._ck(self.lib.gtk_vector_assemble_device(self.h, form, C.byref(p)))
        del keep

    # -- both -----------------------------------------------------------------------
    def assemble_matrix_and_vector(self, mform: int, mparams: dict, vform: int, vparams: dict,
                                   nzval: Optional[np.ndarray] = None, b: Optional[np.ndarray] = None):
        if self.n_vec_rows == 0:
            self.vector_symbolic(FREE)
        pm, k1 = make_params(**mparams)
        pv, k2 = make_params(**vparams)
        nz = np.empty(self.nnz, dtype=np.float64) if nzval is None else nzval
        bb = np.empty(self.n_vec_rows, dtype=np.float64) if b is None else b
        self._ck(self.lib.gtk_assemble_matrix_and_vector(self.h, mform, C.byref(pm), vform, C.byref(pv), _ptr(nz), _ptr(bb)))
        del k1, k2
        return nz, bb

    def assemble_matrix_and_vector_device(self, mform: int, mparams: dict, vform: int, vparams: dict):
        if self.n_vec_rows == 0:
            self.vector_symbolic(FREE)
        pm, k1 = make_params(**mparams)
        pv, k2 = make_params(**vparams)
        self._ck(self.lib.gtk_assemble_matrix_and_vector_device(self.h, mform, C.byref(pm), vform, C.byref(pv)))
        del k1, k2

    # -- DiscreteField parameter, interpolation inputs, scalar integrals ---------------
    def field_set_values(self, free_values=None, dirichlet_values=None):
        """u_h of the current space: free / Dirichlet values to HBM (None = leave as is)."""
        fv = None if free_values is None else _f64(free_values)
        dv = None if dirichlet_values is None else _f64(dirichlet_values)
        if fv is not None and fv.size != self._n_free:
            raise ValueError(f"free_values has {fv.size} entries, the space has {self._n_free} free dofs")
        if dv is not None and dv.size != self._n_diri:
            raise ValueError(f"dirichlet_values has {dv.size} entries, the space has {self._n_diri} Dirichlet dofs")
        self._ck(self.lib.gtk_field_set_values(self.h, _ptr(fv), _ptr(dv)))

    def field_set_values_device(self, d_free: Optional[int] = None, d_dirichlet: Optional[int] = None):
        """same from device pointers (ints), device-to-device on the engine's stream"""
        self._ck(self.lib.gtk_field_set_values_device(self.h, C.c_void_p(d_free or 0), C.c_void_p(d_dirichlet or 0)))

    def field_get_values(self):
        fv = np.empty(self._n_free, dtype=np.float64)
        dv = np.empty(self._n_diri, dtype=np.float64)
        self._ck(self.lib.gtk_field_get_values(self.h, _ptr(fv), _ptr(dv)))
        return fv, dv

    def field_axpy_free(self, a: float, dx):
        dx = _f64(dx)
        if dx.size != self._n_free:
            raise ValueError(f"dx has {dx.size} entries, the space has {self._n_free} free dofs")
        self._ck(self.lib.gtk_field_axpy_free(self.h, float(a), _ptr(dx)))

    def space_dof_coordinates(self, M_at_nodes):
        """(x_free [n_free, D], x_dirichlet [n_dirichlet, D]): node_coordinates(space) seen through the dofs"""
        Mn = _f64(M_at_nodes)
        xf = np.empty((self._n_free, self._D), dtype=np.float64)
        xd = np.empty((self._n_diri, self._D), dtype=np.float64)
        self._ck(self.lib.gtk_space_dof_coordinates(self.h, _ptr(Mn), _ptr(xf), _ptr(xd)))
        return xf, xd

    def scalar_assemble(self, kind: int, **params) -> float:
        p, keep = make_params(**params)
        out = C.c_double(0.0)
        self._ck(self.lib.gtk_scalar_assemble(self.h, kind, C.byref(p), C.byref(out)))
        del keep
        return out.value

    def copy_nzval(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        nz = np.empty(self.nnz, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_copy_nzval(self.h, _ptr(nz)))
        return nz

    def copy_vector(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        b = np.empty(self.n_vec_rows, dtype=np.float64) if out is None else out
        self._ck(self.lib.gtk_copy_vector(self.h, _ptr(b)))
        return b

    def device_pointer(self, which: int):
        p = C.c_void_p()
        n = C.c_int64(0)
        self._ck(self.lib.gtk_device_pointer(self.h, which, C.byref(p), C.byref(n)))
        return p.value, n.value

    def info(self, key: int) -> int:
        return int(self.lib.gtk_info(self.h, key))

    def set_profiling(self, on: bool):
        self._ck(self.lib.gtk_set_profiling(self.h, 1 if on else 0))

    def profile(self):
        """[(kernel name, milliseconds)] of the last numeric call (needs set_profiling(True))."""
        out = []
        for i in range(self.lib.gtk_profile_count(self.h)):
            name = C.create_string_buffer(64)
            ms = C.c_double(0)
            self._ck(self.lib.gtk_profile_get(self.h, i, name, C.byref(ms)))
            out.append((name.value.decode(), ms.value))
        return out

    # -- multi-GPU ------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        lib = load_library()
        buf = C.create_string_buffer(128)
        rc = lib.gtk_comm_unique_id(buf)
        if rc != GTK_OK:
            raise GtkError(rc, "gtk_comm_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def comm_init(self, rank: int, n_ranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.gtk_comm_init(self.h, rank, n_ranks, buf))

    def comm_set_exchange(self, peer: int, send_nz, send_rows, recv_nz, recv_rows):
        sn = np.ascontiguousarray(send_nz, dtype=np.int64); sr = _i32(send_rows)
        rn = np.ascontiguousarray(recv_nz, dtype=np.int64); rr = _i32(recv_rows)
        self._ck(self.lib.gtk_comm_set_exchange(self.h, int(peer), sn.size, _ptr(sn), sr.size, _ptr(sr),
                                                rn.size, _ptr(rn), rr.size, _ptr(rr)))

    def comm_build_exchange(self, gid0: int, own_start):
        """device-side exchange plan for a block row partition (collective over the NCCL communicator)"""
        os_ = np.ascontiguousarray(own_start, dtype=np.int64)
        self._ck(self.lib.gtk_comm_build_exchange(self.h, int(gid0), _ptr(os_)))

    def comm_connect_peer_memory(self):
        self._ck(self.lib.gtk_comm_connect_peer_memory(self.h))

    def comm_ghost_info(self, key: int) -> int:
        return int(self.lib.gtk_comm_ghost_info(self.h, key))

    def comm_p2p_export(self, peer: int) -> bytes:
        """64-byte CUDA IPC handle of the block this rank receives `peer`'s ghost values in."""
        buf = C.create_string_buffer(64)
        self._ck(self.lib.gtk_comm_p2p_export(self.h, int(peer), buf))
        return buf.raw

    def comm_p2p_import(self, peer: int, handle: bytes):
        buf = C.create_string_buffer(bytes(handle), 64)
        self._ck(self.lib.gtk_comm_p2p_import(self.h, int(peer), buf))

    def comm_sum_ghost_rows(self):
        self._ck(self.lib.gtk_comm_sum_ghost_rows(self.h))

    def assemble_and_sum_ghost_rows_device(self, mform: int, mparams: dict, vform: int, vparams: dict):
        """assemble_matrix_and_vector_device + comm_sum_ghost_rows with the exchange overlapped with the sweep."""
        if self.n_vec_rows == 0:
            self.vector_symbolic(FREE)
        pm, k1 = make_params(**mparams)
        pv, k2 = make_params(**vparams)
        self._ck(self.lib.gtk_assemble_and_sum_ghost_rows_device(self.h, mform, C.byref(pm), vform, C.byref(pv)))
        del k1, k2
