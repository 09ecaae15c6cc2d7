"""High-order element-matrix GEMM on the FP64 tensor cores (csrc/elemgemm.cu, fast-path id 3) against the oracle and
against the generic per-entry kernel on identical inputs (BASELINE config 3 element: Q3 hex, 64 dofs, 64 points)."""
import os

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import assert_values_close, make_engine, oracle_matrix, problem

E = gtk_b200.engine
pytestmark = pytest.mark.gpu

CASES = [
    # cells, order, simplexify, bc, warp      -> template instance
    ((4, 3, 3), 3, False, "boundary", 0.15),   # Q3: 8 warps, 48 k-steps
    ((2, 2, 2), 3, False, None, 0.0),          # Q3, no Dirichlet, affine
    ((5, 4, 3), 2, False, [1, 4], 0.2),        # Q2: 27 dofs padded to 32, 27 points padded to 28
    ((4, 4, 3), 2, True, "boundary", 0.2),     # P2 tets: 10 dofs padded to 16, 11 points padded to 12
    ((5, 3, 4), 1, True, [2], 0.2),            # P1 tets
    ((7, 6, 5), 1, False, [1], 0.2),           # Q1 hex outside the structured fast path's (free x free, full BC) domain
]


@pytest.mark.parametrize("cells,order,simplexify,bc,warp", CASES)
def test_dmma_matches_oracle_and_generic(cells, order, simplexify, bc, warp):
    mesh, V, tab = problem(cells, order=order, bc=bc, simplexify=simplexify, warp=warp)
    colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, alpha=0.75)
    eng = make_engine(mesh, V, tab)
    os.environ["GTK_DISABLE_FASTPATH"] = "1"      # keep the structured Q1 sweep out of the way
    os.environ["GTK_DISABLE_Q1CELLS"] = "1"       # ... and the fused Q1 cell kernel of unstructured meshes (tests/test_gpu_unstructured.py)
    try:
        assert eng.matrix_symbolic() == rowval.size
        cp, rv = eng.matrix_pattern()
        assert np.array_equal(cp, colptr) and np.array_equal(rv, rowval)
        nz = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.75)
        assert eng.info(5) == 3, "DMMA path not taken"
        assert_values_close(nz, nzval)
        assert eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.75).tobytes() == nz.tobytes()   # bit-reproducible
        os.environ["GTK_DISABLE_DMMA"] = "1"
        nz_generic = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.75)
        assert eng.info(5) == 0
        assert_values_close(nz, nz_generic)
        os.environ.pop("GTK_DISABLE_DMMA", None)
        os.environ["GTK_ENABLE_DIRECT_WRITE"] = "1"      # same kernel, single-contribution entries written straight to nzval
        nz_direct = eng.matrix_numeric(E.FORM_LAPLACE, alpha=0.75)
        assert eng.info(5) == 3
        assert nz_direct.tobytes() == nz.tobytes()
    finally:
        os.environ.pop("GTK_DISABLE_DMMA", None)
        os.environ.pop("GTK_ENABLE_DIRECT_WRITE", None)
        os.environ.pop("GTK_DISABLE_FASTPATH", None)
        os.environ.pop("GTK_DISABLE_Q1CELLS", None)
        eng.close()


def test_dmma_active_cells_and_blocks():
    """Inactive cells (multi-GPU overlap layer) contribute exact zeros; free x Dirichlet block goes through the same path."""
    mesh, V, tab = problem((3, 3, 4), order=2, bc=[1, 6], warp=0.1)
    eng = make_engine(mesh, V, tab)
    for fd in [(E.FREE, E.DIRICHLET), (E.FREE, E.FREE)]:
        colptr, rowval, nzval = oracle_matrix(O.LAPLACE, mesh, V, tab, fd=fd)
        assert eng.matrix_symbolic(*fd) == rowval.size
        nz = eng.matrix_numeric(E.FORM_LAPLACE)
        assert eng.info(5) == 3
        assert_values_close(nz, nzval)
    eng.close()


def test_config3_full_size_invariants():
    """BASELINE config 3 at FULL size (Q3 hexahedra, 64^3 cells, 849 278 123 nonzeros) through the DMMA path: nnz of the
    tensor-product pattern, finite values, positive diagonal, zero column sums away from the boundary, structurally
    symmetric pattern and symmetric values (off-diagonal 8x8 tiles of an element matrix are mirrored bitwise; inside a
    diagonal tile both halves are accumulated by the tensor core: equal to rounding) — all checked on the device (tools/bench_highorder.py)."""
    import os
    import sys
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_highorder
    out = bench_highorder.run(64, 3, steps=1, warmup=1, check=True)
    c = out["checks"]
    assert out["fast_path"] == 3 and out["nnz"] == 849278123 and out["free_dofs"] == 6967871
    assert c["nnz_ok"] and c["n_free_ok"] and c["finite"] and c["diag_positive"] and c["zero_colsum_ok"], c
    assert c["pattern_symmetric"] and c["values_symmetric_relerr"] <= 1e-14, c
