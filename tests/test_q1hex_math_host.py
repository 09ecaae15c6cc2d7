"""The fast Q1-hex kernel's element algebra (csrc/q1hex_math.cuh: lerped Jacobian columns, adjugate
Gram tensors, sum-factorised contraction) instantiated on the HOST and checked against the oracle's
literal per-point arithmetic.  CPU-only; the device instantiation is covered by the -m gpu tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import gt_oracle as O
import gtk_b200
from util import problem, tab_dict

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostcheck") / "libq1hex_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                           "-o", so, os.path.join(HERE, "hostcheck", "q1hex_host.cpp")])
    lib = C.CDLL(so)
    lib.q1hex_host.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    return lib


def _sym(r, c):
    r, c = min(r, c), max(r, c)
    return r * 8 - (r * (r - 1)) // 2 + (c - r)


@pytest.mark.parametrize("warp", [0.0, 0.25])
def test_sum_factorised_ke_matches_oracle(hostlib, warp):
    mesh, V, tab = problem((4, 3, 3), bc=None, warp=warp, domain=(0, 2, 0, 1, -1, 0.5))
    be_ref = O.element_matrices(O.LAPLACE, mesh.node_coordinates, mesh.cell_nodes, tab_dict(tab), alpha=0.75)
    bv_ref = O.element_vectors(O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, tab_dict(tab), f_const=[2.0])
    for cell in range(mesh.n_cells):
        X = np.ascontiguousarray(mesh.node_coordinates[mesh.cell_nodes[cell] - 1])
        Ke = np.zeros(36); be = np.zeros(8)
        hostlib.q1hex_host(X.ctypes.data, 0.75, 2.0, Ke.ctypes.data, be.ctypes.data)
        full = np.array([[Ke[_sym(r, c)] for c in range(8)] for r in range(8)])
        scale = np.abs(be_ref[cell]).max()
        assert np.abs(full - be_ref[cell]).max() <= 2e-14 * scale
        assert np.abs(be - bv_ref[cell]).max() <= 2e-14 * np.abs(bv_ref[cell]).max()


def test_internal_gauss_constants_match_the_tabulation_inputs():
    tab = gtk_b200.hostprep.measure_tabulation(gtk_b200.hostprep.lagrange_space(gtk_b200.hostprep.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2)), 1), 2)
    a = (1 - 1 / np.sqrt(3)) / 2
    assert np.allclose(tab.xq[0], [a, a, a], atol=1e-15) and np.allclose(tab.xq[1], [1 - a, a, a], atol=1e-15)
    assert np.allclose(tab.w, 0.125, atol=1e-16)
