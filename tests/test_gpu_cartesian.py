"""gtk_set_cartesian_q1_problem: the benchmark inputs generated in HBM equal, bit for bit, the arrays the host-side
restatement of cartesian_mesh.jl:213-263 + space.jl:327-417 produces (hostprep / partition), for the whole mesh in the
reference numbering and for z-slabs in the partition's local numbering."""
import numpy as np
import pytest

import gtk_b200
from util import problem, tab_dict

pytestmark = pytest.mark.gpu
E = gtk_b200.engine
H = gtk_b200.hostprep


@pytest.mark.parametrize("cells,domain", [((4, 3, 5), (0, 1, 0, 1, 0, 1)), ((7, 9, 6), (-1, 2, 0.5, 1.75, 3, 3.3)),
                                          ((2, 2, 2), (0, 1, 0, 1, 0, 1)), ((33, 18, 21), (0, 1, 0, 0.6, 0, 0.7))])
def test_whole_mesh_equals_hostprep(cells, domain):
    mesh = H.cartesian_mesh(domain, cells)
    V = H.lagrange_space(mesh, 1, "boundary")
    eng = E.Engine(0)
    nf, nd = eng.set_cartesian_q1_problem(domain, cells)
    assert (nf, nd) == (V.n_free, V.n_dirichlet)
    assert eng.copy_device_array(8, np.float64).tobytes() == mesh.node_coordinates.tobytes()
    assert np.array_equal(eng.copy_device_array(9, np.int32).reshape(-1, 8), mesh.cell_nodes)
    assert np.array_equal(eng.copy_device_array(10, np.int32).reshape(-1, 8), V.cell_dofs)
    # and the assembly on the generated inputs is the assembly on the uploaded ones
    tab = H.measure_tabulation(V, 2)
    eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    eng.matrix_symbolic()
    cp, rv = eng.matrix_pattern()
    nz, b = eng.assemble_matrix_and_vector(E.FORM_LAPLACE, {}, E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    ref = E.Engine(0)
    ref.set_mesh(mesh.node_coordinates, mesh.cell_nodes); ref.set_space(V.cell_dofs, V.n_free, V.n_dirichlet)
    ref.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
    ref.matrix_symbolic()
    cp2, rv2 = ref.matrix_pattern()
    nz2, b2 = ref.assemble_matrix_and_vector(E.FORM_LAPLACE, {}, E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
    assert np.array_equal(cp, cp2) and np.array_equal(rv, rv2) and nz.tobytes() == nz2.tobytes() and b.tobytes() == b2.tobytes()
    eng.close(); ref.close()


@pytest.mark.parametrize("cells,world", [((5, 4, 9), 3), ((8, 6, 16), 4), ((3, 3, 2), 2)])
def test_slabs_equal_the_partition(cells, world):
    import importlib
    P = importlib.import_module("galerkintoolkit_jl_b200.partition")
    domain = (0, 1, 0, 2, -1, 1)
    for rank in range(world):
        part = P.slab_problem(domain, cells, rank, world)
        kc0 = part.k0 - 1 if rank > 0 else part.k0
        eng = E.Engine(0)
        nf, nd = eng.set_cartesian_q1_problem(domain, cells, kc0, part.k1, slab_local=True)
        assert (nf, nd) == (part.space.n_free, part.space.n_dirichlet)
        assert eng.copy_device_array(8, np.float64).tobytes() == part.mesh.node_coordinates.tobytes()
        assert np.array_equal(eng.copy_device_array(9, np.int32).reshape(-1, 8), part.mesh.cell_nodes)
        assert np.array_equal(eng.copy_device_array(10, np.int32).reshape(-1, 8), part.space.cell_dofs)
        eng.close()


def test_generator_rejects_bad_ranges():
    eng = E.Engine(0)
    with pytest.raises(E.GtkError):
        eng.set_cartesian_q1_problem((0, 1, 0, 1, 0, 1), (4, 4, 4), 1, 3, slab_local=False)
    with pytest.raises(E.GtkError):
        eng.set_cartesian_q1_problem((0, 1, 0, 1, 0, 1), (4, 4, 1))
    with pytest.raises(E.GtkError):
        eng.set_cartesian_q1_problem((0, 1, 0, 1, 0, 1), (4, 4, 4), 2, 5, slab_local=True)
    eng.close()
