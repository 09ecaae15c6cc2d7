"""Two-GPU test of the NCCL ghost-row path (needs >= 2 GPUs: run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, cells, dom, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=rank, world_size=world)
        sys.path[:0] = [os.path.dirname(os.path.abspath(__file__))]
        import gt_oracle as O
        import gtk_b200
        import importlib
        P = importlib.import_module("galerkintoolkit_jl_b200.partition")
        from util import problem, tab_dict
        from test_partition import check_owned_rows
        E = gtk_b200.engine
        torch.cuda.set_device(rank)
        mesh, V, tab = problem(cells, bc="boundary", domain=dom)
        tabd = tab_dict(tab)
        cp, rv, nz = O.assemble_matrix(O.LAPLACE, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd)
        bg = O.assemble_vector(O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd, f_const=[1.0])
        A_glob = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
        part = P.slab_problem(dom, cells, rank, world)
        eng = E.Engine(rank)
        colptr, rowval, n_owned = P.attach(eng, part, tab, dist)
        results = []
        for rep in range(3):
            eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
            fast = eng.info(5)
            eng.comm_sum_ghost_rows()
            results.append((eng.copy_nzval(), eng.copy_vector()))
        # overlapped variant (exchange hidden behind the sweep): bitwise the same as the two separate calls
        for rep in range(2):
            eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
            results.append((eng.copy_nzval(), eng.copy_vector()))
        for early in ("0", "1"):     # unpack after the whole sweep / concurrently with its middle (forced both ways)
            os.environ["GTK_EARLY_UNPACK"] = early
            eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
            results.append((eng.copy_nzval(), eng.copy_vector()))
        del os.environ["GTK_EARLY_UNPACK"]
        # the same two ways over NCCL instead of peer memory (both transports are set up by attach): bitwise equal
        assert eng.comm_ghost_info(3) == 1, "peer-memory transport not active"
        os.environ["GTK_DISABLE_P2P"] = "1"
        assert eng.comm_ghost_info(3) == 0
        eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        eng.comm_sum_ghost_rows()
        results.append((eng.copy_nzval(), eng.copy_vector()))
        eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        results.append((eng.copy_nzval(), eng.copy_vector()))
        del os.environ["GTK_DISABLE_P2P"]
        eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        results.append((eng.copy_nzval(), eng.copy_vector()))
        # the same exchange plan built ON THE DEVICE (gtk_comm_build_exchange + gtk_comm_connect_peer_memory): same counts,
        # same transport, bitwise the same owned rows
        eng2 = E.Engine(rank)
        cp2, rv2, n_owned2 = P.attach_device(eng2, part, tab, dist)
        assert np.array_equal(cp2, colptr) and np.array_equal(rv2, rowval) and n_owned2 == n_owned
        assert [eng2.comm_ghost_info(k) for k in range(4)] == [eng.comm_ghost_info(k) for k in range(4)]
        eng2.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        results.append((eng2.copy_nzval(), eng2.copy_vector()))
        eng2.assemble_matrix_and_vector_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        eng2.comm_sum_ghost_rows()
        results.append((eng2.copy_nzval(), eng2.copy_vector()))
        eng2.close()
        # communication-avoiding mode: the halo cell layer is assembled too, nothing is exchanged; the own rows are BITWISE
        # those of the single-GPU assembly of the whole mesh (computed here on this rank's GPU), for any number of ranks
        eng3 = E.Engine(rank)
        cp3, rv3, n_owned3 = P.attach_recompute(eng3, part, tab)
        assert np.array_equal(cp3, colptr) and np.array_equal(rv3, rowval) and n_owned3 == n_owned
        eng3.assemble_matrix_and_vector_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        nz3, b3 = eng3.copy_nzval(), eng3.copy_vector()
        eng3.close()
        check_owned_rows(part, colptr, rowval, nz3, b3, A_glob, bg)
        eng1 = E.Engine(rank)
        eng1.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
        eng1.set_space(V.cell_dofs, V.n_free, V.n_dirichlet)
        eng1.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
        eng1.matrix_symbolic(); eng1.vector_symbolic()
        cp1, rv1 = eng1.matrix_pattern()
        eng1.assemble_matrix_and_vector_device(E.FORM_LAPLACE, dict(alpha=1.0), E.FORM_SOURCE_CONST, dict(f_const=[1.0]))
        A1 = sp.csc_matrix((eng1.copy_nzval(), rv1.astype(np.int64) - 1, cp1.astype(np.int64) - 1), shape=(V.n_free, V.n_free)).tocsr()
        b1 = eng1.copy_vector()
        eng1.close()
        A3 = sp.csc_matrix((nz3, rv3.astype(np.int64) - 1, cp3.astype(np.int64) - 1), shape=(part.space.n_free, part.space.n_free)).tocsr()
        own_rows = np.flatnonzero(part.row_owner == part.rank)
        g = part.row_gid - 1                                   # local row / column -> global (0-based)
        for r in own_rows[:: max(1, own_rows.size // 400)]:
            lo, hi = A3.indptr[r], A3.indptr[r + 1]
            ref_row = A1.getrow(g[r])
            assert np.array_equal(g[A3.indices[lo:hi]], ref_row.indices)
            assert A3.data[lo:hi].tobytes() == ref_row.data.tobytes(), "own rows must be bitwise the single-GPU rows"
        assert b3[own_rows].tobytes() == b1[g[own_rows]].tobytes()
        nzval, b = results[0]
        check_owned_rows(part, colptr, rowval, nzval, b, A_glob, bg)
        own = part.row_owner == part.rank
        for nz_i, b_i in results[1:]:   # same GPU count => bitwise identical
            assert nz_i[own[rowval - 1]].tobytes() == nzval[own[rowval - 1]].tobytes() and b_i[own].tobytes() == b[own].tobytes()
        tot = torch.tensor([n_owned]); dist.all_reduce(tot)
        assert int(tot.item()) == rv.size
        eng.close()
        dist.barrier(); dist.destroy_process_group()
        q.put((rank, "ok", fast))
    except Exception:
        import traceback
        q.put((rank, traceback.format_exc(), -1))


@pytest.mark.parametrize("cells,world", [((6, 5, 8), 2), ((20, 12, 17), 2), ((33, 18, 60), 2), ((18, 11, 41), 4), ((12, 9, 43), 8)])
def test_multi_gpu_ghost_rows_match_single_domain(cells, world):
    """world = 4 has interior ranks that both send and receive (top, bottom and middle parts of the overlapped sweep)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, cells, (0, 1, 0, 1, 0, 2), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg, fast in results:
        assert msg == "ok", f"rank {rank}: {msg}"
        assert fast in (1, 2), "slab meshes must take a structured fast path"
