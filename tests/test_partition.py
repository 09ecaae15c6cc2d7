"""Multi-GPU host logic on CPU: z-slab partition, ghost-row exchange plan, and the exchange itself over a
world_size-2 gloo process group.  The device data path (NCCL) is covered by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import gt_oracle as O
import gtk_b200
from util import problem, tab_dict

P = sys.modules["galerkintoolkit_jl_b200"].partition if hasattr(sys.modules.get("galerkintoolkit_jl_b200"), "partition") else None
if P is None:
    import importlib
    P = importlib.import_module("galerkintoolkit_jl_b200.partition")
H = gtk_b200.hostprep


def local_oracle(part, tab):
    """Oracle assembly of one rank's local matrix/vector: pattern from all local cells, values from active cells only."""
    m, V = part.mesh, part.space
    be = O.element_matrices(O.LAPLACE, m.node_coordinates, m.cell_nodes, tab)
    bv = O.element_vectors(O.SOURCE_CONST, m.node_coordinates, m.cell_nodes, tab, f_const=[1.0])
    f, c = part.active_cells
    inactive = np.ones(m.n_cells, dtype=bool); inactive[f:f + c] = False
    be[inactive] = 0.0; bv[inactive] = 0.0
    I, J, Vv = O.coo_matrix(be, V.cell_dofs, V.cell_dofs)
    colptr, rowval, nzval = O.sparse_csc(I, J, Vv, V.n_free, V.n_free)
    Ib, Vb = O.coo_vector(bv, V.cell_dofs)
    return colptr, rowval, nzval, O.dense_vector(Ib, Vb, V.n_free)


def check_owned_rows(part, colptr, rowval, nzval, b, A_glob, b_glob):
    n = part.space.n_free
    A = sp.csc_matrix((nzval, rowval.astype(np.int64) - 1, colptr.astype(np.int64) - 1), shape=(n, n)).tocsr()
    own = np.flatnonzero(part.row_owner == part.rank)
    g = part.row_gid - 1
    Ag = A_glob.tocsr()[g[own]]
    # embed local columns into global columns
    inj = sp.csr_matrix((np.ones(n), (np.arange(n), g)), shape=(n, A_glob.shape[1]))
    Aloc = (A[own] @ inj).tocsr()
    diff = abs(Aloc - Ag)
    scale = abs(A_glob).max()
    assert diff.max() <= 1e-12 * scale
    # structural completeness: every stored entry of the global row is stored locally
    Aloc.data[:] = 1; Ag = Ag.copy(); Ag.data[:] = 1
    assert (Ag - Aloc.multiply(Ag)).nnz == 0 or abs(Ag - Aloc.multiply(Ag)).max() == 0
    assert np.abs(b[own] - b_glob[g[own]]).max() <= 1e-12 * np.abs(b_glob).max()


@pytest.mark.parametrize("cells,world,bc", [((4, 3, 6), 2, "boundary"), ((5, 4, 9), 3, "boundary"), ((3, 3, 8), 4, [1, 2, 3, 4, 5, 6])])
def test_slab_partition_and_exchange_in_process(cells, world, bc):
    dom = (0, 1, 0, 2, 0, 3)
    mesh, V, tab = problem(cells, bc=bc, domain=dom)
    tabd = tab_dict(tab)
    cp, rv, nz = O.assemble_matrix(O.LAPLACE, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd)
    bg = O.assemble_vector(O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd, f_const=[1.0])
    A_glob = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
    parts = [P.slab_problem(dom, cells, r, world, bc) for r in range(world)]
    # the local problems tile the global one
    assert sum(p.active_cells[1] for p in parts) == mesh.n_cells
    assert sum(int((p.row_owner == p.rank).sum()) for p in parts) == V.n_free
    assert all(p.n_global_free == V.n_free for p in parts)
    for p in parts:   # local coordinates are the global ones
        f, c = p.active_cells
        k = p.k0 * cells[0] * cells[1]
        assert np.array_equal(p.mesh.node_coordinates[p.mesh.cell_nodes[f:f + c] - 1], mesh.node_coordinates[mesh.cell_nodes[k:k + c] - 1])
        assert np.all(np.diff(p.row_gid) > 0)
    local = [local_oracle(p, tabd) for p in parts]
    # in-process all-to-all
    outboxes = {}
    def fake_alltoall_factory(r):
        def fn(outbox):
            outboxes[r] = outbox
            return None
        return fn
    sends = [P.ghost_send_lists(p, local[r][0], local[r][1]) for r, p in enumerate(parts)]
    plans = []
    for r, p in enumerate(parts):
        def a2a(outbox, r=r):
            return [None if (src == r or r not in sends[src]) else (sends[src][r][1], sends[src][r][2], sends[src][r][4]) for src in range(world)]
        plans.append(P.build_exchange_plan(p, local[r][0], local[r][1], a2a))
    # exchange: add in increasing peer rank
    vals = [(l[2].copy(), l[3].copy()) for l in local]
    for r in range(world):
        for peer in sorted(plans[r]):
            pl = plans[r][peer]
            if pl["recv_nz"].size or pl["recv_rows"].size:
                snd = plans[peer][r]
                assert snd["send_nz"].size == pl["recv_nz"].size and snd["send_rows"].size == pl["recv_rows"].size
                vals[r][0][pl["recv_nz"]] += local[peer][2][snd["send_nz"]]
                vals[r][1][pl["recv_rows"]] += local[peer][3][snd["send_rows"]]
    for r, p in enumerate(parts):
        check_owned_rows(p, local[r][0], local[r][1], vals[r][0], vals[r][1], A_glob, bg)


@pytest.mark.parametrize("cells,world", [((4, 3, 6), 2), ((5, 4, 9), 3), ((3, 3, 8), 4)])
def test_recompute_mode_completes_the_own_rows_bitwise(cells, world):
    """Partition mode "recompute" (no exchange): a rank that ALSO assembles its halo cell layer holds, in the rows it owns, bitwise
    the rows of the single-domain matrix — every contributing cell is local and is visited in the same (increasing cell id) order"""
    dom = (0, 1, 0, 2, 0, 3)
    mesh, V, tab = problem(cells, bc="boundary", domain=dom)
    tabd = tab_dict(tab)
    cp, rv, nz = O.assemble_matrix(O.LAPLACE, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd)
    bg = O.assemble_vector(O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd, f_const=[1.0])
    A_glob = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free)).tocsr()
    for r in range(world):
        p = P.slab_problem(dom, cells, r, world)
        m, W = p.mesh, p.space
        lcp, lrv, lnz = O.assemble_matrix(O.LAPLACE, m.node_coordinates, m.cell_nodes, W.cell_dofs, W.n_free, W.n_dirichlet, tabd)   # ALL local cells
        lb = O.assemble_vector(O.SOURCE_CONST, m.node_coordinates, m.cell_nodes, W.cell_dofs, W.n_free, W.n_dirichlet, tabd, f_const=[1.0])
        A_loc = sp.csc_matrix((lnz, lrv.astype(np.int64) - 1, lcp.astype(np.int64) - 1), shape=(W.n_free, W.n_free)).tocsr()
        g = p.row_gid - 1
        own = np.flatnonzero(p.row_owner == p.rank)
        assert own.size > 0
        for i in own:
            lo, hi = A_loc.indptr[i], A_loc.indptr[i + 1]
            ref = A_glob.getrow(g[i])
            assert np.array_equal(g[A_loc.indices[lo:hi]], ref.indices)
            assert A_loc.data[lo:hi].tobytes() == ref.data.tobytes()
        assert lb[own].tobytes() == bg[g[own]].tobytes()


def _gloo_worker(rank, world, port, cells, dom, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import torch
        mesh, V, tab = problem(cells, bc="boundary", domain=dom)
        tabd = tab_dict(tab)
        cp, rv, nz = O.assemble_matrix(O.LAPLACE, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd)
        bg = O.assemble_vector(O.SOURCE_CONST, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, V.n_dirichlet, tabd, f_const=[1.0])
        A_glob = sp.csc_matrix((nz, rv.astype(np.int64) - 1, cp.astype(np.int64) - 1), shape=(V.n_free, V.n_free))
        part = P.slab_problem(dom, cells, rank, world)
        colptr, rowval, nzval, b = local_oracle(part, tabd)
        plan = P.build_exchange_plan(part, colptr, rowval, P.torch_alltoall_objects(dist))
        # the data path of comm.cu, restated with gloo: pack -> send/recv -> add in increasing peer rank
        reqs, recv_bufs = [], {}
        for peer in sorted(plan):
            pl = plan[peer]
            if pl["send_nz"].size + pl["send_rows"].size:
                buf = torch.from_numpy(np.concatenate([nzval[pl["send_nz"]], b[pl["send_rows"]]]))
                reqs.append(dist.isend(buf, peer))
            if pl["recv_nz"].size + pl["recv_rows"].size:
                recv_bufs[peer] = torch.empty(pl["recv_nz"].size + pl["recv_rows"].size, dtype=torch.float64)
                reqs.append(dist.irecv(recv_bufs[peer], peer))
        for rq in reqs:
            rq.wait()
        for peer in sorted(recv_bufs):
            pl = plan[peer]; buf = recv_bufs[peer].numpy()
            nzval[pl["recv_nz"]] += buf[:pl["recv_nz"].size]
            b[pl["recv_rows"]] += buf[pl["recv_nz"].size:]
        check_owned_rows(part, colptr, rowval, nzval, b, A_glob, bg)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))


def test_ghost_row_exchange_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, (4, 4, 6), (0, 1, 0, 1, 0, 1), q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg in results:
        assert msg == "ok", f"rank {rank}: {msg}"


class _FakeEngine:
    """Stands in for engine.Engine in attach(): pattern from the oracle, records the exchange plan and transport calls."""
    fail_import_on_rank = None

    def __init__(self, part, tabd):
        self.part, self.tabd = part, tabd
        self.exchanges, self.imported = {}, {}

    def set_mesh(self, *a): pass
    def set_space(self, *a): pass
    def set_tabulation(self, *a): pass
    def set_active_cells(self, *a): pass
    def vector_symbolic(self, *a): pass

    def matrix_symbolic(self):
        self.colptr, self.rowval, _, _ = local_oracle(self.part, self.tabd)
        return self.rowval.size

    def matrix_pattern(self):
        return self.colptr, self.rowval

    @staticmethod
    def comm_unique_id():
        return b"\x00" * 128

    def comm_init(self, rank, world, uid):
        assert len(uid) == 128

    def comm_set_exchange(self, peer, send_nz, send_rows, recv_nz, recv_rows):
        self.exchanges[peer] = (send_nz.size, send_rows.size, recv_nz.size, recv_rows.size)

    def comm_p2p_export(self, peer):
        return bytes([self.part.rank, peer]) + b"\x00" * 62

    def comm_p2p_import(self, peer, handle):
        if self.part.rank == _FakeEngine.fail_import_on_rank:
            raise RuntimeError("cudaIpcOpenMemHandle: peer access is not supported (simulated)")
        assert handle[:2] == bytes([peer, self.part.rank])       # the block the PEER keeps for this rank
        self.imported[peer] = handle


def _attach_worker(rank, world, port, fail_rank, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        os.environ.pop("GTK_DISABLE_P2P", None)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dom, cells = (0, 1, 0, 1, 0, 1), (4, 3, 9)
        _, _, tab = problem(cells, bc="boundary", domain=dom)
        part = P.slab_problem(dom, cells, rank, world)
        _FakeEngine.fail_import_on_rank = fail_rank
        eng = _FakeEngine(part, tab_dict(tab))
        P.attach(eng, part, tab, dist)
        peers = sorted(eng.exchanges)
        assert peers == [p for p in (rank - 1, rank + 1) if 0 <= p < world]
        disabled = os.environ.get("GTK_DISABLE_P2P") == "1"
        dist.barrier(); dist.destroy_process_group()
        q.put((rank, "ok", disabled, sorted(eng.imported)))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc(), None, None))


@pytest.mark.parametrize("fail_rank", [None, 1])
def test_attach_agrees_on_the_ghost_row_transport_over_gloo_world3(fail_rank):
    """attach(): every rank exports/imports the peer blocks; if ANY rank cannot map a peer's block, ALL ranks fall back
    to NCCL (GTK_DISABLE_P2P) — otherwise one side would push into memory the other side never polls."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 3
    port = 29100 + (os.getpid() % 2000) + (7 if fail_rank else 0)
    procs = [ctx.Process(target=_attach_worker, args=(r, world, port, fail_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, msg, disabled, imported in results:
        assert msg == "ok", f"rank {rank}: {msg}"
        assert disabled == (fail_rank is not None)
        if fail_rank is None:
            assert imported == [p for p in (rank - 1, rank + 1) if 0 <= p < world]
