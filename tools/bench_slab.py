"""Multi-GPU slab experiment (development tool): `cells` per GPU stacked in z over the ranks of a torchrun job, inputs
generated in HBM, overlapped assemble + ghost-row exchange.  Prints per-rank device time and, with GTK_COMM_TIMING=1, the
library's device timeline of one overlapped step.

    torchrun --nproc-per-node N tools/bench_slab.py --cells 512,512,64 [--steps 20]
"""
import argparse
import datetime
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", default="512,512,64")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import gtk_b200
    E, H = gtk_b200.engine, gtk_b200.hostprep
    P = importlib.import_module("galerkintoolkit_jl_b200.partition")
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr), timeout=datetime.timedelta(minutes=4))
    nx, ny, nz = (int(c) for c in a.cells.split(","))
    cells = (nx, ny, nz * world)
    dom = (0.0, 1.0, 0.0, float(ny) / nx, 0.0, float(nz * world) / nx)
    tab = H.measure_tabulation(H.lagrange_space(H.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2)), 1, "boundary"), 2)
    eng = E.Engine(lr)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    mp, vp = dict(alpha=1.0), dict(f_const=[1.0])
    tm = {}
    if world > 1:
        lay = P.slab_layout(cells, rank, world)
        P.attach_generated(eng, dom, cells, lay, tab, dist, tm)
        step = lambda: eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
    else:
        eng.set_cartesian_q1_problem(dom, cells)
        eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
        eng.matrix_symbolic(); eng.vector_symbolic()
        step = lambda: eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    eng.set_profiling(True)
    step(); torch.cuda.synchronize()
    prof = eng.profile()
    eng.set_profiling(False)
    os.environ["GTK_COMM_TIMING"] = "1"
    rec = {"rank": rank, "ms_per_step": ms, "setup_ms": {k: round(v, 2) for k, v in tm.items()}, "kernels_ms": [(n, round(t, 4)) for n, t in prof],
           "launches": eng.info(0)}
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, rec)
        if rank == 0:
            for r in out:
                print(json.dumps(r))
            print("max ms_per_step", max(r["ms_per_step"] for r in out))
        dist.barrier()
        dist.destroy_process_group()
    else:
        print(json.dumps(rec))
    eng.close()


if __name__ == "__main__":
    main()
