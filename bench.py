#!/usr/bin/env python
"""bench.py — 3D Poisson CSC assembly throughput (nonzeros/s) on 1..8 B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n 128]

One "step" = one numeric assembly (GT.update_matrix!/update_vector! analogue, problems.jl:276-285,
352-361) of the 3D Q1 Poisson matrix AND right-hand side on BASELINE config 2 (128^3 hex cells,
full Dirichlet boundary, Float64/Int32): geometry, quadrature, element matrices, deterministic
scatter into CSC nzval and b.  The sparsity pattern (symbolic phase) is built once before the
timed region and reported separately as `symbolic_ms`.

  value     whole-job nnz/s with every input resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the public C ABI with HOST buffers: per step the node
            coordinates go host->device from pinned memory and nzval + b come back
  roofline  algorithmic bytes of the step / device time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline   the C restatement of the reference's CPU path (oracle/, "port") on this box

N > 1: one process per GPU (torchrun), the mesh is a stack of N z-slabs of 128^3 cells each
(weak scaling); every rank assembles its slab, ghost-row contributions of the slab interfaces
are summed into their owner over NCCL (send/recv between z-neighbours) inside the timed step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "3D Poisson Q1 CSC matrix+RHS assembly throughput (128^3 cells per GPU, numeric re-assembly)"
UNIT = "nnz/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + clock-event reasons DURING the timed region.  The timed region of this bench is milliseconds long,
    so `nvidia-smi -lms` (>= 100 ms period, B200_PROFILING.md recipe) cannot land a sample inside it: poll NVML
    (the library nvidia-smi itself queries) from a thread every ~0.5 ms instead; nvidia-smi is the fallback."""
    REASONS = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
               ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th, self.nv, self.h = index, [], False, None, None, None
        self.t_begin = self.t_end = None
        self.repeat = False
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((time.perf_counter(), sm, rs))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        if self.nv is None:
            return
        self.th = threading.Thread(target=self._poll, daemon=True)
        self.th.start()

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.nv is None:
            return self._smi_once()
        self.stop_flag = True
        self.th.join(timeout=2)
        nv = self.nv
        inside = [s for s in self.samples if self.t_begin is not None and self.t_begin <= s[0] <= self.t_end]
        use = inside if inside else self.samples
        sm = [s[1] for s in use]
        reasons = set()
        for _, _, rs in use:
            for name, attr in self.REASONS:
                if rs & getattr(nv, attr):
                    reasons.add(name)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(use), "samples_inside_timed_region": 0 if self.repeat else len(inside),
                "source": "nvml polling thread" + (" during an untimed 0.3 s repeat of the timed loop right after it (the timed "
                                                   "region itself is too short for NVML sampling)" if self.repeat else " during the timed region")}

    def _smi_once(self):
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                                  "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                    "reasons": [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")], "samples": 1,
                    "source": "nvidia-smi single query after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


def build_problem(n, rank=0, world=1, cells=None):
    """Config 2 inputs.  world>1: this rank's z-slab of the n x n x (n*world) mesh.
    cells = (nx, ny, nz): that many cells PER GPU instead of n^3 (e.g. 512,512,64 = one GPU's share of config 5)."""
    import gtk_b200
    H = gtk_b200.hostprep
    nx, ny, nz = cells if cells else (n, n, n)
    zmax = float(nz * world) / nx
    if world == 1:
        mesh = H.cartesian_mesh((0, 1, 0, float(ny) / nx, 0, zmax), (nx, ny, nz))
        V = H.lagrange_space(mesh, 1, "boundary")
        part = None
    else:
        from galerkintoolkit_jl_b200 import partition as P
        part = P.slab_problem((0, 1, 0, float(ny) / nx, 0, zmax), (nx, ny, nz * world), rank, world)
        mesh, V = part.mesh, part.space
    tab = H.measure_tabulation(V, 2)
    return mesh, V, tab, part


def algorithmic_bytes(mesh, V, nnz, n_rows):
    """BASELINE.md §3 / SURVEY.md §8d for a numeric-only step: coordinates + cell->nodes + cell->dofs read once,
    nzval and b written once (rowval/colptr are written once by the symbolic phase, outside the timed step)."""
    return 8 * mesh.D * mesh.n_nodes + 4 * mesh.n_lnodes * mesh.n_cells + 4 * V.n_ldofs * mesh.n_cells + 8 * nnz + 8 * n_rows


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (C port in oracle/; the Julia package cannot run here)
    on the host cores, same metric.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import c_oracle
    import gtk_b200
    H = gtk_b200.hostprep
    threads = os.cpu_count() or 1

    def one(n, nthreads):
        mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
        V = H.lagrange_space(mesh, 1, "boundary")
        tab = H.measure_tabulation(V, 2)
        tabd = dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)
        t = time.perf_counter()
        out = c_oracle.assemble(1, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tabd, nthreads=nthreads)
        return time.perf_counter() - t, out[1].size, out[4]

    t32, _, _ = one(32, threads)
    est = lambda n: t32 * (n / 32.0) ** 3
    budget = 150.0
    n = next((m for m in (128, 96, 64, 48, 32) if (args.steps + args.warmup) * est(m) <= budget), 32)
    for _ in range(args.warmup):
        one(n, threads)
    times, nnz, phases = [], 0, None
    for _ in range(args.steps):
        dt, nnz, phases = one(n, threads)
        times.append(dt)
    total = sum(times)
    value = nnz * args.steps / total
    sample = (f"{n}^3-cell Q1 Poisson matrix+RHS per step (count + cell loop + COO->CSC compress), "
              f"cell loop on {threads} pthreads, compress serial; phases(s) count/loop/compress/vector="
              f"{[round(float(x), 3) for x in phases]}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Poisson Q1 hex {n}^3 cells (bounded sample of config 2: 128^3), Float64/Int32, CPU",
                       "cells": n ** 3},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(n_full):
    """Timed C-oracle run on this box's host cores, bounded to ~10-30 s (rank 0, N=1 only)."""
    import c_oracle
    import gtk_b200
    H = gtk_b200.hostprep
    threads = os.cpu_count() or 1

    def one(n):
        mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
        V = H.lagrange_space(mesh, 1, "boundary")
        tab = H.measure_tabulation(V, 2)
        tabd = dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)
        t = time.perf_counter()
        out = c_oracle.assemble(1, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tabd, nthreads=threads)
        return time.perf_counter() - t, out[1].size

    t32, _ = one(32)
    n = next((m for m in (n_full, 96, 64, 48, 32) if m <= n_full and t32 * (m / 32.0) ** 3 <= 30.0), 32)
    dt, nnz = one(n)
    return {"value": nnz / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one full assembly (count + cell loop + COO->CSC + RHS) of {n}^3 cells, {dt:.2f} s; cell loop on "
                      f"{threads} pthreads, compress serial (the reference itself is single-threaded Julia)"}


def cpu_baseline_high_order(order=3, n=16):
    """The same C-oracle run for BASELINE config 3's element on a bounded sample (Q3 hexahedra, n^3 cells): the
    reference's generated loop costs n_q * n_ldofs^2 integrand evaluations per cell, which is what this times."""
    import c_oracle
    import gtk_b200
    H = gtk_b200.hostprep
    threads = os.cpu_count() or 1
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
    V = H.lagrange_space(mesh, order, "boundary")
    tab = H.measure_tabulation(V, 2 * order)
    tabd = dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)
    cap = int(mesh.n_cells) * V.cell_dofs.shape[1] ** 2
    t = time.perf_counter()
    out = c_oracle.assemble(1, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tabd, nthreads=threads, nnz_cap=cap)
    dt = time.perf_counter() - t
    return {"value": out[1].size / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one full assembly of Q{order} hexahedra on {n}^3 cells ({out[1].size} nnz), {dt:.2f} s; phases(s) "
                      f"count/loop/compress/vector={[round(float(x), 3) for x in out[4]]}; cell loop on {threads} pthreads, compress serial"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="cells per direction per GPU (128 = BASELINE config 2)")
    ap.add_argument("--cells", default=None, help="nx,ny,nz cells per GPU instead of n^3 (not a BASELINE bench line; e.g. 512,512,64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-high-order", action="store_true", help="skip the BASELINE config 3 (Q3 hex 64^3, DMMA path) entry")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import gtk_b200
    E = gtk_b200.engine
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    n = args.n
    cells = tuple(int(c) for c in args.cells.split(",")) if args.cells else None
    if cells:
        args.no_cpu_baseline = args.no_high_order = True
    mesh, V, tab, part = build_problem(n, rank, world, cells)
    eng = E.Engine(local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if world == 1:
        eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
        eng.set_space(V.cell_dofs, V.n_free, V.n_dirichlet)
        eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
        t0 = time.perf_counter()
        nnz_local = eng.matrix_symbolic()
        eng.vector_symbolic()
        nnz_owned = nnz_local
        n_owned_rows = V.n_free
    else:
        from galerkintoolkit_jl_b200 import partition as P
        _, rowval_h, nnz_owned = P.attach(eng, part, tab, dist)
        nnz_local = eng.nnz
        n_owned_rows = int(P.owned_rows_mask(part).sum())
    torch.cuda.synchronize()
    symbolic_first_ms = 1e3 * (time.perf_counter() - t0)   # includes lazy CUDA module load + first allocations
    symbolic_ms = symbolic_first_ms
    if world == 1:                                           # steady-state cost of the symbolic phase (pattern + plan)
        reps = []
        for _ in range(3):
            t0 = time.perf_counter()
            eng.matrix_symbolic()
            eng.vector_symbolic()
            torch.cuda.synchronize()
            reps.append(1e3 * (time.perf_counter() - t0))
        symbolic_ms = sorted(reps)[1]
    mp = dict(alpha=1.0)
    vp = dict(f_const=[1.0])

    def step():
        if world > 1:   # one call: sweep + NCCL ghost-row summation, the exchange overlapped with the sweep
            eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
        else:
            eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    launches_per_step = eng.info(0)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    sampler.mark_end()
    ms_total = ev0.elapsed_time(ev1)
    if rank == 0 and sampler.nv is not None:
        inside = sum(1 for smp in sampler.samples if sampler.t_begin <= smp[0] <= sampler.t_end)
        if inside < 5 and world == 1:
            # the timed region is a few ms: too short for NVML to land samples in it.  Repeat the same loop untimed for
            # ~0.3 s right away and sample the clocks of that (identical) load instead; stated in `clocks.source`.
            sampler.samples.clear()
            sampler.mark_begin()
            t_rep = time.perf_counter()
            while time.perf_counter() - t_rep < 0.3:
                for _ in range(20):
                    step()
                torch.cuda.synchronize()
            sampler.mark_end()
            sampler.repeat = True
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        tn = torch.tensor([nnz_owned], device="cuda", dtype=torch.int64)
        dist.all_reduce(tn)
        nnz_total = int(tn.item())
        td = torch.tensor([n_owned_rows], device="cuda", dtype=torch.int64)
        dist.all_reduce(td)
        dofs_total = int(td.item())
    else:
        nnz_total = nnz_owned
        dofs_total = n_owned_rows
    ms_per_step = ms_total / args.steps
    value = nnz_total / (ms_per_step * 1e-3)

    # per-kernel device times (events inside the lib around every launch), separate short loop
    eng.set_profiling(True)
    acc = {}
    for _ in range(5):
        step()
        torch.cuda.synchronize()
        for name, ms in eng.profile():
            acc.setdefault(name, []).append(ms)
    eng.set_profiling(False)
    kernels = {k: float(np.mean(v)) for k, v in acc.items()}
    dominant = max(kernels, key=kernels.get) if kernels else None

    # the same step on a NON-affine mesh (interior nodes displaced by 0.2 h U(-1,1)): exercises the general sweep
    # kernel instead of the exactly-affine one (reported next to the headline, never as the headline)
    general = None
    if world == 1 and not cells:
        rng = np.random.default_rng(0)
        warped = mesh.node_coordinates.copy()
        inner = ~gtk_b200.hostprep.boundary_node_mask(mesh)
        warped[inner] += 0.2 / n * rng.uniform(-1, 1, size=(int(inner.sum()), 3))
        eng.update_coordinates(warped)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gsteps = max(5, min(args.steps, 20))
        g0.record(stream)
        for _ in range(gsteps):
            step()
        g1.record(stream)
        torch.cuda.synchronize()
        gms = g0.elapsed_time(g1) / gsteps
        general = {"mesh": "same topology, interior nodes displaced by 0.2 h U(-1,1) (trilinear, non-affine cells)",
                   "ms_per_step": gms, "value": nnz_total / (gms * 1e-3), "unit": UNIT, "fast_path": eng.info(5),
                   "roofline_frac": algorithmic_bytes(mesh, V, nnz_local, V.n_free) / (gms * 1e-3) / 1e9 / measured_peak_gbs()[0]}
        eng.update_coordinates(mesh.node_coordinates)
        for _ in range(2):
            step()
        torch.cuda.synchronize()

    # end to end through the C ABI with pinned host buffers
    if nnz_local >= 2 ** 31 - 1:
        # beyond SparseMatrixCSC{Float64,Int32}: device-resident experiment only (--cells 512,512,512), no host copy
        if rank == 0:
            peak, peak_src = measured_peak_gbs()
            alg = algorithmic_bytes(mesh, V, nnz_local, V.n_free)
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "ms_per_step": ms_per_step, "config": {"workload": f"{cells} cells on one GPU, device-resident only (nnz exceeds Int32 colptr)",
                                                                     "nnz": int(nnz_local), "free_dofs": int(V.n_free)},
                              "symbolic_ms": symbolic_ms, "e2e": None, "kernels_ms": kernels,
                              "roofline": {"bound": "hbm", "achieved": alg / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": alg / (ms_per_step * 1e-3) / 1e9 / peak, "peak_source": peak_src}}), flush=True)
        eng.close()
        return
    xyz_pin = torch.from_numpy(mesh.node_coordinates).pin_memory()
    nz_pin = torch.empty(nnz_local, dtype=torch.float64).pin_memory()
    b_pin = torch.empty(V.n_free, dtype=torch.float64).pin_memory()
    xyz_np, nz_np, b_np = xyz_pin.numpy(), nz_pin.numpy(), b_pin.numpy()

    def e2e_step():
        eng.update_coordinates(xyz_np)                                # H2D
        step()
        eng.copy_nzval(nz_np)                                         # D2H
        eng.copy_vector(b_np)                                         # D2H

    e2e_steps = max(3, min(args.steps, 10))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    checksum = float(nz_np.sum() + b_np.sum())

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        alg = algorithmic_bytes(mesh, V, nnz_local, V.n_free)
        step_dev_ms = ms_per_step
        achieved = alg / (step_dev_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(dominant, {}).get("dram_bytes_per_launch") if (n == 128 and not cells) else None
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": (f"BASELINE config 2: 3D Poisson Q1 hex {n}^3 cells per GPU, full Dirichlet boundary, "
                                    f"Float64/Int32, matrix+RHS numeric assembly on a cached pattern") if not cells else
                                   (f"3D Poisson Q1 hex {cells[0]}x{cells[1]}x{cells[2]} cells per GPU (z-slab; 512x512x64 is one GPU's "
                                    f"share of BASELINE config 5), full Dirichlet boundary, matrix+RHS numeric assembly on a cached pattern"),
                       "cells_per_gpu": int(mesh.n_cells) if cells else n ** 3, "nnz_per_gpu": nnz_local, "free_dofs_per_gpu": V.n_free,
                       "partition": "none" if world == 1 else (f"{world} z-slabs of " + (f"{cells[0]}x{cells[1]}x{cells[2]}" if cells else f"{n}^3") +
                                                                  f" cells, ghost-row sum over " + ("peer memory (NVLink stores + flags)" if eng.comm_ghost_info(3) == 1 else "NCCL send/recv") +
                                                                  f", overlapped with the sweep ({eng.comm_ghost_info(2)} B/step on rank 0)"),
                       "l2": "per-step traffic (>0.6 GB) exceeds the 126 MB L2; no explicit flush",
                       "fast_path": eng.info(5)},
            "symbolic_ms": symbolic_ms, "symbolic_first_ms": symbolic_first_ms,
            "dofs_per_s": dofs_total / (ms_per_step * 1e-3),
            "e2e": {"value": nnz_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(xyz_np.nbytes),
                    "d2h_bytes_per_step": int(nz_np.nbytes + b_np.nbytes), "ms_per_step": 1e3 * e2e_s, "checksum": checksum},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_step": alg,
                         "bytes_per_nnz": alg / max(nnz_local, 1), "kernel": dominant,
                         "kernels_ms": kernels, "basis": "whole step device time (all kernels of the step)"},
            "clocks": clocks,
            "general_path": general,
        }
        if world == 1 and not args.no_high_order:
            # BASELINE config 3 next to the headline (never as the headline): Q3 hexahedra, element-matrix GEMM on the
            # FP64 tensor cores, roofline = measured DMMA issue rate
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_highorder
                eng.close()
                eng = None
                line["high_order"] = bench_highorder.run(64, 3, steps=5, warmup=2, device=local_rank, check=False)
                if not args.no_cpu_baseline:
                    line["high_order"]["cpu_baseline"] = cpu_baseline_high_order()
            except Exception as exc:   # reported, never hidden
                line["high_order"] = {"error": f"{type(exc).__name__}: {exc}"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n)
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
