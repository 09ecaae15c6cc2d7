#!/usr/bin/env python
"""bench.py — 3D Poisson CSC assembly throughput (nonzeros/s) on 1..8 B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--n 128]

One "step" = one numeric RE-ASSEMBLY (GT.update_matrix!/update_vector!, problems.jl:276-285, 352-361) of the 3D Q1
Poisson matrix AND right-hand side of BASELINE config 2 (128^3 hex cells, full Dirichlet boundary, Float64/Int32) on a
cached pattern: geometry, quadrature, element matrices, deterministic scatter into CSC nzval and b.  Both arms time
THIS step (`--impl reference`: the C port of the reference's loop + sparse_matrix!(A,V,cache) through the cached nz
index, assembly.jl:584-588), and both also report the FIRST assembly (symbolic + numeric) next to it.

  value           whole-job nnz/s of the re-assembly with every input resident in HBM (CUDA events, max over ranks)
  e2e             the same step through the public C ABI with HOST buffers: per step the node coordinates go
                  host->device from pinned memory and nzval + b come back
  reassembly / first_assembly   {device_ms, e2e_ms, value, e2e_value} of update_matrix! / of assemble_matrix
  roofline        algorithmic bytes of the step / device time, against MEASURED_PEAKS.json hbm_gbs
  general_path    same mesh with non-affine cells (general sweep kernel)
  mixed_path      same mesh with ONE displaced node (affine kernel + general kernel on the affected tiles)
  unstructured_path   same mesh with the cells in a random order (what a Gmsh mesh hits: element kernel + staged reduction)
  high_order      BASELINE config 3 (Q3 hexahedra 64^3, FP64 tensor cores)
  elasticity      BASELINE config 4 element (P2 x 3 on tetrahedra, Strang degree-4 rule) at 64^3 x 6 tetrahedra
  multifield      SURVEY §8 f4: Stokes (Q2 x 3 x Q1 product space, 24^3 hexahedra) and ∫_Λ jump(u) jump(v) over the interior faces of 48^3 cells
  config5         BASELINE config 5 at THIS GPU count: 512 x 512 x (512/N) cells per GPU, T_N, T_1 (rank 0 alone, device
                  resident) and the strong-scaling efficiency T_1 / (N T_N)
  cpu_baseline    the C restatement of the reference's CPU path (oracle/, "port") on this box

N > 1: one process per GPU (torchrun); the headline mesh is a stack of N z-slabs of 128^3 cells (weak scaling), generated
in HBM (gtk_set_cartesian_q1_problem).  A rank's own rows are completed either by also assembling the lower neighbour's top
cell layer (GTK_PARTITION_MODE=recompute, default: no data-path exchange, own rows bitwise the single-GPU rows) or by summing
ghost-row contributions into their owner over NVLink peer memory / NCCL inside the timed step (=exchange).  config5 reports both.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "3D Poisson Q1 CSC matrix+RHS assembly throughput (128^3 cells per GPU, numeric re-assembly on a cached pattern)"
UNIT = "nnz/s"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def config_dict(n):
    """identical on both arms (the driver compares them): everything here follows from n alone"""
    m = n - 1
    return {"workload": f"BASELINE config 2: 3D Poisson Q1 hex {n}^3 cells per GPU on GT.cartesian_mesh, full Dirichlet boundary, "
                        f"Float64/Int32, matrix+RHS",
            "step": "numeric re-assembly on a cached pattern (update_matrix! + update_vector!, problems.jl:276-285, 352-361)",
            "cells_per_gpu": n ** 3, "nnz_per_gpu": (3 * m - 2) ** 3, "free_dofs_per_gpu": m ** 3,
            "l2": "per-step traffic (> 0.6 GB) exceeds the 126 MB L2; no explicit flush"}


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
            self.h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(index))
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


def physical_index(index):
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    if visible:
        parts = visible.split(",")
        if index < len(parts) and parts[index].strip().isdigit():
            return int(parts[index])
    return index


def bind_to_gpu_numa_node(local_rank):
    """Pin this rank (and, by first touch, its pinned host buffers) to the CPU cores NVML reports as local to its GPU:
    with 8 ranks each copying ~0.45 GB per step to the host, buffers that all live on one socket cap the aggregate
    device-to-host rate far below 8 PCIe links.  Returns a description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index(local_rank))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1 and 64 * w + b < ncpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cores local to the GPU (NVML cpu affinity), first {allowed[0]}"
        return "NVML affinity empty: unchanged"
    except Exception as exc:      # noqa: BLE001 - reported in the line
        return f"unchanged ({type(exc).__name__})"


def algorithmic_bytes(n_nodes, n_cells, nnz, n_rows, n_lnodes=8, n_ldofs=8, D=3):
    """BASELINE.md §3 / SURVEY.md §8d for a numeric-only step: coordinates + cell->nodes + cell->dofs read once,
    nzval and b written once (rowval/colptr are written once by the symbolic phase, outside the timed step)."""
    return 8 * D * n_nodes + 4 * n_lnodes * n_cells + 4 * n_ldofs * n_cells + 8 * nnz + 8 * n_rows


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the C port of the reference's CPU path (oracle/)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reassembly(n, steps, warmup, threads):
    """-> dict: the reference's update_matrix!/update_vector! step and its first assembly, timed on the host cores"""
    import c_oracle
    import gtk_b200
    H = gtk_b200.hostprep
    mesh = H.cartesian_mesh((0, 1, 0, 1, 0, 1), (n, n, n))
    V = H.lagrange_space(mesh, 1, "boundary")
    tab = H.measure_tabulation(V, 2)
    tabd = dict(w=tab.w, N=tab.N, dN=tab.dN, M=tab.M, dM=tab.dM)
    t = time.perf_counter()
    R = c_oracle.Reassembly(1, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tabd, nthreads=threads)
    first_s = time.perf_counter() - t
    first_ph = [float(x) for x in R.first_phases]
    for _ in range(warmup):
        R.step()
    times, ph = [], np.zeros(3)
    for _ in range(steps):
        t = time.perf_counter()
        R.step()
        times.append(time.perf_counter() - t)
        ph += R.phases
    nnz = int(R.nzval.size)
    s = sum(times) / len(times)
    return {"n": n, "nnz": nnz, "reassembly_s": s, "reassembly_phases_s": [round(float(x), 4) for x in ph / steps],
            "first_assembly_s": first_s, "first_assembly_phases_s": [round(x, 4) for x in first_ph], "threads": threads}


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (C port in oracle/; the Julia package cannot run here) on the host
    cores, SAME step (re-assembly on a cached pattern), same metric/config.  Rank 0 only; the others exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    n = args.n
    r = cpu_reassembly(n, args.steps, args.warmup, threads)
    value = r["nnz"] / r["reassembly_s"]
    sample = (f"{n}^3-cell Q1 Poisson matrix+RHS re-assembly per step = cell loop + contribute! on {threads} pthreads, then "
              f"sparse_matrix!(A,V,cache) and dense_vector! serial (the reference is single-threaded Julia); phases(s) "
              f"loop/compress!/vector = {r['reassembly_phases_s']}; first assembly {r['first_assembly_s']:.2f} s, phases(s) "
              f"count/loop/compress/vector = {r['first_assembly_phases_s']}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * r["reassembly_s"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(n),
            "reassembly": {"device_ms": None, "e2e_ms": 1e3 * r["reassembly_s"], "value": value, "e2e_value": value},
            "first_assembly": {"device_ms": None, "e2e_ms": 1e3 * r["first_assembly_s"], "value": r["nnz"] / r["first_assembly_s"],
                               "e2e_value": r["nnz"] / r["first_assembly_s"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cpu_baseline(n_full):
    """Bounded (~10-30 s) C-oracle run on this box's host cores (rank 0, N=1 only): re-assembly step + first assembly."""
    threads = os.cpu_count() or 1
    r = cpu_reassembly(n_full, 5, 1, threads)
    return {"value": r["nnz"] / r["reassembly_s"], "unit": UNIT, "cores": threads, "kind": "port",
            "first_assembly_value": r["nnz"] / r["first_assembly_s"],
            "sample": f"5 re-assembly steps of {n_full}^3 cells ({r['reassembly_s']:.3f} s each: loop/compress!/vector = "
                      f"{r['reassembly_phases_s']}) after one first assembly ({r['first_assembly_s']:.2f} s: count/loop/compress/"
                      f"vector = {r['first_assembly_phases_s']}); cell loops on {threads} pthreads, compression serial (the "
                      f"reference itself is single-threaded Julia)"}


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
    R = c_oracle.Reassembly(1, mesh.node_coordinates, mesh.cell_nodes, V.cell_dofs, V.n_free, tabd, nthreads=threads)
    t = time.perf_counter()
    R.step()
    dt = time.perf_counter() - t
    return {"value": R.nzval.size / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one re-assembly of Q{order} hexahedra on {n}^3 cells ({R.nzval.size} nnz), {dt:.2f} s; phases(s) "
                      f"loop/compress!/vector={[round(float(x), 3) for x in R.phases]}; cell loop on {threads} pthreads, compression serial"}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
SPIN_MS = float(os.environ.get("GTK_BENCH_SPIN_MS", "0"))   # optional untimed device work right before every timed region (default off:
# measured on B200, a 60 ms spin-up changes nothing — 0.1254 vs 0.1260 ms per step — so the K timed steps follow the W warm-up steps directly)


def spin_count(est_ms_per_step):
    """steps that keep the GPU busy for ~SPIN_MS: the SM clock and the memory system are in their sustained state when the
    timed region starts (a 20-step region of a 0.12 ms kernel is 2.5 ms long — shorter than the clock ramp after an idle gap)"""
    return 0 if SPIN_MS <= 0 else max(1, int(SPIN_MS / max(est_ms_per_step, 1e-3)))


def timed_loop(torch, stream, fn, steps, barrier=None, spin=0):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(spin):          # untimed; same count on every rank (the fused exchange runs in lockstep)
        fn()
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(steps):
        fn()
    ev1.record(stream)
    if barrier:
        barrier()
    else:
        torch.cuda.synchronize()
    return ev0.elapsed_time(ev1) / steps


PARTITION_MODE = os.environ.get("GTK_PARTITION_MODE", "recompute")   # "recompute" (communication-avoiding) | "exchange" (ghost-row sum)
MODE_TEXT = {
    "recompute": "every rank also assembles the lower neighbour's top cell layer (it holds those cells and their node coordinates anyway), so "
                 "all cells that touch its own rows are local: NO data-path exchange; own rows bitwise equal to the single-GPU matrix",
    "exchange": "every cell assembled once; ghost-row partial sums summed into their owner over NVLink peer memory (NCCL fallback) inside the "
                "sweep kernel's copy-out; exchange plan built on the device",
}


def run_config5(torch, dist, E, P, tab, rank, world, local_rank, stream, steps=10, mode=None, t1_known=None):
    """BASELINE config 5 at this GPU count: 512^3 cells as `world` z-slabs, all inputs generated in HBM.  T_1 is measured
    in the same job by rank 0 alone on the whole mesh (device resident: nnz exceeds the Int32 colptr of the ABI's copy-out)."""
    cells = (512, 512, 512)
    dom = (0, 1, 0, 1, 0, 1)
    mp, vp = dict(alpha=1.0), dict(f_const=[1.0])
    out = {"workload": "BASELINE config 5: 3D Poisson Q1 hex 512^3 cells, full Dirichlet boundary, matrix+RHS numeric re-assembly, "
                       f"{world} z-slab(s) of 512x512x{512 // world} cells", "n_gpus": world}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # T_N
    eng = E.Engine(local_rank)
    eng.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    if world == 1:
        nf, _ = eng.set_cartesian_q1_problem(dom, cells)
        eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        nnz_owned = eng.matrix_symbolic()
        eng.vector_symbolic()
        step = lambda: eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
    else:
        lay = P.slab_layout(cells, rank, world)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        tm = {}
        mode = mode or PARTITION_MODE
        nnz_owned = P.attach_generated(eng, dom, cells, lay, tab, dist, tm, mode=mode)
        out["setup_phases_ms_rank0"] = {k: round(v, 3) for k, v in tm.items()}
        out["partition_mode"] = mode
        out["partition"] = MODE_TEXT[mode]
        if mode == "exchange":
            step = lambda: eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
        else:
            step = lambda: eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
    torch.cuda.synchronize()
    sym_ms = 1e3 * (time.perf_counter() - t1)
    if world > 1:   # the NCCL communicator is created once per process in an application: not part of the symbolic phase
        sym_ms = tm["symbolic"] + tm.get("exchange_plan", 0.0) + tm.get("peer_memory", 0.0)
    for _ in range(3):
        step()
    ms = timed_loop(torch, stream, step, steps, barrier, spin=spin_count(8.0 / world))
    tot = torch.tensor([float(nnz_owned), ms, sym_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
        nnz_total, ms, sym_ms = int(tot[0].item()), float(mx[1].item()), float(mx[2].item())
        if mode == "exchange":
            transport = "peer memory (NVLink stores + flags)" if eng.comm_ghost_info(3) == 1 else "NCCL send/recv"
            out["exchange"] = {"transport": transport, "bytes_per_step_rank0": eng.comm_ghost_info(2), "plan": "built on the device"}
    else:
        nnz_total = int(nnz_owned)
    out.update({"nnz": nnz_total, "ms_per_step": ms, "value": nnz_total / (ms * 1e-3), "unit": UNIT,
                "symbolic_ms": sym_ms, "symbolic_includes": "pattern + sweep plan" +
                ("" if world == 1 or mode != "exchange" else " + exchange plan (device) + peer-memory handles; NCCL communicator creation and input generation listed in setup_phases_ms_rank0")})
    eng.close()
    del eng
    torch.cuda.synchronize()
    # T_1 on rank 0 (the same measurement when world == 1)
    t1_ms = ms
    if world > 1 and t1_known is not None:
        t1_ms = t1_known
    elif world > 1:
        t1_ms = 0.0
        if rank == 0:
            e1 = E.Engine(local_rank)
            e1.set_stream(stream.cuda_stream)
            e1.set_cartesian_q1_problem(dom, cells)
            e1.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
            e1.matrix_symbolic()
            e1.vector_symbolic()
            s1 = lambda: e1.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
            for _ in range(3):
                s1()
            t1_ms = timed_loop(torch, stream, s1, steps, spin=spin_count(8.0))
            e1.close()
        t = torch.tensor([t1_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t1_ms = float(t.item())
    peak, _ = measured_peak_gbs()
    alg1 = algorithmic_bytes(513 ** 3, 512 ** 3, nnz_total, 511 ** 3)
    out.update({"t1_ms": t1_ms, "t1": "the whole 512^3 mesh on ONE GPU, device resident, same job (rank 0)",
                "strong_efficiency": t1_ms / (world * ms),
                "t1_roofline_frac": alg1 / (t1_ms * 1e-3) / 1e9 / peak})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="cells per direction per GPU (128 = BASELINE config 2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-high-order", action="store_true", help="skip the BASELINE config 3 (Q3 hex 64^3, DMMA path) entry")
    ap.add_argument("--no-config5", action="store_true", help="skip the BASELINE config 5 (512^3 over the N GPUs) entry")
    ap.add_argument("--no-elasticity", action="store_true", help="skip the BASELINE config 4 (P2 x 3 elasticity on tetrahedra) entry")
    ap.add_argument("--no-multifield", action="store_true", help="skip the Stokes (product space) and skeleton-integral entries (SURVEY §8 f4)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no general/unstructured/high-order/config5/cpu entries")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    if args.no_extras:
        args.no_cpu_baseline = args.no_high_order = args.no_config5 = args.no_elasticity = args.no_multifield = True

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    affinity = bind_to_gpu_numa_node(local_rank)
    import torch
    import gtk_b200
    import importlib
    E, H = gtk_b200.engine, gtk_b200.hostprep
    P = importlib.import_module("galerkintoolkit_jl_b200.partition")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=4))

    n = args.n
    cfg = config_dict(n)
    stream = torch.cuda.current_stream()
    mp, vp = dict(alpha=1.0), dict(f_const=[1.0])
    dom = (0.0, 1.0, 0.0, 1.0, 0.0, float(world))
    cells_total = (n, n, n * world)
    tab = H.measure_tabulation(H.lagrange_space(H.cartesian_mesh((0, 1, 0, 1, 0, 1), (2, 2, 2)), 1, "boundary"), 2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = E.Engine(local_rank)
    eng.set_stream(stream.cuda_stream)
    torch.cuda.synchronize()
    mesh = V = None
    t0 = time.perf_counter()
    if world == 1:
        mesh = H.cartesian_mesh(dom, cells_total)
        V = H.lagrange_space(mesh, 1, "boundary")
        host_prep_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        eng.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
        eng.set_space(V.cell_dofs, V.n_free, V.n_dirichlet)
        eng.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
        torch.cuda.synchronize()
        upload_ms = 1e3 * (time.perf_counter() - t0)
        t0 = time.perf_counter()
        nnz_local = eng.matrix_symbolic()
        eng.vector_symbolic()
        nnz_owned, n_owned_rows, n_free_local = nnz_local, V.n_free, V.n_free
        n_nodes_local, n_cells_local = mesh.n_nodes, mesh.n_cells
    else:
        lay = P.slab_layout(cells_total, rank, world)
        host_prep_s, upload_ms = time.perf_counter() - t0, 0.0
        t0 = time.perf_counter()
        setup_tm = {}
        nnz_owned = P.attach_generated(eng, dom, cells_total, lay, tab, dist, setup_tm, mode=PARTITION_MODE)
        nnz_local, n_free_local, n_owned_rows = eng.nnz, lay.n_free, lay.own_hi - lay.own_lo
        n_nodes_local, n_cells_local = (n + 1) ** 2 * (lay.k1 - lay.kc0 + 1), n * n * (lay.k1 - lay.kc0)
    torch.cuda.synchronize()
    symbolic_first_ms = 1e3 * (time.perf_counter() - t0)   # includes lazy CUDA module load + first allocations (+ NCCL init for N > 1)
    symbolic_ms = symbolic_first_ms
    if world > 1:
        symbolic_ms = setup_tm["symbolic"] + setup_tm.get("exchange_plan", 0.0) + setup_tm.get("peer_memory", 0.0)

    def step():
        if world > 1 and PARTITION_MODE == "exchange":   # one call: sweep + ghost-row summation, the exchange inside the sweep
            eng.assemble_and_sum_ghost_rows_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
        else:
            eng.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)

    first_dev_ms = None
    if world == 1:                                           # steady-state cost of the symbolic phase / of a first assembly
        reps, firsts = [], []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.matrix_symbolic()
            eng.vector_symbolic()
            torch.cuda.synchronize()
            reps.append(1e3 * (time.perf_counter() - t0))
            step()
            torch.cuda.synchronize()
            firsts.append(1e3 * (time.perf_counter() - t0))
        symbolic_ms, first_dev_ms = sorted(reps)[1], sorted(firsts)[1]

    for _ in range(args.warmup):
        step()
    launches_per_step = eng.info(0)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    sampler.mark_begin()
    ms_per_step = timed_loop(torch, stream, step, args.steps, barrier, spin=spin_count(0.15))
    sampler.mark_end()
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
        t = torch.tensor([ms_per_step, symbolic_first_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t2 = torch.tensor([symbolic_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_per_step, symbolic_first_ms, symbolic_ms = float(t[0].item()), float(t[1].item()), float(t2.item())
        tn = torch.tensor([nnz_owned, n_owned_rows], device="cuda", dtype=torch.int64)
        dist.all_reduce(tn)
        nnz_total, dofs_total = int(tn[0].item()), int(tn[1].item())
    else:
        nnz_total, dofs_total = nnz_owned, n_owned_rows
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
    peak, peak_src = measured_peak_gbs()
    alg = algorithmic_bytes(n_nodes_local, n_cells_local, nnz_local, n_free_local)

    # ---- end to end through the C ABI with pinned host buffers (first touch after the affinity binding) ----
    if world == 1:
        xyz_host = mesh.node_coordinates
    else:
        xyz_host = eng.copy_device_array(8, np.float64).reshape(-1, 3)
    xyz_pin = torch.from_numpy(np.ascontiguousarray(xyz_host)).pin_memory()
    nz_pin = torch.empty(nnz_local, dtype=torch.float64).pin_memory()
    b_pin = torch.empty(n_free_local, dtype=torch.float64).pin_memory()
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

    # ---- first assembly end to end (N = 1): inputs H2D, symbolic, numeric, pattern + values D2H ----
    first = None
    if world == 1:
        cp_pin = np.empty(V.n_free + 1, dtype=np.int32)
        rv_pin = torch.empty(nnz_local, dtype=torch.int32).pin_memory().numpy()
        reps = []
        for _ in range(3):
            e2 = E.Engine(local_rank)
            e2.set_stream(stream.cuda_stream)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2.set_mesh(mesh.node_coordinates, mesh.cell_nodes)
            e2.set_space(V.cell_dofs, V.n_free, V.n_dirichlet)
            e2.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
            e2.matrix_symbolic()
            e2.vector_symbolic()
            e2.lib.gtk_matrix_pattern(e2.h, cp_pin.ctypes.data, rv_pin.ctypes.data)
            e2.assemble_matrix_and_vector(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp, nzval=nz_np, b=b_np)
            reps.append(1e3 * (time.perf_counter() - t0))
            e2.close()
        first = {"device_ms": first_dev_ms, "e2e_ms": sorted(reps)[1], "value": nnz_total / (first_dev_ms * 1e-3),
                 "e2e_value": nnz_total / (sorted(reps)[1] * 1e-3),
                 "what": "device_ms: symbolic (pattern + plan) + numeric with the inputs already in HBM, warm; e2e_ms: a fresh "
                         "context — mesh/space/tabulation host->device, symbolic, numeric, colptr/rowval/nzval/b device->host"}

    # ---- the same step on a NON-affine mesh and on an UNSTRUCTURED (cell-permuted) mesh (N = 1) ----
    general = unstructured = mixed = None
    if world == 1:
        rng = np.random.default_rng(0)
        warped = mesh.node_coordinates.copy()
        inner = ~H.boundary_node_mask(mesh)
        warped[inner] += 0.2 / n * rng.uniform(-1, 1, size=(int(inner.sum()), 3))
        eng.update_coordinates(warped)
        for _ in range(3):
            step()
        gsteps = max(5, min(args.steps, 20))
        gms = timed_loop(torch, stream, step, gsteps, spin=spin_count(0.4))
        general = {"mesh": "same topology, interior nodes displaced by 0.2 h U(-1,1) (trilinear, non-affine cells)",
                   "ms_per_step": gms, "value": nnz_total / (gms * 1e-3), "unit": UNIT, "fast_path": eng.info(5),
                   "roofline_frac": alg / (gms * 1e-3) / 1e9 / peak}
        # ONE displaced node: the affine kernel everywhere + the general kernel on the tiles around the 8 non-affine cells
        one = mesh.node_coordinates.copy()
        one[int(np.flatnonzero(inner)[inner.sum() // 2])] += 0.2 / n
        eng.update_coordinates(one)
        for _ in range(3):
            step()
        mms = timed_loop(torch, stream, step, gsteps, spin=spin_count(0.2))
        mixed = {"mesh": "config 2 with ONE interior node displaced by 0.2 h (8 non-affine cells)", "ms_per_step": mms,
                 "value": nnz_total / (mms * 1e-3), "unit": UNIT, "fast_path": eng.info(5), "roofline_frac": alg / (mms * 1e-3) / 1e9 / peak}
        eng.update_coordinates(mesh.node_coordinates)
        if not args.no_extras:
            try:
                perm = rng.permutation(mesh.n_cells)
                eu = E.Engine(local_rank)
                eu.set_stream(stream.cuda_stream)
                eu.set_mesh(mesh.node_coordinates, mesh.cell_nodes[perm])
                eu.set_space(V.cell_dofs[perm], V.n_free, V.n_dirichlet)
                eu.set_tabulation(tab.w, tab.N, tab.dN, tab.M, tab.dM)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                eu.matrix_symbolic()
                eu.vector_symbolic()
                ustep = lambda: eu.assemble_matrix_and_vector_device(E.FORM_LAPLACE, mp, E.FORM_SOURCE_CONST, vp)
                ustep()
                torch.cuda.synchronize()
                usym_ms = 1e3 * (time.perf_counter() - t0)
                for _ in range(2):
                    ustep()
                ums = timed_loop(torch, stream, ustep, 5, spin=spin_count(1.3))
                eu.set_profiling(True)
                ustep()
                torch.cuda.synchronize()
                uk = {k: v for k, v in eu.profile()}
                n_coo = eu.info(4)
                # what this path moves by construction: element matrices + vectors staged once (write + read), the 4-byte
                # slot index of the reduction plan read once, on top of the compulsory bytes
                staged = 2 * 8 * (64 + 8) * mesh.n_cells + 4 * n_coo
                unstructured = {"mesh": "config 2 with the cells in a random order (no lattice structure to exploit: what any Gmsh mesh hits)",
                                "ms_per_step": ums, "value": nnz_total / (ums * 1e-3), "unit": UNIT, "fast_path": eu.info(5),
                                "kernels_ms": uk, "symbolic_plus_first_numeric_ms": usym_ms,
                                "roofline": {"bound": "hbm", "achieved": alg / (ums * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                             "frac": alg / (ums * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_step": alg,
                                             "bytes_moved_by_construction": alg + staged,
                                             "traffic_ratio_by_construction": (alg + staged) / alg}}
                eu.close()
            except Exception as exc:   # reported, never hidden
                unstructured = {"error": f"{type(exc).__name__}: {exc}"}

    line = None
    if rank == 0:
        traffic, traffic_src = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                rec = json.load(f).get(dominant, {}) if n == 128 else {}
                traffic = rec.get("dram_bytes_per_launch")
                traffic_src = rec.get("source", "committed ncu --set full capture of this kernel on this workload (profiles/), NOT measured in this run")
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "partition": "none" if world == 1 else (f"{world} z-slabs of {n}^3 cells generated in HBM; mode '{PARTITION_MODE}': " + MODE_TEXT[PARTITION_MODE] +
                                                    (f" ({eng.comm_ghost_info(2)} B/step on rank 0, " + ("peer memory" if eng.comm_ghost_info(3) == 1 else "NCCL") + ")"
                                                     if PARTITION_MODE == "exchange" else "") +
                                                    "; the other mode: GTK_PARTITION_MODE=exchange|recompute; config5 reports both"),
            "fast_path": eng.info(5), "cpu_affinity": affinity,
            "setup_phases_ms_rank0": None if world == 1 else {k: round(v, 3) for k, v in setup_tm.items()},
            "symbolic_ms": symbolic_ms, "symbolic_first_ms": symbolic_first_ms, "input_upload_ms": upload_ms, "host_prep_s": host_prep_s,
            "dofs_per_s": dofs_total / (ms_per_step * 1e-3),
            "reassembly": {"device_ms": ms_per_step, "e2e_ms": 1e3 * e2e_s, "value": value, "e2e_value": nnz_total / e2e_s},
            "first_assembly": first,
            "e2e": {"value": nnz_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(xyz_np.nbytes),
                    "d2h_bytes_per_step": int(nz_np.nbytes + b_np.nbytes), "ms_per_step": 1e3 * e2e_s, "checksum": checksum},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "achieved": alg / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms_per_step * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": alg / (ms_per_step * 1e-3) / 1e9 / 8000.0,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": alg, "bytes_per_nnz": alg / max(nnz_local, 1), "kernel": dominant,
                         "kernels_ms": kernels, "basis": "whole step device time (all kernels of the step)"},
            "clocks": clocks, "general_path": general, "mixed_path": mixed, "unstructured_path": unstructured,
        }
    eng.close()
    eng = None
    del xyz_pin, nz_pin, b_pin
    torch.cuda.synchronize()

    if world == 1 and not args.no_high_order:
        # BASELINE config 3 next to the headline (never as the headline): Q3 hexahedra, element-matrix GEMM on the
        # FP64 tensor cores, roofline = measured DMMA issue rate; full-size invariants checked in the same run
        try:
            import bench_highorder
            line["high_order"] = bench_highorder.run(64, 3, steps=5, warmup=2, device=local_rank, check=True)
            if not args.no_cpu_baseline:
                line["high_order"]["cpu_baseline"] = cpu_baseline_high_order()
        except Exception as exc:   # reported, never hidden
            line["high_order"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world == 1 and not args.no_elasticity:
        # BASELINE config 4 element at the largest size whose host preparation stays within the bench's time budget
        try:
            import bench_elasticity
            line["elasticity"] = bench_elasticity.run(64, steps=3, warmup=1, device=local_rank, check=True)
            if not args.no_cpu_baseline:
                line["elasticity"]["cpu_baseline"] = bench_elasticity.cpu_baseline()
        except Exception as exc:   # reported, never hidden
            line["elasticity"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world == 1 and not args.no_multifield:
        # a product space (Stokes, Q2 x 3 x Q1) and a skeleton integral through the block kernels
        try:
            import bench_multifield
            line["multifield"] = bench_multifield.run(24, 48, steps=3, warmup=1, device=local_rank)
        except Exception as exc:   # reported, never hidden
            line["multifield"] = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_config5:
        try:
            c5 = run_config5(torch, dist, E, P, tab, rank, world, local_rank, stream)
            if world > 1:      # the other way of completing the own rows, same job, same T_1
                other = "exchange" if PARTITION_MODE == "recompute" else "recompute"
                c5o = run_config5(torch, dist, E, P, tab, rank, world, local_rank, stream, mode=other, t1_known=c5.get("t1_ms"))
                c5["other_mode"] = {k: c5o.get(k) for k in ("partition_mode", "partition", "ms_per_step", "value", "strong_efficiency", "symbolic_ms",
                                                           "exchange", "setup_phases_ms_rank0")}
        except Exception as exc:   # reported, never hidden
            c5 = {"error": f"{type(exc).__name__}: {exc}"}
        if rank == 0:
            line["config5"] = c5
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(n)
        if SPIN_MS > 0:
            line["spin_up_ms"] = SPIN_MS
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
