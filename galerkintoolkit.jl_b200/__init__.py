"""galerkintoolkit.jl_b200 — B200-native assembly engine for GalerkinToolkit.jl's hot path.

    csrc/ + libgtkasm.so   hand-written sm_100a CUDA kernels behind the C ABI of include/gtk_assembly.h
    engine.py              ctypes binding of that ABI (stand-in for the Julia ccall shim)
    gt.py                  host-side mirror of the reference's user API (∫, lagrange_space, assemble_*)
    hostprep.py            input preparation (mesh / dof maps / tabulations) as flat arrays
    partition.py           z-slab partition + ghost-row plan for the multi-GPU path

The directory name contains a dot, so import it through the repo-root shim:  ``import gtk_b200``.
Nothing in this package imports ``oracle/`` and nothing computes on the CPU: if libgtkasm.so or a
GPU is missing, engine construction raises.
"""
from . import hostprep  # noqa: F401
from . import engine  # noqa: F401
from . import gt  # noqa: F401
from . import gt as GT  # noqa: F401
from .build import build as build_library  # noqa: F401

__all__ = ["hostprep", "engine", "gt", "GT", "build_library"]
