"""B200-native batched convex-MPC engine for the TRON1 point-foot controller.

Layout: csrc/ (sm_100a kernels + the C ABI of include/mpc_b200.h), host/ (C++ facade mirroring
the reference's QPSolver / mpcQP / MPC classes), engine.py + _capi.py (ctypes/torch plumbing for
tests and bench), synth.py (synthetic workloads).  Importing this package does not load CUDA;
constructing an Engine does, and fails loudly when the extension or the GPU is missing."""
from . import synth  # noqa: F401

__all__ = ["synth"]
