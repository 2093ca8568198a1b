"""bluerov2_b200 -- batched, B200-native SQP-RTI solver for the BlueROV2 NMPC OCP (hot path only).

The compute lives in ``csrc/`` (hand-written sm_100a CUDA behind a C-ABI shared library,
``libacados_ocp_solver_bluerov2.so``); this package is the thin host-side mirror of that ABI plus the
synthetic-workload generators used by tests and bench.py.  There is no CPU fallback: every solver entry
point raises if the CUDA library is missing or no GPU is present.
"""
from . import traj, workloads  # noqa: F401

__all__ = ["traj", "workloads"]
