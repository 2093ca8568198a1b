"""oracle/ -- CPU checkers for the BlueROV2 SQP-RTI hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in ``bluerov2_b200`` (the product) may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

Two libraries (both built by ``oracle/Makefile``):

* ``oracle/libbluerov2_oracle.so``  -- our plain-C restatement (``bluerov2_oracle.c``) of one RTI step
  (ERK4 + sensitivities, Gauss-Newton NLS cost, box-constrained QP solved to 1e-12) and of the EKF.
* ``oracle/_ref/libbluerov2_casadi_ref.so`` -- the reference's own CasADi-generated C files compiled in
  place from ``/root/reference`` (bit-level oracle for f, J*Sx, J*Su + Ju).
"""
from .oracle import (  # noqa: F401
    Oracle,
    CasadiRef,
    build,
    NOMINAL_P,
    W_DEFAULT,
    WE_DEFAULT,
    LBU,
    UBU,
    X_INIT,
)
