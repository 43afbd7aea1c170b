"""Import the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  /root/reference does not exist on
the GPU box, so nothing that runs there may call into this file;  it is used by
``oracle/gen_golden.py`` (to freeze golden vectors) and by the ``not gpu`` test
``tests/test_oracle_vs_reference_live.py`` (skipped when the tree is absent).

``core/controllers.py`` imports, at module top, seven third-party modules that are
not installed here and that the hot-path arithmetic never touches, except
``simple_pid.PID`` and ``control.dlqr`` (core/controllers.py:6-11).  We inject
``sys.modules`` stubs for them:

  cvxpy, msgpackrpc, fsd_path_planning(.utils(.math_utils))  -> empty modules
  simple_pid.PID   -> ``FixedDtPID``: restated simple-pid 2.x semantics with a FIXED
                      dt (the real one reads time.monotonic(); "parity unpinned")
  control.dlqr     -> scipy DARE based restatement ("parity unpinned")

Everything else executed is the reference's own bytes.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("BGR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "controllers.py"))


class FixedDtPID:
    """Restated ``simple_pid.PID`` (2.x) with a fixed sample period.

    Call sites in the reference: core/controllers.py:29 (ctor), :44 (setpoint), :45
    (``self.pid(current_speed)``).  Upstream semantics restated (parity unpinned, the
    package is neither vendored nor pinned):
        error      = setpoint - input
        P          = Kp * error
        I         += Ki * error * dt          (output_limits default (None, None))
        D          = -Kd * (input - last_input) / dt   (derivative on measurement; 0 on first call)
        output     = P + I + D
    Upstream takes dt from time.monotonic() and holds the previous output when
    dt < sample_time (0.01 s); a faster-than-real-time loop cannot use that, so dt
    is the fixed simulation step (SURVEY.md section 7 "simple_pid is wall-clock based").
    """

    dt = 0.05  # class-level so the shim can set it before the reference builds a PID

    def __init__(self, Kp=1.0, Ki=0.0, Kd=0.0, setpoint=0.0):
        self.Kp, self.Ki, self.Kd = Kp, Ki, Kd
        self.setpoint = setpoint
        self._integral = 0.0
        self._last_input = None

    def __call__(self, input_):
        dt = self.dt
        error = self.setpoint - input_
        d_input = input_ - (self._last_input if self._last_input is not None else input_)
        proportional = self.Kp * error
        self._integral += self.Ki * error * dt
        derivative = -self.Kd * d_input / dt
        self._last_input = input_
        return proportional + self._integral + derivative


def dlqr_scipy(A, B, Q, R):
    """Restated ``control.dlqr`` (core/controllers.py:97): K = (R + B'PB)^-1 B'PA."""
    from scipy.linalg import solve_discrete_are

    A = np.asarray(A, dtype=float)
    B = np.asarray(B, dtype=float)
    P = solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)
    return K, P, None


_loaded = None


def load_reference():
    """Return a namespace with the reference's unmodified modules (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in (
        "cvxpy",
        "simple_pid",
        "control",
        "msgpackrpc",
        "fsd_path_planning",
        "fsd_path_planning.utils",
        "fsd_path_planning.utils.math_utils",
    ):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["simple_pid"].PID = FixedDtPID
    sys.modules["control"].dlqr = dlqr_scipy

    class ConeTypes:  # only referenced by out-of-scope sim glue
        UNKNOWN, RIGHT, LEFT, ORANGE_SMALL, ORANGE_BIG = range(5)

    sys.modules["fsd_path_planning"].ConeTypes = ConeTypes
    sys.modules["fsd_path_planning.utils.math_utils"].unit_2d_vector_from_angle = (
        lambda a: np.array([np.cos(a), np.sin(a)])
    )

    import importlib

    ns = types.SimpleNamespace()
    ns.vehicle_config = importlib.import_module("vehicle_config")
    ns.conf = ns.vehicle_config.Vehicle_config
    ns.car_state = importlib.import_module("core.data.car_state")
    ns.controllers = importlib.import_module("core.controllers")
    ns.metrics_calculator = importlib.import_module("preformance_analysis.metrics_calculator")
    ns.FixedDtPID = FixedDtPID
    _loaded = ns
    return ns
