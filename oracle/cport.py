"""ctypes binding of the plain-C oracle twin (oracle/c/bgr_oracle.c).  TEST INFRASTRUCTURE (oracle/__init__.py).

``rollout`` runs E episodes with OpenMP over episodes and returns arrays in the product's BGR_M_* / BGR_S_* /
BGR_T_* column layout (as float64), so parity tests can compare thousands of full-length episodes directly.
The C twin is pinned against the numpy oracle in tests/test_oracle_c.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import laws, loop

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "c", "liboracle.so")
N_METRICS, N_STATUS, N_TRAJ = 12, 10, 12


class Cfg(C.Structure):
    _fields_ = [("steer", C.c_int32), ("accel", C.c_int32), ("model", C.c_int32), ("max_steps", C.c_int32),
                ("laps", C.c_int32), ("laps_target", C.c_int32), ("closed", C.c_int32), ("n", C.c_int32),
                ("n_cones", C.c_int32),
                ("dt", C.c_double), ("WB", C.c_double), ("l_f", C.c_double), ("l_r", C.c_double),
                ("max_steer", C.c_double), ("max_accel", C.c_double), ("max_decel", C.c_double),
                ("hit_radius", C.c_double), ("lqr_k0", C.c_double), ("lqr_k1", C.c_double),
                ("sigma_xy", C.c_double), ("sigma_yaw", C.c_double), ("sigma_v", C.c_double),
                ("sigma_cone", C.c_double), ("seed", C.c_uint32), ("noise_on", C.c_uint32),
                ("episode_id_base", C.c_int64),
                ("replan_horizon", C.c_int32), ("replan_stride", C.c_int32), ("replan_after", C.c_int32),
                ("reserved", C.c_int32), ("planner_lookahead", C.c_double)]


_lib = None


def available():
    return os.path.isfile(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError(f"{LIB_PATH} is not built: run `make -C oracle/c` (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        _lib.oracle_rollout.restype = C.c_int
        _lib.oracle_rollout.argtypes = [C.POINTER(Cfg)] + [C.c_void_p] * 4 + [C.c_int64] + [C.c_void_p] * 5 + [C.c_int]
    return _lib


def rollout(track: loop.Track, spec: loop.RolloutSpec, params, init, episode_id_base=0, traj=False, threads=None):
    """``params`` (E, 16), ``init`` (E, 6) -> (metrics (E, 12) f64, status (E, 10) i32, traj (E, S, 12) f64 | None)."""
    lib = load()
    conf = spec.conf
    params = np.ascontiguousarray(params, dtype=np.float64)
    init = np.ascontiguousarray(init, dtype=np.float64)
    E = params.shape[0]
    assert params.shape == (E, 16) and init.shape == (E, 6)
    cx, cy, kap = (np.ascontiguousarray(a, dtype=np.float64) for a in (track.x, track.y, track.kappa))
    cones = np.ascontiguousarray(track.cones, dtype=np.float64).reshape(-1, 2)
    c = Cfg()
    c.steer = {"pure_pursuit": 0, "stanley": 1}[spec.steer]
    c.accel = {"pid": 0, "lqg": 1}[spec.accel]
    c.model = {"kinematic": 0, "dynamic": 1}[spec.model]
    c.max_steps, c.laps, c.laps_target = spec.max_steps, spec.laps, spec.laps_target
    c.closed, c.n, c.n_cones = int(track.closed), len(cx), len(cones)
    c.dt, c.WB, c.l_f, c.l_r = spec.dt, conf.WB, conf.l_f, conf.l_r
    c.max_steer, c.max_accel, c.max_decel = conf.MAX_STEER, conf.MAX_ACCEL, conf.MAX_DECEL
    c.hit_radius = spec.hit_radius
    if spec.accel == "lqg":
        K = laws.lqr_gain(spec.dt)
        c.lqr_k0, c.lqr_k1 = float(K[0, 0]), float(K[0, 1])
    if spec.noise is not None:
        nz = spec.noise
        c.sigma_xy, c.sigma_yaw, c.sigma_v, c.sigma_cone = nz.sigma_xy, nz.sigma_yaw, nz.sigma_v, nz.sigma_cone
        c.seed, c.noise_on = nz.seed & 0xFFFFFFFF, 1
    c.episode_id_base = episode_id_base
    c.replan_horizon, c.replan_stride, c.replan_after = spec.replan_horizon, spec.replan_stride, spec.replan_after
    c.planner_lookahead = conf.LOOKAHEAD_DISTANCE
    metrics = np.empty((E, N_METRICS))
    status = np.empty((E, N_STATUS), dtype=np.int32)
    tr = np.zeros((E, spec.max_steps, N_TRAJ)) if traj else None
    threads = threads or len(os.sched_getaffinity(0))
    rc = lib.oracle_rollout(C.byref(c), cx.ctypes.data, cy.ctypes.data, kap.ctypes.data,
                            cones.ctypes.data if len(cones) else None, E, params.ctypes.data, init.ctypes.data,
                            metrics.ctypes.data, status.ctypes.data, None if tr is None else tr.ctypes.data, threads)
    assert rc == 0
    return metrics, status, tr
