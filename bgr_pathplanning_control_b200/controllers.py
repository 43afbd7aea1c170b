"""Drop-in controller classes: the reference's names, signatures, return conventions and error behaviour
(core/controllers.py), with every law evaluated by the CUDA kernels through ``bgr_step_host`` /
``bgr_mpc_eval_host``.  One controller instance = one episode, as in the reference (the PID integral and
the Kalman state live in the instance).  For many episodes use ``engine.rollout`` -- one boundary crossing
per rollout instead of one per step.

There is no CPU fallback: these classes raise if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import hashlib
from collections import OrderedDict

import numpy as np

from . import _abi as abi
from . import engine
from .vehicle_config import conf

MAX_STEER_ANGLE = conf.MAX_STEER   # core/controllers.py:15

_DEVICE = 0


def set_device(index):
    """CUDA device the scalar drop-in classes run on (default 0)."""
    global _DEVICE
    _DEVICE = int(index)
    _track_cache.clear()


# ---- device-resident copies of the `path` arrays callers pass every step ----
_track_cache: "OrderedDict[tuple, engine.Track]" = OrderedDict()
_TRACK_CACHE_SIZE = 8


def _track_for(cx, cy, curve=None):
    cx = np.ascontiguousarray(cx, dtype=np.float64)
    cy = np.ascontiguousarray(cy, dtype=np.float64)
    h = hashlib.blake2b(digest_size=16)
    h.update(cx.tobytes())
    h.update(cy.tobytes())
    if curve is not None:
        curve = np.ascontiguousarray(curve, dtype=np.float64)
        h.update(curve.tobytes())
    key = (_DEVICE, len(cx), h.digest())
    t = _track_cache.get(key)
    if t is None:
        t = engine.Track(cx, cy, curve, closed=False, device=_DEVICE)
        _track_cache[key] = t
        while len(_track_cache) > _TRACK_CACHE_SIZE:
            _track_cache.popitem(last=False)[1].close()
    else:
        _track_cache.move_to_end(key)
    return t


def _dummy_track():
    """The accel laws and ``State.update`` take no path; the kernel still stages one."""
    return _track_for(np.array([0.0, 1.0, 2.0]), np.array([0.0, 0.0, 0.0]))


def _state_row(state):
    return np.array([[float(state.x), float(state.y), float(state.yaw), float(state.v),
                      float(getattr(state, "r", 0.0)), float(getattr(state, "beta", 0.0))]])


def _step(track, cfg, phases, params, state_row, target_ind=0, ctrl=None, cmd=None):
    """One ``bgr_step_host`` call for a single vehicle.  Returns (out row, new target index)."""
    lib = abi.load()
    ti = np.array([int(target_ind)], dtype=np.int32)
    out = np.zeros((1, abi.N_OUT))
    c = cfg.to_struct()
    abi.check(lib.bgr_step_host(track.handle, C.byref(c), int(phases), 1, params.ctypes.data, state_row.ctypes.data,
                                ti.ctypes.data, None if ctrl is None else ctrl.ctypes.data,
                                None if cmd is None else cmd.ctypes.data, out.ctypes.data, None))
    return out[0], int(ti[0])


def _path_columns(path):
    path = np.asarray(path, dtype=np.float64)
    if path.ndim != 2 or path.shape[1] < 3:
        raise IndexError("path must be an (N, 3) array [index, x, y]")   # the reference indexes path[:, 1], path[:, 2]
    return path[:, 1], path[:, 2]


class PurePursuitController:
    """core/controllers.py:183-230."""

    def __init__(self, look_ahead_distance=conf.LOOKAHEAD_DISTANCE):
        self.look_ahead_distance = look_ahead_distance

    def compute_steering(self, state, path, target_ind):
        """-> (steering [rad] as np.float64, target index as int)   (:193-230)"""
        cx, cy = _path_columns(path)
        track = _track_for(cx, cy)
        params = engine.default_params(1, np.float64)
        params[0, abi.P_LF] = self.look_ahead_distance
        cfg = engine.RolloutConfig(steer="pure_pursuit", precision="fp64")
        out, ti = _step(track, cfg, abi.PHASE_STEER, params, _state_row(state), target_ind,
                        cmd=np.zeros((1, abi.N_CMD)))
        return np.float64(out[abi.O_STEER]), ti


class StanleyController:
    """The LIVE definition, core/controllers.py:325-385 (k = 1.0, k_soft = 0.1)."""

    def __init__(self, k=1.0, k_soft=0.1):
        self.k = k
        self.k_soft = k_soft

    def compute_steering(self, state, path, target_ind):
        """-> (steering as np.float64, nearest index as np.int64); ``target_ind`` is ignored (:356)."""
        cx, cy = _path_columns(path)
        track = _track_for(cx, cy)
        params = engine.default_params(1, np.float64)
        params[0, abi.P_K_STANLEY], params[0, abi.P_K_SOFT] = self.k, self.k_soft
        cfg = engine.RolloutConfig(steer="stanley", precision="fp64")
        out, ti = _step(track, cfg, abi.PHASE_STEER | abi.PHASE_ERRORS, params, _state_row(state), 0,
                        cmd=np.zeros((1, abi.N_CMD)))
        return np.float64(out[abi.O_STEER]), np.int64(ti)


class ShadowedStanleyController:
    """The SHADOWED first body of ``class StanleyController`` (core/controllers.py:234-324): adaptive gain
    ``k / (v + k_soft)``, steering-rate limit, low-pass filter, carried ``previous_steering``.  Dead code in the
    reference (Python keeps the later definitions, :325-385) -- offered as an explicitly labelled extra with the
    dead code's own signature ``compute_steering(state, path, target_ind, dt)``."""

    def __init__(self, k=conf.k_stanley, k_soft=conf.k_soft_stanley, max_steer_rate=np.deg2rad(30), alpha=0.5,
                 lookahead_distance=0.0):
        self.k, self.k_soft = k, k_soft
        self.max_steer_rate, self.alpha = max_steer_rate, alpha
        self.lookahead_distance = lookahead_distance
        self.previous_steering = 0.0                                      # :250

    def compute_steering(self, state, path, target_ind, dt):
        cx, cy = _path_columns(path)
        track = _track_for(cx, cy)
        params = engine.default_params(1, np.float64)
        params[0, abi.P_K_STANLEY], params[0, abi.P_K_SOFT] = self.k, self.k_soft
        cfg = engine.RolloutConfig(steer="stanley", precision="fp64", dt=dt, stanley_shadowed=True,
                                   stanley_max_steer_rate=self.max_steer_rate, stanley_alpha=self.alpha,
                                   stanley_lookahead=max(0.0, float(self.lookahead_distance)))     # :276-281
        ctrl = np.zeros((1, abi.N_CTRL))
        ctrl[0, abi.C_PREV_STEER] = self.previous_steering
        out, ti = _step(track, cfg, abi.PHASE_STEER | abi.PHASE_ERRORS, params, _state_row(state), 0, ctrl,
                        cmd=np.zeros((1, abi.N_CMD)))
        self.previous_steering = float(ctrl[0, abi.C_PREV_STEER])        # :319
        return np.float64(out[abi.O_STEER]), np.int64(ti)


class AccelerationPIDController:
    """core/controllers.py:18-76; the PID arithmetic is simple-pid's with a fixed sample period
    ``conf.dt`` (the real package reads the wall clock -- see DESIGN.md 'deviations')."""

    def __init__(self, kp, ki, kd, setpoint):
        self.kp, self.ki, self.kd = kp, ki, kd
        self.maxspeed = setpoint
        self._ctrl = np.zeros((1, abi.N_CTRL))

    def compute_acceleration(self, current_speed, curvature):
        params = engine.default_params(1, np.float64)
        params[0, abi.P_KP], params[0, abi.P_KI], params[0, abi.P_KD] = self.kp, self.ki, self.kd
        cfg = engine.RolloutConfig(accel="pid", precision="fp64")
        state = np.zeros((1, abi.N_STATE))
        state[0, abi.V] = float(current_speed)
        cmd = np.zeros((1, abi.N_CMD))
        cmd[0, abi.CMD_KAPPA] = float(curvature)
        out, _ = _step(_dummy_track(), cfg, abi.PHASE_ACCEL | abi.PHASE_KAPPA_GIVEN, params, state, 0, self._ctrl, cmd)
        self.maxspeed = np.float64(out[abi.O_V_DESIRED])          # :43
        return np.float64(out[abi.O_ACCEL])

    def compute_breaking(self, curvature):
        """Desired speed for a curvature (:58-73), evaluated by the same kernel (no state change)."""
        params = engine.default_params(1, np.float64)
        cfg = engine.RolloutConfig(accel="pid", precision="fp64")
        cmd = np.zeros((1, abi.N_CMD))
        cmd[0, abi.CMD_KAPPA] = float(curvature)
        out, _ = _step(_dummy_track(), cfg, abi.PHASE_ACCEL | abi.PHASE_KAPPA_GIVEN, params, np.zeros((1, abi.N_STATE)),
                       0, self._ctrl.copy(), cmd)
        return np.float64(out[abi.O_V_DESIRED])


class LQGAccelerationController:
    """core/controllers.py:78-180: discrete LQR on the augmented speed-error model + 2-state Kalman filter."""

    def __init__(self, dt):
        self.dt = dt
        # the public matrices of the reference instance (:83-94, :101-104); the gain K and every later P come from the
        # library's own recursions (bgr_lqg_design / bgr_lqg_covariance), which also feed the kernels
        self.A = np.array([[1, 0], [dt, 1]])                       # :83-84
        self.B = np.array([[dt], [0]])                             # :85-86
        self.C = np.array([[1, 0]])                                # :87
        self.Q = np.diag([10.0, 100.0])                            # :90-92
        self.R = np.array([[0.001]])                               # :93-94
        K, _ = engine.lqg_design(dt, 0)
        self.K = K.reshape(1, 2)                                   # :97
        self.Q_kf = np.diag([0.1, 0.1])                            # :101-103
        self.R_kf = np.array([[0.5]])                              # :104
        self.v_desired = 0.0
        self.a_k_prev = 0.0
        self._ctrl = np.zeros((1, abi.N_CTRL))

    @property
    def x_hat(self):
        return self._ctrl[0, abi.C_LQG_XHAT0:abi.C_LQG_XHAT1 + 1].reshape(2, 1).copy()

    @property
    def P(self):
        """Estimate covariance (:109), advanced once per ``compute_acceleration`` call (:126, :160).  It does not depend
        on the measurements, so it is the shared recursion evaluated at this instance's call count."""
        return engine.lqg_covariance(self.dt, int(self._ctrl[0, abi.C_STEP]))

    def compute_acceleration(self, current_speed, curvature):
        params = engine.default_params(1, np.float64)
        cfg = engine.RolloutConfig(accel="lqg", precision="fp64", dt=self.dt)
        state = np.zeros((1, abi.N_STATE))
        state[0, abi.V] = float(current_speed)
        cmd = np.zeros((1, abi.N_CMD))
        cmd[0, abi.CMD_KAPPA] = float(curvature)
        out, _ = _step(_dummy_track(), cfg, abi.PHASE_ACCEL | abi.PHASE_KAPPA_GIVEN, params, state, 0, self._ctrl, cmd)
        self.v_desired = np.float64(out[abi.O_V_DESIRED])          # :120
        self.a_k_prev = float(out[abi.O_ACCEL])                    # :136
        return float(out[abi.O_ACCEL])

    def compute_desired_speed(self, curvature):
        return AccelerationPIDController(0, 0, 0, 0).compute_breaking(curvature)   # same formula, :165-180


class MPCController:
    """core/controllers.py:471-568 with the cvxpy/OSQP solve replaced by GPU evaluation of the same cost
    over a bank of candidate control sequences: the optimum over the bank, which for the default ``qp_family`` bank
    lies within 0.2 % of the QP optimum in cost (tests/test_gpu_mpc.py)."""

    def __init__(self, N=10, dt=0.1, n_candidates=4096, seed=20261021, bank="qp_family"):
        self.N = N
        self.dt = dt
        self.Q = np.diag([1.0, 1.0, 0.5, 0.1])                     # :483
        self.R = np.diag([0.1, 0.1])                               # :484
        # default: the bank built from the QP's own separable structure -- its best candidate is within 0.2 % of the
        # optimum OSQP converges to (engine.candidate_bank); bank="uniform" gives the generic random sample
        self.bank = engine.candidate_bank(n_candidates, N, seed, kind=bank, dt=dt)
        self._resident = None
        self.last_cost = None

    def compute_control(self, state, path):
        """-> (acceleration, steering) -- note the order and the negated steering (:568)."""
        cx, cy = _path_columns(path)
        track = _track_for(cx, cy)
        cfg = engine.MpcConfig(horizon=self.N, dt=self.dt, q=tuple(np.diag(self.Q)), r=tuple(np.diag(self.R)))
        s = np.array([[float(state.x), float(state.y), float(state.yaw), float(state.v)]])
        if self._resident is None or self._resident.device != _DEVICE:
            self._resident = engine.MpcBank(self.bank, device=_DEVICE)      # uploaded once per controller
        res = engine.mpc_evaluate(track, cfg, s, np.zeros(1, np.int32), self._resident)
        if int(res.best_idx[0]) < 0:
            print("MPC optimization failed. Using zero acceleration and steering.")   # :562
        self.last_cost = float(res.best_cost[0])
        return np.float64(res.best_u[0, 0]), np.float64(res.best_u[0, 1])


def get_steering_controller(name):
    """core/controllers.py:572-589."""
    if name == 'pure_pursuit':
        return PurePursuitController(look_ahead_distance=conf.LOOKAHEAD_DISTANCE)
    elif name == 'stanley':
        return StanleyController(k=1.0, k_soft=0.1)
    elif name == 'mpc':
        return MPCController(N=10, dt=0.05)
    else:
        raise ValueError(f"Unknown steering controller name: {name}")


def normalize_angle(angle):
    """core/controllers.py:592-603: floor-mod wrap to [-pi, pi).  Host one-liner kept for API parity; the
    kernels carry their own device version."""
    return (angle + np.pi) % (2 * np.pi) - np.pi


def find_target_point(state, cx, cy, target_ind):
    """core/controllers.py:605-612: the look-ahead scan with ``conf.LOOKAHEAD_DISTANCE``, capped at
    ``len(cx) - 1`` -- the same device scan as pure pursuit."""
    track = _track_for(cx, cy)
    params = engine.default_params(1, np.float64)
    cfg = engine.RolloutConfig(steer="pure_pursuit", precision="fp64")
    _, ti = _step(track, cfg, abi.PHASE_STEER, params, _state_row(state), target_ind, cmd=np.zeros((1, abi.N_CMD)))
    return ti


def compute_control_commands(controller, path, state, target_ind=0, accel_controller=None, curve=None):
    """README-derived convenience (README.md:98-103 names a ``compute_control_commands`` that does not exist
    in the reference code): one steering + one acceleration command in the order of
    nodes/controller.py:61-63.  Returns (steering, acceleration, new_target_ind)."""
    if hasattr(controller, "compute_control"):
        acceleration, steering = controller.compute_control(state, path)
        return steering, acceleration, target_ind
    steering, target_ind = controller.compute_steering(state, path, target_ind)
    acceleration = 0.0
    if accel_controller is not None:
        kappa = 0.0
        if curve is not None:
            kappa = curve[target_ind] if target_ind < len(curve) else curve[-1]      # nodes/controller.py:62
        acceleration = accel_controller.compute_acceleration(state.v, kappa)
    return steering, acceleration, target_ind
