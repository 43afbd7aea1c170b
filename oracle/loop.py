"""The composed closed loop: the "reference numpy path" of one episode.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference has no FSDS-free loop; this file
COMPOSES the restated reference functions of ``oracle/laws.py`` in the order of the live caller
(nodes/controller.py:61-63) and of the legacy loop (old/animation_loop.py:71,87-114), replacing
the simulator by ``State.update`` (core/data/car_state.py:118-125).

Per step k (t = k * dt), for an episode that has not terminated:
  1. observation   obs = true state (+ ORACLE-DEFINED Philox noise, config 5)
  2. steering      PurePursuit (core/controllers.py:193) over the UNROLLED path (base lap tiled
                   ``laps`` times; index monotone, :212-220)  or the live Stanley (:336) over the
                   base lap (global argmin, :353-356)
  3. curvature     kappa = curve[ti] if ti < len(curve) else curve[-1]   (nodes/controller.py:62)
  4. acceleration  PID (:32, restated simple_pid with fixed dt) or LQG (:117)
  5. logging       ORACLE-DEFINED use of Stanley's own formulas (:353-376) on the TRUE state over the
                   base lap: lateral_error := e_ct, heading_error := theta_e at the global-argmin
                   nearest waypoint; timestamp := k * dt  (columns of simulation_logger.py:8-10)
  6. cone hits     ORACLE-DEFINED: cone c is hit when hypot(cone - (x, y)) <= hit_radius at a logged
                   step; the count is of DISTINCT cones
  7. lap counter   ORACLE-DEFINED (closed tracks): nearest index wrapping forward by more than half a
                   lap => lap += 1 (backward => lap -= 1)
  8. vehicle       State.update(a, delta) -> dynamic bicycle (car_state.py:178-223), or the
                   ORACLE-DEFINED kinematic bicycle
  2'. re-planning  (``replan_horizon`` > 0, pure pursuit only) the regime of the live stack: Planner.update
                   (nodes/planner.py:17-102) keeps a SHORT local path and its own look-ahead index on it; it plans a
                   new path (index back to 0, :62) whenever that index has passed ``replan_after`` (:26), advances the
                   index by its own scan with conf.LOOKAHEAD_DISTANCE (:72-77) and hands path + index to the
                   controller, which runs compute_steering on the local path from that index (nodes/controller.py:
                   56-62).  ORACLE-DEFINED stand-in for fsd-path-planning (out of scope): the path planned from a car
                   position is the window of the centreline starting at the waypoint nearest to it, every
                   ``replan_stride``-th waypoint, ``replan_horizon`` points (modulo n when closed, clipped when open).
  9. termination   after the step: target index reached the last path index (old/animation_loop.py:71
                   ``lastIndex > target_ind``), or ``laps_target`` laps completed, or non-finite state;
                   else the step cap (``T >= curr_time`` in the legacy loop).
Metrics = preformance_analysis/metrics_calculator.py:1-8 over the logged rows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import laws, philox

# ---- per-episode parameter columns (restated; tests assert equality with the product's layout) ----
P_LF, P_KP, P_KI, P_KD, P_TARGET_SPEED, P_K_EXPO, P_K_STANLEY, P_K_SOFT = range(8)
P_MU, P_C_F, P_C_R, P_MASS, P_I_Z = range(8, 13)
N_PARAMS = 16

# ---- status flag bits ----
ST_PATH_END = 1      # target index reached the last path index
ST_LAPS_DONE = 2     # laps_target completed
ST_TIMEOUT = 4       # step cap reached
ST_NONFINITE = 8     # state became inf/nan


def default_params(conf: laws.VehicleConfig = laws.CONF) -> np.ndarray:
    p = np.zeros(N_PARAMS)
    p[P_LF] = conf.LOOKAHEAD_DISTANCE
    p[P_KP], p[P_KI], p[P_KD] = conf.kp_accel, conf.ki_accel, conf.kd_accel
    p[P_TARGET_SPEED] = conf.TARGET_SPEED
    p[P_K_EXPO] = conf.k_expo
    p[P_K_STANLEY], p[P_K_SOFT] = 1.0, 0.1          # factory values, core/controllers.py:585
    p[P_MU], p[P_C_F], p[P_C_R] = conf.mu, conf.c_f, conf.c_r
    p[P_MASS], p[P_I_Z] = conf.m, conf.I_z
    return p


@dataclass
class NoiseSpec:
    seed: int = 0
    sigma_xy: float = 0.0
    sigma_yaw: float = 0.0
    sigma_v: float = 0.0
    sigma_cone: float = 0.0


@dataclass
class RolloutSpec:
    steer: str = "pure_pursuit"       # "pure_pursuit" | "stanley" | "stanley_shadowed" (dead-code variant, :234-324)
    shadowed_rate: float = laws.SHADOWED_RATE
    shadowed_alpha: float = laws.SHADOWED_ALPHA
    shadowed_lookahead: float = 0.0   # the variant's optional look-ahead advance (:276-281)
    accel: str = "pid"                # "pid" | "lqg"
    model: str = "kinematic"          # "kinematic" | "dynamic"
    max_steps: int = 2000
    dt: float = laws.CONF.dt
    laps: int = 1                     # base lap tiled this many times for the monotone PP index
    laps_target: int = 0              # stop after this many completed laps (0 = off)
    hit_radius: float = 0.0           # 0 = cone hits off
    noise: NoiseSpec | None = None
    replan_horizon: int = 0           # > 0: receding-horizon replay with local paths of this many points
    replan_stride: int = 1
    replan_after: int = 3             # nodes/planner.py:26
    conf: laws.VehicleConfig = field(default_factory=lambda: laws.CONF)


@dataclass
class Track:
    """Centreline as the controllers see it: base lap x, y, curvature (float64 arrays)."""
    x: np.ndarray
    y: np.ndarray
    kappa: np.ndarray
    closed: bool = True
    cones: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))


def window_path_indices(n: int, closed: bool, j0: int, horizon: int, stride: int) -> np.ndarray:
    """ORACLE-DEFINED stand-in for the planner: waypoint indices of the local path planned at waypoint ``j0``."""
    g = j0 + stride * np.arange(horizon, dtype=np.int64)
    return g % n if closed else np.minimum(g, n - 1)


def planner_scan(cx, cy, x, y, target_ind: int, lookahead: float) -> int:
    """The planner's own index update, nodes/planner.py:72-77 (note ``< len - 1``, unlike the controller's scan)."""
    while target_ind < len(cx) - 1:
        distance = np.hypot(cx[target_ind] - x, cy[target_ind] - y)
        if distance >= lookahead:
            break
        target_ind += 1
    return target_ind


def rollout_episode(track: Track, spec: RolloutSpec, params, init_state, episode_id: int = 0,
                    record: bool = True):
    """One closed-loop episode in fp64.  ``init_state`` = [x, y, yaw, v, r, beta].

    Returns a dict: ``metrics`` (dict), ``status`` (int flags), ``steps`` (rows logged),
    ``cone_hits``, final ``state``, ``target_ind``, ``nearest_ind``; and if ``record``: arrays
    ``traj`` (steps, 6) state BEFORE each step, ``steer``, ``accel``, ``ti``, ``near``, ``e_ct``,
    ``theta_e`` and the decision margins ``pp_margin`` (min |d - Lf| over the waypoints the scan
    tested) and ``near_margin`` (second-best minus best distance of the nearest search) that the
    parity tests use to recognise exact-distance near-ties.
    """
    conf = spec.conf
    dt = spec.dt
    n_lap = len(track.x)
    cx_lap, cy_lap, curve_lap = track.x, track.y, track.kappa
    if spec.steer == "pure_pursuit":
        cx_tot = np.tile(cx_lap, spec.laps)
        cy_tot = np.tile(cy_lap, spec.laps)
        curve_tot = np.tile(curve_lap, spec.laps)
    else:
        cx_tot, cy_tot, curve_tot = cx_lap, cy_lap, curve_lap
    n_tot = len(cx_tot)
    p = np.asarray(params, dtype=np.float64)
    state = [float(s) for s in init_state]

    pid = laws.PIDState()
    lqg = laws.LQGState()
    if spec.accel == "lqg":
        mats = laws.lqg_matrices(dt)
        K_lqr = laws.lqr_gain(dt)

    cones = np.asarray(track.cones, dtype=np.float64).reshape(-1, 2)
    noise = spec.noise
    if noise is not None and noise.sigma_cone > 0.0 and len(cones):
        z = philox.gaussians4(np.arange(len(cones), dtype=np.uint32), philox.STREAM_CONE, 0, 0,
                              noise.seed, episode_id)
        cones = cones + noise.sigma_cone * np.stack([z[0], z[1]], axis=1)
    cone_hit = np.zeros(len(cones), dtype=bool)

    replan = spec.replan_horizon > 0
    if replan and spec.steer != "pure_pursuit":
        raise ValueError("re-planning replay is defined for pure pursuit (nodes/controller.py:23)")
    plan_idx = None                  # Planner.current_path (as waypoint indices), nodes/planner.py:13
    plan_ti = 0                      # Planner.target_ind, :14
    ti = 0
    prev_steer = 0.0                 # StanleyController.previous_steering (:250), shadowed variant only
    prev_near = None
    lap = 0
    status = 0
    ts, lat, hed = [], [], []
    rec = {k: [] for k in ("traj", "steer", "accel", "ti", "near", "e_ct", "theta_e",
                           "pp_margin", "near_margin")}

    for k in range(spec.max_steps):
        x, y, yaw, v, r, beta = state
        # 1. observation
        ox, oy, oyaw, ov = x, y, yaw, v
        if noise is not None and (noise.sigma_xy > 0 or noise.sigma_yaw > 0 or noise.sigma_v > 0):
            z = philox.gaussians4(k, philox.STREAM_OBS, 0, 0, noise.seed, episode_id)
            ox = x + noise.sigma_xy * float(z[0])
            oy = y + noise.sigma_xy * float(z[1])
            oyaw = yaw + noise.sigma_yaw * float(z[2])
            ov = v + noise.sigma_v * float(z[3])
        # 2. steering
        pp_margin = math.inf
        if replan:
            if plan_idx is None or plan_ti > spec.replan_after:                         # nodes/planner.py:26
                d_obs = np.hypot(cx_lap - ox, cy_lap - oy)
                j0 = int(np.argmin(d_obs))
                if record:           # decision margin of the planner's start point (near-tie recognition in the tests)
                    part = np.partition(d_obs, 1)
                    pp_margin = min(pp_margin, float(part[1] - part[0]))
                plan_idx = window_path_indices(n_lap, track.closed, j0, spec.replan_horizon, spec.replan_stride)
                plan_ti = 0                                                             # :62
            lx, ly = cx_lap[plan_idx], cy_lap[plan_idx]
            plan_ti = planner_scan(lx, ly, ox, oy, plan_ti, conf.LOOKAHEAD_DISTANCE)    # :72-77
            if record:
                for i in range(plan_ti + 1):      # (margins of both scans; the planner's own tests first)
                    pp_margin = min(pp_margin, abs(float(np.hypot(lx[i] - ox, ly[i] - oy)) - conf.LOOKAHEAD_DISTANCE))
                i = plan_ti
                while i < len(lx):
                    d = float(np.hypot(lx[i] - ox, ly[i] - oy))
                    pp_margin = min(pp_margin, abs(d - p[P_LF]))
                    if d >= p[P_LF]:
                        break
                    i += 1
            delta, ti = laws.pure_pursuit_steering(ox, oy, oyaw, lx, ly, plan_ti, p[P_LF],   # nodes/controller.py:61
                                                   conf.WB, conf.MAX_STEER)
        elif spec.steer == "pure_pursuit":
            if record:
                i = ti
                while i < n_tot:
                    d = float(np.hypot(cx_tot[i] - ox, cy_tot[i] - oy))
                    pp_margin = min(pp_margin, abs(d - p[P_LF]))
                    if d >= p[P_LF]:
                        break
                    i += 1
            delta, ti = laws.pure_pursuit_steering(ox, oy, oyaw, cx_tot, cy_tot, ti, p[P_LF],
                                                   conf.WB, conf.MAX_STEER)
        elif spec.steer == "stanley":
            delta, ti = laws.stanley_steering(ox, oy, oyaw, ov, cx_lap, cy_lap,
                                              p[P_K_STANLEY], p[P_K_SOFT], conf.MAX_STEER)
        elif spec.steer == "stanley_shadowed":
            delta, ti, prev_steer = laws.stanley_shadowed_steering(
                ox, oy, oyaw, ov, cx_lap, cy_lap, prev_steer, dt, p[P_K_STANLEY], p[P_K_SOFT],
                spec.shadowed_rate, spec.shadowed_alpha, conf.MAX_STEER, spec.shadowed_lookahead)
        else:
            raise ValueError(spec.steer)
        # 3. curvature lookup (nodes/controller.py:62)
        if replan:
            curve_loc = curve_lap[plan_idx]
            kappa = curve_loc[ti] if ti < len(curve_loc) else curve_loc[-1]
        else:
            kappa = curve_tot[ti] if ti < len(curve_tot) else curve_tot[-1]
        # 4. acceleration
        if spec.accel == "pid":
            a, _ = laws.pid_acceleration(pid, ov, kappa, p[P_KP], p[P_KI], p[P_KD], dt,
                                         p[P_TARGET_SPEED], p[P_K_EXPO], conf.MAX_DECEL, conf.MAX_ACCEL)
        elif spec.accel == "lqg":
            a, _ = laws.lqg_acceleration(lqg, K_lqr, mats, ov, kappa, p[P_TARGET_SPEED], p[P_K_EXPO],
                                         conf.MAX_DECEL, conf.MAX_ACCEL)
        else:
            raise ValueError(spec.accel)
        # 5. logging on the TRUE state
        near, theta_e, e_ct, _ = laws.nearest_point_errors(x, y, yaw, cx_lap, cy_lap)
        ts.append(k * dt)
        lat.append(e_ct)
        hed.append(theta_e)
        # 6. cone hits
        if spec.hit_radius > 0.0 and len(cones):
            cone_hit |= np.hypot(cones[:, 0] - x, cones[:, 1] - y) <= spec.hit_radius
        # 7. lap counter
        if prev_near is not None and track.closed:
            if prev_near - near > n_lap // 2:
                lap += 1
            elif near - prev_near > n_lap // 2:
                lap -= 1
        prev_near = near
        if record:
            d_all = np.hypot(cx_lap - x, cy_lap - y)
            part = np.partition(d_all, 1)
            rec["traj"].append(list(state))
            rec["steer"].append(delta)
            rec["accel"].append(a)
            rec["ti"].append(ti)
            rec["near"].append(near)
            rec["e_ct"].append(e_ct)
            rec["theta_e"].append(theta_e)
            rec["pp_margin"].append(pp_margin)
            rec["near_margin"].append(float(part[1] - part[0]))
        # 8. vehicle step (State.update(a, delta): car_state.py:118-125)
        if spec.model == "dynamic":
            state = laws.dynamic_bicycle_step(state, delta, a, dt, p[P_MASS], p[P_I_Z], conf.l_f,
                                              conf.l_r, p[P_C_F], p[P_C_R], p[P_MU])
        elif spec.model == "kinematic":
            state = laws.kinematic_bicycle_step(state, delta, a, dt, conf.WB)
        else:
            raise ValueError(spec.model)
        # 9. termination
        if not all(math.isfinite(s) for s in state):
            status |= ST_NONFINITE
        if replan:
            if not track.closed and plan_idx[ti] >= n_lap - 1:      # the target is the last waypoint of an open path
                status |= ST_PATH_END
        elif spec.steer == "pure_pursuit" and ti >= n_tot - 1:
            status |= ST_PATH_END
        if spec.steer in ("stanley", "stanley_shadowed") and not track.closed and ti >= n_tot - 1:
            status |= ST_PATH_END
        if spec.laps_target > 0 and lap >= spec.laps_target:
            status |= ST_LAPS_DONE
        if status:
            break
    else:
        status |= ST_TIMEOUT

    out = {
        "metrics": laws.lap_metrics(ts, lat, hed),
        "max_abs_lateral": float(np.max(np.abs(lat))),
        "status": status,
        "steps": len(ts),
        "laps": lap,
        "cone_hits": int(cone_hit.sum()),
        "state": list(state),
        "target_ind": int(ti),
        "nearest_ind": int(prev_near),
    }
    if record:
        out.update({k: np.asarray(v) for k, v in rec.items()})
    return out
