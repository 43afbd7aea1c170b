"""fp64 restatement of every reference function on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Each function cites the reference
file:line it follows and keeps the reference's operation order, so that the golden
vectors frozen from the reference itself (tests/golden/, oracle/gen_golden.py) are
reproduced to the last bit on the same libm.

Scalars are Python floats; arrays are numpy float64.  ``np.*`` vs ``math.*`` calls
mirror the reference's own choice at each line (numpy's float64 sin/cos/hypot may be
SIMD kernels that differ from libm by an ulp).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------------------
# vehicle_config.py:4-39  (class attributes of Vehicle_config, restated as a dataclass)
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class VehicleConfig:
    m: float = 255.0                       # vehicle_config.py:7
    I_z: float = 1700.0                    # :8
    l_f: float = 0.9                       # :9
    l_r: float = 0.9                       # :10
    c_f: float = 16000.0                   # :11
    c_r: float = 17000.0                   # :12
    mu: float = 1.0                        # :13
    TARGET_SPEED: float = 25 / 3.6         # :14
    kp_accel: float = 0.35                 # :19
    ki_accel: float = 0.4                  # :20
    kd_accel: float = 0.3                  # :21
    k_expo: float = 0.05                   # :24
    k_stanley: float = 0.7                 # :26 (unused by the live Stanley law)
    k_soft_stanley: float = 0.01           # :27 (unused by the live Stanley law)
    MAX_ACCEL: float = 1.5                 # :30
    MAX_DECEL: float = -2.0                # :31
    MAX_SPEED: float = 40 / 3.6            # :33 (unused by any law)
    MAX_STEER: float = float(np.deg2rad(30))  # :35
    LOOKAHEAD_DISTANCE: float = 5.5        # :37
    dt: float = 0.05                       # :39

    @property
    def WB(self) -> float:                 # :16  WB = l_f + l_r
        return self.l_f + self.l_r


CONF = VehicleConfig()


# --------------------------------------------------------------------------------------
# core/controllers.py:592-603
# --------------------------------------------------------------------------------------
def normalize_angle(angle: float) -> float:
    """``(angle + pi) % (2 pi) - pi`` with Python floor-mod: result in [-pi, pi)."""
    return (angle + np.pi) % (2 * np.pi) - np.pi


# --------------------------------------------------------------------------------------
# core/controllers.py:193-230  PurePursuitController.compute_steering
# --------------------------------------------------------------------------------------
def pure_pursuit_steering(x, y, yaw, cx, cy, target_ind, Lf, WB=CONF.WB, max_steer=CONF.MAX_STEER):
    """Forward scan for the first index with distance >= Lf (:212-220), then the law (:223-229).

    Returns (steering, ind).  ``cx, cy`` are the path columns 1 and 2 (:206-207).
    """
    ind = int(target_ind)
    n = len(cx)
    while ind < n:                                       # :212
        dx = cx[ind] - x                                 # :213
        dy = cy[ind] - y                                 # :214
        distance = np.hypot(dx, dy)                      # :215
        if distance >= Lf:                               # :216
            break
        ind += 1                                         # :218
    if ind >= n:                                         # :219
        ind = n - 1                                      # :220
    tx = cx[ind]
    ty = cy[ind]
    alpha = math.atan2(ty - y, tx - x) - yaw             # :225
    alpha = normalize_angle(alpha)                       # :226
    steering = math.atan2(2.0 * WB * math.sin(alpha) / Lf, 1.0)   # :228
    steering = float(np.clip(steering, -max_steer, max_steer))    # :229
    return steering, ind


# --------------------------------------------------------------------------------------
# core/controllers.py:605-612  find_target_point
# --------------------------------------------------------------------------------------
def find_target_point(x, y, cx, cy, target_ind, lookahead=CONF.LOOKAHEAD_DISTANCE):
    target_ind = int(target_ind)
    while target_ind < len(cx) - 1:                      # :607
        distance = np.hypot(cx[target_ind] - x, cy[target_ind] - y)   # :608
        if distance >= lookahead:                        # :609
            break
        target_ind += 1                                  # :611
    return target_ind


# --------------------------------------------------------------------------------------
# core/controllers.py:353-376  (live StanleyController) -- nearest point + error terms.
# Shared by the Stanley law and by the per-step error logging (ORACLE-DEFINED use).
# --------------------------------------------------------------------------------------
def path_heading(cx, cy, ind):
    """Forward difference, backward at the last index (:361-368)."""
    n = len(cx)
    if ind < n - 1:                                      # :361
        path_dx = cx[ind + 1] - cx[ind]                  # :362
        path_dy = cy[ind + 1] - cy[ind]                  # :363
    else:
        path_dx = cx[ind] - cx[ind - 1]                  # :365
        path_dy = cy[ind] - cy[ind - 1]                  # :366
    return math.atan2(path_dy, path_dx)                  # :368


def nearest_point_errors(x, y, yaw, cx, cy):
    """Global argmin + heading error + cross-track error (:353-376).

    Returns (ind, theta_e, e_ct, path_yaw).
    """
    dx = cx - x                                          # :353
    dy = cy - y                                          # :354
    d = np.hypot(dx, dy)                                 # :355
    ind = int(np.argmin(d))                              # :356 (first minimum wins)
    path_yaw = path_heading(cx, cy, ind)
    theta_e = normalize_angle(path_yaw - yaw)            # :371
    ddx = cx[ind] - x                                    # :374
    ddy = cy[ind] - y                                    # :375
    e_ct = float(-ddx * np.sin(path_yaw) + ddy * np.cos(path_yaw))   # :376
    return ind, float(theta_e), e_ct, path_yaw


# --------------------------------------------------------------------------------------
# core/controllers.py:336-385  StanleyController.compute_steering (the LIVE definition)
# --------------------------------------------------------------------------------------
def stanley_steering(x, y, yaw, v, cx, cy, k=1.0, k_soft=0.1, max_steer=CONF.MAX_STEER):
    """Returns (steering, ind).  Ignores any incoming target index (:356 overwrites it)."""
    ind, theta_e, e_ct, _ = nearest_point_errors(x, y, yaw, cx, cy)
    steering = theta_e + math.atan2(k * e_ct, v + k_soft)             # :379
    steering = normalize_angle(steering)                              # :380
    steering = float(np.clip(steering, -max_steer, max_steer))        # :381
    return steering, ind


# --------------------------------------------------------------------------------------
# core/controllers.py:234-324  the SHADOWED (dead-code) Stanley variant: adaptive gain, steering-rate
# limit, low-pass filter, carried ``previous_steering`` (stored BEFORE the final clip, :319-322).
# Labelled extra (SURVEY.md section 8 row a2'); pinned by tests/golden/shadowed_stanley_kat.json, which
# oracle/gen_golden_shadowed.py produces by executing the reference's own source for these lines.
# ``lookahead_distance`` (:276-281): the optional advance of the index from the nearest point (default 0.0 = off).
# --------------------------------------------------------------------------------------
SHADOWED_RATE = float(np.deg2rad(30))      # max_steer_rate default (:234)
SHADOWED_ALPHA = 0.5                       # alpha default (:234)


def stanley_shadowed_steering(x, y, yaw, v, cx, cy, previous_steering, dt, k=CONF.k_stanley,
                              k_soft=CONF.k_soft_stanley, max_steer_rate=SHADOWED_RATE, alpha=SHADOWED_ALPHA,
                              max_steer=CONF.MAX_STEER, lookahead_distance=0.0):
    """Returns (steering, ind, new_previous_steering)."""
    ind, theta_e, e_ct, _ = nearest_point_errors(x, y, yaw, cx, cy)             # :269-273, :283-302
    if lookahead_distance > 0:                                                   # :276
        while ind < len(cx) - 1:                                                 # :277
            distance = np.hypot(cx[ind] - x, cy[ind] - y)                        # :278
            if distance >= lookahead_distance:                                   # :279
                break
            ind += 1                                                             # :281
        path_yaw = path_heading(cx, cy, ind)                                     # :286-294
        theta_e = float(normalize_angle(path_yaw - yaw))                         # :297
        ddx = cx[ind] - x                                                        # :300
        ddy = cy[ind] - y                                                        # :301
        e_ct = float(-ddx * np.sin(path_yaw) + ddy * np.cos(path_yaw))           # :302
    adaptive_k = k / (v + k_soft)                                                # :305
    steering = theta_e + math.atan2(adaptive_k * e_ct, 1.0)                      # :308
    steering = normalize_angle(steering)                                         # :309
    max_delta = max_steer_rate * dt                                              # :312
    steering = np.clip(steering, previous_steering - max_delta, previous_steering + max_delta)   # :313
    steering = alpha * steering + (1 - alpha) * previous_steering                # :316
    new_prev = float(steering)                                                   # :319
    steering = float(np.clip(steering, -max_steer, max_steer))                   # :322
    return steering, ind, new_prev


# --------------------------------------------------------------------------------------
# core/controllers.py:58-73 and :165-180  curvature -> desired speed
# --------------------------------------------------------------------------------------
def desired_speed(curvature, target_speed=CONF.TARGET_SPEED, k_expo=CONF.k_expo):
    return float(target_speed * np.exp(-k_expo * np.abs(curvature)))  # :72 / :179


# --------------------------------------------------------------------------------------
# core/controllers.py:18-47  AccelerationPIDController (+ restated simple_pid, fixed dt)
# --------------------------------------------------------------------------------------
@dataclass
class PIDState:
    integral: float = 0.0
    last_input: float | None = None


def pid_acceleration(st: PIDState, current_speed, curvature, kp, ki, kd, dt,
                     target_speed=CONF.TARGET_SPEED, k_expo=CONF.k_expo,
                     max_decel=CONF.MAX_DECEL, max_accel=CONF.MAX_ACCEL):
    """compute_acceleration (:32-47).  PID arithmetic = oracle.ref_shim.FixedDtPID
    (restated simple-pid, PARITY UNPINNED).  Returns (acceleration, desired_v)."""
    desired_v = desired_speed(curvature, target_speed, k_expo)        # :42
    setpoint = desired_v                                              # :43-44
    error = setpoint - current_speed
    d_input = current_speed - (st.last_input if st.last_input is not None else current_speed)
    proportional = kp * error
    st.integral += ki * error * dt
    derivative = -kd * d_input / dt
    st.last_input = current_speed
    acceleration = proportional + st.integral + derivative            # :45
    acceleration = float(np.clip(acceleration, max_decel, max_accel))  # :46
    return acceleration, desired_v


# --------------------------------------------------------------------------------------
# core/controllers.py:78-162  LQGAccelerationController
# --------------------------------------------------------------------------------------
def lqg_matrices(dt):
    A = np.array([[1, 0], [dt, 1]], dtype=float)                      # :83-84
    B = np.array([[dt], [0]], dtype=float)                            # :85-86
    C = np.array([[1, 0]], dtype=float)                               # :87
    Q = np.diag([10.0, 100.0])                                        # :90-92
    R = np.array([[0.001]])                                           # :93-94
    Q_kf = np.diag([0.1, 0.1])                                        # :101-103
    R_kf = np.array([[0.5]])                                          # :104
    return A, B, C, Q, R, Q_kf, R_kf


def lqr_gain(dt):
    """``control.dlqr`` restated with scipy's DARE (:97).  PARITY UNPINNED (python-control absent)."""
    from scipy.linalg import solve_discrete_are

    A, B, _, Q, R, _, _ = lqg_matrices(dt)
    P = solve_discrete_are(A, B, Q, R)
    return np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A)              # shape (1, 2)


@dataclass
class LQGState:
    x_hat: np.ndarray = field(default_factory=lambda: np.zeros((2, 1)))   # :107-108
    P: np.ndarray = field(default_factory=lambda: np.eye(2) * 1.0)        # :109
    a_k_prev: float = 0.0                                                 # :112


def kalman_filter_update(A, B, C, Q_kf, R_kf, x_hat, P, a_k_prev, z_k):
    """:142-162, same operation order."""
    x_pred = A @ x_hat + B * a_k_prev                                 # :144
    P_pred = A @ P @ A.T + Q_kf                                       # :145
    y = z_k - (C @ x_pred)                                            # :148
    S = C @ P_pred @ C.T + R_kf                                       # :151
    K_kf = P_pred @ C.T @ np.linalg.inv(S)                            # :154
    x_hat_new = x_pred + K_kf @ y                                     # :157
    P_new = (np.eye(A.shape[0]) - K_kf @ C) @ P_pred                  # :160
    return x_hat_new, P_new, K_kf


def kalman_gain_sequence(dt, n_steps):
    """The covariance recursion (:145,151,154,160) is data independent: the Kalman gain at
    step k is the same for every episode.  Returns (n_steps, 2) float64 -- the table the
    GPU path shares between all episodes (SURVEY.md section 7 'Register pressure')."""
    A, B, C, _, _, Q_kf, R_kf = lqg_matrices(dt)
    P = np.eye(2) * 1.0
    out = np.zeros((n_steps, 2))
    for k in range(n_steps):
        P_pred = A @ P @ A.T + Q_kf
        S = C @ P_pred @ C.T + R_kf
        K_kf = P_pred @ C.T @ np.linalg.inv(S)
        P = (np.eye(2) - K_kf @ C) @ P_pred
        out[k] = K_kf[:, 0]
    return out


def lqg_acceleration(st: LQGState, K, mats, current_speed, curvature,
                     target_speed=CONF.TARGET_SPEED, k_expo=CONF.k_expo,
                     max_decel=CONF.MAX_DECEL, max_accel=CONF.MAX_ACCEL):
    """compute_acceleration (:117-138).  Returns (acceleration, desired_v)."""
    A, B, C, _, _, Q_kf, R_kf = mats
    v_desired = desired_speed(curvature, target_speed, k_expo)        # :119
    e_v_measurement = current_speed - v_desired                       # :123
    st.x_hat, st.P, _ = kalman_filter_update(A, B, C, Q_kf, R_kf, st.x_hat, st.P,
                                             st.a_k_prev, e_v_measurement)   # :126-127
    a_k = -K @ st.x_hat                                               # :130
    a_k = np.clip(a_k, max_decel, max_accel)                          # :133
    st.a_k_prev = a_k.item()                                          # :136
    return a_k.item(), v_desired


# --------------------------------------------------------------------------------------
# core/data/car_state.py:178-223  DynamicBicycleModel.state_equations
# (State.update(a, delta) calls it with u = [delta, a], :118-125)
# --------------------------------------------------------------------------------------
def dynamic_bicycle_step(state6, delta, a, dt, m=CONF.m, I_z=CONF.I_z, l_f=CONF.l_f, l_r=CONF.l_r,
                         c_f=CONF.c_f, c_r=CONF.c_r, mu=CONF.mu):
    """state6 = [x, y, yaw, v, r, beta] -> next state (list of 6 floats)."""
    x_pos, y_pos, yaw, v, r, beta = state6
    epsilon = 1e-5                                                    # :203
    v = max(v, epsilon)                                               # :204
    F_yf = -c_f * (beta + (l_f * r) / v - delta)                      # :207
    F_yr = -c_r * (beta - (l_r * r) / v)                              # :208
    F_yf = float(np.clip(F_yf, -mu * m * 9.81, mu * m * 9.81))        # :211
    F_yr = float(np.clip(F_yr, -mu * m * 9.81, mu * m * 9.81))        # :212
    x_pos_next = x_pos + v * np.cos(yaw + beta) * dt                  # :215
    y_pos_next = y_pos + v * np.sin(yaw + beta) * dt                  # :216
    yaw_next = yaw + r * dt                                           # :217
    v_next = v + a * dt                                               # :218
    r_next = r + (l_f * F_yf - l_r * F_yr) / I_z * dt                 # :219
    beta_next = beta + (F_yr / (m * v) - r) * dt                      # :220
    return [float(x_pos_next), float(y_pos_next), float(yaw_next), float(v_next),
            float(r_next), float(beta_next)]


# --------------------------------------------------------------------------------------
# ORACLE-DEFINED: kinematic bicycle step (no reference code).
# Standard rear-axle explicit Euler with WB = conf.WB (vehicle_config.py:16):
#   x += v cos(yaw) dt;  y += v sin(yaw) dt;  yaw += v / WB * tan(delta) dt;  v += a dt
# All right-hand sides use the OLD state.  r and beta pass through unchanged.
# --------------------------------------------------------------------------------------
def kinematic_bicycle_step(state6, delta, a, dt, WB=CONF.WB):
    x_pos, y_pos, yaw, v, r, beta = state6
    x_next = x_pos + v * math.cos(yaw) * dt
    y_next = y_pos + v * math.sin(yaw) * dt
    yaw_next = yaw + v / WB * math.tan(delta) * dt
    v_next = v + a * dt
    return [x_next, y_next, yaw_next, v_next, r, beta]


# --------------------------------------------------------------------------------------
# core/controllers.py:486-568  MPCController -- cost / linearised dynamics of ONE candidate
# control sequence (the QP optimum itself is PARITY UNPINNED: cvxpy/OSQP absent).
# --------------------------------------------------------------------------------------
MPC_Q = (1.0, 1.0, 0.5, 0.1)                                          # :483
MPC_R = (0.1, 0.1)                                                    # :484


def mpc_candidate_cost(x0, y0, yaw0, v0, ref_x, ref_y, u, dt, WB=CONF.WB,
                       target_speed=CONF.TARGET_SPEED):
    """Cost of control sequence ``u`` (H, 2) = [accel, steer] from state (x0,y0,yaw0,v0).

    Dynamics :522-527 (x, y advance with the INITIAL speed and heading; yaw and v integrate
    the controls), reference state :539-542 (indexed by stage k, yaw reference 0.0, clamped to
    the last path point), stage cost :545, terminal cost with the LAST stage's ref_state :548.
    Returns (cost, feasible) where feasible = all(v_{k+1} >= 0) (:536); control bounds
    (:530-533) are the candidate generator's responsibility.
    """
    H = len(u)
    cos_t = float(np.cos(yaw0))                                       # :514
    sin_t = float(np.sin(yaw0))                                       # :515
    sx, sy, syaw, sv = x0, y0, yaw0, v0                               # :511
    cost = 0.0
    feasible = True
    n_ref = len(ref_x)
    rx = ry = 0.0
    for k in range(H):
        if k < n_ref:                                                 # :539
            rx, ry = ref_x[k], ref_y[k]                               # :540
        else:
            rx, ry = ref_x[-1], ref_y[-1]                             # :542
        ex, ey, eyaw, ev = sx - rx, sy - ry, syaw - 0.0, sv - target_speed
        cost += (MPC_Q[0] * ex * ex + MPC_Q[1] * ey * ey + MPC_Q[2] * eyaw * eyaw
                 + MPC_Q[3] * ev * ev)                                # :545 quad_form(x - ref, Q)
        cost += MPC_R[0] * u[k][0] * u[k][0] + MPC_R[1] * u[k][1] * u[k][1]   # :545 quad_form(u, R)
        nx = sx + v0 * cos_t * dt                                     # :523
        ny = sy + v0 * sin_t * dt                                     # :524
        nyaw = syaw + v0 / WB * u[k][1] * dt                          # :525
        nv = sv + u[k][0] * dt                                        # :526
        sx, sy, syaw, sv = nx, ny, nyaw, nv
        if sv < 0.0:                                                  # :536
            feasible = False
    ex, ey, eyaw, ev = sx - rx, sy - ry, syaw - 0.0, sv - target_speed
    cost += (MPC_Q[0] * ex * ex + MPC_Q[1] * ey * ey + MPC_Q[2] * eyaw * eyaw
             + MPC_Q[3] * ev * ev)                                    # :548
    return cost, feasible


def mpc_costs_batch(x0, y0, yaw0, v0, ref_x, ref_y, bank, dt, WB=CONF.WB,
                    target_speed=CONF.TARGET_SPEED):
    """Vectorised ``mpc_candidate_cost`` over a candidate bank (K, H, 2); same stage order.

    Returns (costs (K,), feasible (K,) bool)."""
    bank = np.asarray(bank, dtype=np.float64)
    K, H, _ = bank.shape
    cos_t = float(np.cos(yaw0))
    sin_t = float(np.sin(yaw0))
    n_ref = len(ref_x)
    sx = x0
    sy = y0
    syaw = np.full(K, float(yaw0))
    sv = np.full(K, float(v0))
    cost = np.zeros(K)
    feas = np.ones(K, dtype=bool)
    rx = ry = 0.0
    for k in range(H):
        if k < n_ref:
            rx, ry = ref_x[k], ref_y[k]
        else:
            rx, ry = ref_x[-1], ref_y[-1]
        ex, ey = sx - rx, sy - ry
        ev = sv - target_speed
        cost += MPC_Q[0] * ex * ex + MPC_Q[1] * ey * ey + MPC_Q[2] * syaw * syaw + MPC_Q[3] * ev * ev
        cost += MPC_R[0] * bank[:, k, 0] ** 2 + MPC_R[1] * bank[:, k, 1] ** 2
        sx = sx + v0 * cos_t * dt
        sy = sy + v0 * sin_t * dt
        syaw = syaw + v0 / WB * bank[:, k, 1] * dt
        sv = sv + bank[:, k, 0] * dt
        feas &= sv >= 0.0
    ex, ey = sx - rx, sy - ry
    ev = sv - target_speed
    cost += MPC_Q[0] * ex * ex + MPC_Q[1] * ey * ey + MPC_Q[2] * syaw * syaw + MPC_Q[3] * ev * ev
    return cost, feas


def mpc_qp_optimum(x0, y0, yaw0, v0, ref_x, ref_y, H, dt, WB=CONF.WB, target_speed=CONF.TARGET_SPEED,
                   max_accel=CONF.MAX_ACCEL, max_steer=CONF.MAX_STEER):
    """The EXACT optimum of the QP that compute_control hands to cvxpy / OSQP (:503-552) -- what the reference's solver
    approximates to its tolerance (cvxpy 1.x -> OSQP, both absent from the reference tree and from this image: the
    solver's own iterate stays PARITY UNPINNED; the optimum it converges to is pinned here).

    Under the linearisation x and y do not depend on the controls (:523-524), yaw is driven by the steering only
    (:525) and v by the acceleration only (:526), and Q, R are diagonal: the QP separates into two bounded
    least-squares chains of H variables,
        min_s  sum_{k=0..H} q2 (yaw0 + v0/WB dt sum_{i<k} s_i)^2 + r1 sum_k s_k^2      |s_k| <= MAX_STEER      (:533)
        min_a  sum_{k=0..H} q3 (v0 - v_t + dt sum_{i<k} a_i)^2  + r0 sum_k a_k^2      |a_k| <= MAX_ACCEL      (:532)
    (``cp.abs(u[k, 0]) <= MAX_ACCEL``: the QP's acceleration bound is symmetric; MAX_DECEL only enters the final clip,
    :565), plus v_k >= 0 (:536), which is checked on the solution: if it binds, the v chain is re-solved with SLSQP.
    Each chain is solved by scipy's bounded-variable least squares (active set: exact up to round-off).
    Returns (cost, accel (H,), steer (H,))."""
    from scipy.optimize import lsq_linear, minimize
    low = np.tril(np.ones((H, H)))                         # row k - 1: sum over i < k, k = 1..H

    def chain(scale, qq, rr, e0, bound):
        a_mat = np.vstack([math.sqrt(qq) * scale * low, math.sqrt(rr) * np.eye(H)])
        b_vec = np.concatenate([-math.sqrt(qq) * e0 * np.ones(H), np.zeros(H)])
        return lsq_linear(a_mat, b_vec, bounds=(-bound, bound), method="bvls", tol=1e-15).x

    steer = chain(v0 / WB * dt, MPC_Q[2], MPC_R[1], yaw0, max_steer)
    accel = chain(dt, MPC_Q[3], MPC_R[0], v0 - target_speed, max_accel)
    if np.any(v0 + dt * np.cumsum(accel) < 0.0):           # :536 binds: general inequality constraints
        def fun(a):
            e = v0 - target_speed + dt * np.concatenate([[0.0], np.cumsum(a)])
            return MPC_Q[3] * float(e @ e) + MPC_R[0] * float(a @ a)
        cons = [{"type": "ineq", "fun": lambda a: v0 + dt * np.cumsum(a)}]
        res = minimize(fun, np.clip(accel, -max_accel, max_accel), method="SLSQP", bounds=[(-max_accel, max_accel)] * H,
                       constraints=cons, options={"ftol": 1e-14, "maxiter": 500})
        accel = res.x
    cost, _ = mpc_candidate_cost(x0, y0, yaw0, v0, ref_x, ref_y, np.stack([accel, steer], axis=1), dt, WB, target_speed)
    return float(cost), accel, steer


def mpc_select(costs, feasible, bank, max_decel=CONF.MAX_DECEL, max_accel=CONF.MAX_ACCEL,
               max_steer=CONF.MAX_STEER):
    """Pick the cheapest feasible candidate (first minimum wins), then apply the return
    convention of compute_control (:565-568): clip, steering sign flipped.
    If no candidate is feasible: the solver-failure fallback (0, -0) (:558-562)."""
    c = np.where(feasible, costs, np.inf)
    best = int(np.argmin(c))
    if not np.isfinite(c[best]):
        return -1, 0.0, -0.0, float("inf")
    acceleration = float(np.clip(bank[best][0][0], max_decel, max_accel))   # :565
    steering = float(np.clip(bank[best][0][1], -max_steer, max_steer))      # :566
    return best, acceleration, -steering, float(c[best])                    # :568


# --------------------------------------------------------------------------------------
# preformance_analysis/metrics_calculator.py:1-8 on the CSV schema of
# preformance_analysis/simulation_logger.py:8-10  (pandas .std() => ddof = 1)
# --------------------------------------------------------------------------------------
def lap_metrics(timestamp, lateral_error, heading_error):
    """Returns dict with the five reference scalars.  numpy restatement of the pandas calls:
    Series.mean(), Series.std() (ddof=1; NaN for a single row), iloc[-1] - iloc[0]."""
    lat = np.asarray(lateral_error, dtype=np.float64)
    hed = np.asarray(heading_error, dtype=np.float64)
    ts = np.asarray(timestamp, dtype=np.float64)
    n = len(lat)

    def _std(a):
        return float(np.std(a, ddof=1)) if n > 1 else float("nan")

    return {
        "lateral_mean": float(np.mean(lat)), "lateral_std": _std(lat),       # :1-2
        "heading_mean": float(np.mean(hed)), "heading_std": _std(hed),       # :4-5
        "lap_time": float(ts[-1] - ts[0]),                                   # :7-8
    }


# --------------------------------------------------------------------------------------
# old/test_runs_system_id.py:84-111  identify_dynamics: LinearRegression (with intercept) of the next
# state on hstack([state, input]) -- restated on arrays instead of a CSV file.  The reference logs
# [x, y, yaw, v, yaw_rate, beta, throttle, steering]; here the inputs are the logged (accel, steering).
# --------------------------------------------------------------------------------------
def identify_dynamics(states, inputs):
    """states (T, 6), inputs (T, 2) -> (A (6, 6), B (6, 2), intercept (6,), C, D) as the reference returns them."""
    from sklearn.linear_model import LinearRegression

    states = np.asarray(states, dtype=np.float64)
    inputs = np.asarray(inputs, dtype=np.float64)
    X = states[:-1]                                                   # :97
    U = inputs[:-1]                                                   # :98
    Y = states[1:]                                                    # :99
    reg = LinearRegression()                                          # :101
    reg.fit(np.hstack([X, U]), Y)                                     # :102
    AB = reg.coef_                                                    # :103
    A = AB[:, :X.shape[1]]                                            # :104
    B = AB[:, X.shape[1]:]                                            # :105
    C = np.eye(states.shape[1])                                       # :106
    D = np.zeros((states.shape[1], inputs.shape[1]))                  # :107
    return A, B, reg.intercept_, C, D
