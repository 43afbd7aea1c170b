"""Freeze golden vectors from the UNMODIFIED reference (run in the build container only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Usage:  python -m oracle.gen_golden
Writes tests/golden/reference_kat.json.  Every value in that file is an output of the
reference's own bytes under /root/reference (imported through oracle/ref_shim.py) on the seeded
inputs recorded next to it; floats are stored as hex strings (float.hex) so they round-trip exactly.

The MPC section runs the reference's ``MPCController.compute_control`` (core/controllers.py:486-568)
against an EAGER stand-in for cvxpy (``_EagerCvxpy`` below): variables carry concrete candidate
values, ``quad_form`` / ``abs`` / comparisons evaluate numerically, and ``Problem.solve`` records
the objective value and the residual of every constraint instead of optimising.  This pins the
reference's cost and dynamics DEFINITIONS (not the OSQP optimum, which stays parity-unpinned).
"""
from __future__ import annotations

import json
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "reference_kat.json")


def hx(v):
    """Exact, JSON-safe encoding of floats / arrays of floats."""
    if isinstance(v, (list, tuple, np.ndarray)):
        return [hx(e) for e in v]
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    return float(v).hex()


# ------------------------------------------------------------------ eager cvxpy stand-in
class _Expr:
    def __init__(self, v):
        self.v = np.asarray(v, dtype=np.float64)

    @staticmethod
    def _val(o):
        return o.v if isinstance(o, _Expr) else np.asarray(o, dtype=np.float64)

    def __add__(self, o): return _Expr(self.v + self._val(o))
    __radd__ = __add__
    def __sub__(self, o): return _Expr(self.v - self._val(o))
    def __rsub__(self, o): return _Expr(self._val(o) - self.v)
    def __mul__(self, o): return _Expr(self.v * self._val(o))
    __rmul__ = __mul__
    def __truediv__(self, o): return _Expr(self.v / self._val(o))
    def __rtruediv__(self, o): return _Expr(self._val(o) / self.v)
    def __neg__(self): return _Expr(-self.v)
    def __getitem__(self, idx): return _Expr(self.v[idx])
    def __eq__(self, o): return ("eq", float(np.max(np.abs(self.v - self._val(o)))))
    def __le__(self, o): return ("le", float(np.max(self.v - self._val(o))))
    def __ge__(self, o): return ("ge", float(np.max(self._val(o) - self.v)))
    __hash__ = None


class _EagerCvxpy(types.ModuleType):
    OPTIMAL = "optimal"
    OPTIMAL_INACCURATE = "optimal_inaccurate"
    OSQP = "OSQP"

    def __init__(self):
        super().__init__("cvxpy")
        self.seed_values = []     # values handed to successive Variable() calls
        self.captured = None

    def Variable(self, shape):
        val = self.seed_values.pop(0)
        assert tuple(val.shape) == tuple(shape), (val.shape, shape)
        e = _Expr(val)
        e.value = np.asarray(val, dtype=np.float64)
        return e

    @staticmethod
    def quad_form(x, Q):
        v = _Expr._val(x)
        return _Expr(v @ np.asarray(Q) @ v)

    @staticmethod
    def abs(x):
        return _Expr(np.abs(_Expr._val(x)))

    @staticmethod
    def Minimize(cost):
        return cost

    def Problem(self, objective, constraints):
        mod = self

        class _P:
            status = None

            def solve(self_inner, **kw):
                mod.captured = {"cost": float(_Expr._val(objective)), "constraints": list(constraints)}
                self_inner.status = mod.OPTIMAL

        return _P()


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shim

    eager = _EagerCvxpy()
    sys.modules["cvxpy"] = eager            # before the reference is imported
    ref = ref_shim.load_reference()
    C, CS, conf = ref.controllers, ref.car_state, ref.conf
    assert C.cp is eager

    G = {"_meta": {"generator": "oracle/gen_golden.py", "reference_root": ref_shim.REFERENCE_ROOT,
                   "numpy": np.__version__, "python": sys.version.split()[0],
                   "note": "floats are float.hex(); produced by the reference's unmodified code"}}

    # ---- Vehicle_config (vehicle_config.py:4-39)
    G["vehicle_config"] = {k: hx(getattr(conf, k)) for k in (
        "m", "I_z", "l_f", "l_r", "c_f", "c_r", "mu", "TARGET_SPEED", "WB", "kp_accel", "ki_accel",
        "kd_accel", "k_expo", "k_stanley", "k_soft_stanley", "MAX_ACCEL", "MAX_DECEL", "MAX_SPEED",
        "MAX_STEER", "LOOKAHEAD_DISTANCE", "dt")}

    # ---- normalize_angle (core/controllers.py:592-603)
    rng = np.random.default_rng(1001)
    angles = [0.0, 3.5, -3.5, math.pi, -math.pi, 2 * math.pi, -2 * math.pi, 1e-9, -1e-9, 24.7, -31.2,
              np.nextafter(math.pi, 0.0), np.nextafter(math.pi, 4.0)] + list(rng.uniform(-40, 40, 64))
    G["normalize_angle"] = {"in": hx(angles), "out": hx([C.normalize_angle(a) for a in angles])}

    # ---- paths used below
    def ellipse(n, a, b):
        th = np.linspace(0, 2 * np.pi, n, endpoint=False)
        return a * np.cos(th), b * np.sin(th)

    paths = {"ellipse720": ellipse(720, 40.0, 20.0), "ellipse400": ellipse(400, 30.0, 15.0)}
    G["paths"] = {"ellipse720": {"n": 720, "a": hx(40.0), "b": hx(20.0)},
                  "ellipse400": {"n": 400, "a": hx(30.0), "b": hx(15.0)},
                  "recipe": "th=linspace(0,2pi,n,endpoint=False); cx=a*cos(th); cy=b*sin(th)"}

    def rand_states_near(cx, cy, rng, n):
        out = []
        for _ in range(n):
            i = int(rng.integers(0, len(cx)))
            j = (i + 1) % len(cx)
            hdg = math.atan2(cy[j] - cy[i], cx[j] - cx[i])
            out.append((float(cx[i] + rng.normal(0, 0.6)), float(cy[i] + rng.normal(0, 0.6)),
                        float(hdg + rng.normal(0, 0.25) + 2 * math.pi * rng.integers(-2, 3)),
                        float(rng.uniform(0.0, 9.0)), i))
        return out

    # ---- PurePursuit / Stanley / find_target_point (core/controllers.py:193, :336, :605)
    for pname, (cx, cy) in paths.items():
        path = np.column_stack((np.arange(len(cx)), cx, cy))          # nodes/controller.py:56
        rng = np.random.default_rng(2002 + len(cx))
        cases = rand_states_near(cx, cy, rng, 48)
        # SURVEY starter vectors
        if pname == "ellipse720":
            cases.insert(0, (40.3, -0.2, 1.5, 5.0, 0))
        else:
            cases.insert(0, (30.0, 0.0, math.pi / 2, 5.0, 0))
        pp_rows, st_rows, ft_rows = [], [], []
        for (x, y, yaw, v, i) in cases:
            st = CS.State(x=x, y=y, yaw=yaw, v=v)
            for Lf in (conf.LOOKAHEAD_DISTANCE, 2.0, 9.0):
                for start in (0, max(0, i - 5), i):
                    s, ind = C.PurePursuitController(look_ahead_distance=Lf).compute_steering(st, path, start)
                    pp_rows.append({"state": hx([x, y, yaw, v]), "Lf": hx(Lf), "start": int(start),
                                    "steering": hx(s), "ind": int(ind)})
            for (k, ks) in ((1.0, 0.1), (0.7, 0.01), (2.0, 1.0)):
                s, ind = C.StanleyController(k=k, k_soft=ks).compute_steering(st, path, 0)
                st_rows.append({"state": hx([x, y, yaw, v]), "k": hx(k), "k_soft": hx(ks),
                                "steering": hx(s), "ind": int(ind)})
            for start in (0, max(0, i - 5)):
                ft_rows.append({"state": hx([x, y, yaw, v]), "start": int(start),
                                "ind": int(C.find_target_point(st, cx, cy, start))})
        G[f"pure_pursuit/{pname}"] = pp_rows
        G[f"stanley/{pname}"] = st_rows
        G[f"find_target_point/{pname}"] = ft_rows
    # end-of-path clamp (core/controllers.py:219-220)
    cx, cy = paths["ellipse400"]
    path = np.column_stack((np.arange(50), cx[:50], cy[:50]))
    st = CS.State(x=float(cx[48]), y=float(cy[48]), yaw=2.0, v=3.0)
    s, ind = C.PurePursuitController().compute_steering(st, path, 45)
    G["pure_pursuit/clamp"] = {"n": 50, "state": hx([st.x, st.y, st.yaw, st.v]), "start": 45,
                               "steering": hx(s), "ind": int(ind)}

    # ---- speed target (core/controllers.py:58-73) and PID law (:32-47) with the fixed-dt PID
    kap = [0.0, 0.08, -0.08, 1 / 15, 0.5, -2.0]
    pidc = C.AccelerationPIDController(conf.kp_accel, conf.ki_accel, conf.kd_accel, conf.TARGET_SPEED)
    G["compute_breaking"] = {"kappa": hx(kap), "v": hx([pidc.compute_breaking(k) for k in kap])}
    rng = np.random.default_rng(3003)
    seq_v = list(np.abs(np.cumsum(rng.normal(0.15, 0.2, 60))))
    seq_k = list(rng.choice([0.0, 1 / 15, -1 / 15, 0.2], 60))
    ref_shim.FixedDtPID.dt = conf.dt
    pidc = C.AccelerationPIDController(conf.kp_accel, conf.ki_accel, conf.kd_accel, conf.TARGET_SPEED)
    acc, msp = [], []
    for v, k in zip(seq_v, seq_k):
        acc.append(pidc.compute_acceleration(v, k))
        msp.append(pidc.maxspeed)
    G["pid_sequence"] = {"dt": hx(conf.dt), "gains": hx([conf.kp_accel, conf.ki_accel, conf.kd_accel]),
                         "v": hx(seq_v), "kappa": hx(seq_k), "accel": hx(acc), "maxspeed": hx(msp),
                         "note": "PID arithmetic = restated simple_pid with fixed dt (parity unpinned)"}

    # ---- LQG (core/controllers.py:78-162) with the scipy-backed dlqr
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        lqg = C.LQGAccelerationController(conf.dt)
    G["lqg_K"] = hx(np.asarray(lqg.K).ravel())
    acc, xh, Pm, vd = [], [], [], []
    for v, k in zip(seq_v, seq_k):
        acc.append(lqg.compute_acceleration(v, k))
        xh.append(np.asarray(lqg.x_hat).ravel().tolist())
        Pm.append(np.asarray(lqg.P).ravel().tolist())
        vd.append(lqg.v_desired)
    G["lqg_sequence"] = {"dt": hx(conf.dt), "v": hx(seq_v), "kappa": hx(seq_k), "accel": hx(acc),
                         "x_hat": hx(xh), "P": hx(Pm), "v_desired": hx(vd),
                         "note": "K from scipy DARE (control.dlqr absent: parity unpinned)"}

    # ---- DynamicBicycleModel.state_equations (core/data/car_state.py:178-223), State.update (:118-125)
    rng = np.random.default_rng(4004)
    model = CS.DynamicBicycleModel()
    rows = [([1, 2, 0.3, 6, 0.2, 0.05], [0.15, 1.0]), ([0, 0, 0, 0, 0, 0], [0.3, 1.5])]
    for _ in range(64):
        rows.append(([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-30, 30),
                      rng.choice([0.0, 1e-6, rng.uniform(0, 12)]), rng.normal(0, 0.6), rng.normal(0, 0.15)],
                     [rng.uniform(-0.55, 0.55), rng.uniform(-2.0, 1.5)]))
    G["state_equations"] = [{"x": hx(x), "u": hx(u), "dt": hx(conf.dt),
                             "out": hx(model.state_equations(list(map(float, x)), list(map(float, u)), conf.dt))}
                            for x, u in rows]
    # low-grip variants exercise the friction clip (:211-212)
    low = CS.DynamicBicycleModel(mu=0.3, c_f=20000.0, c_r=15000.0)
    G["state_equations_lowmu"] = [{"x": hx(x), "u": hx(u), "dt": hx(conf.dt), "mu": hx(0.3), "c_f": hx(20000.0),
                                   "c_r": hx(15000.0),
                                   "out": hx(low.state_equations(list(map(float, x)), list(map(float, u)), conf.dt))}
                                  for x, u in rows[:24]]
    st = CS.State(0, 0, 0, 5)
    st.update(0.5, 0.1)                      # NB argument order (a, delta)
    G["state_update"] = {"init": hx([0, 0, 0, 5, 0, 0]), "a": hx(0.5), "delta": hx(0.1),
                         "out": hx([st.x, st.y, st.yaw, st.v, st.r, st.beta]),
                         "axles": hx([st.rear_x, st.rear_y, st.front_x, st.front_y])}

    # ---- metrics_calculator (preformance_analysis/metrics_calculator.py:1-8)
    import pandas as pd
    rng = np.random.default_rng(5005)
    mrows = []
    for n in (3, 2, 500, 2000):
        lat = [0.1, 0.2, 0.4] if n == 3 else list(rng.normal(0.05, 0.2, n))
        hed = list(rng.normal(-0.01, 0.05, n))
        ts = list(np.arange(n) * conf.dt + 12.5)
        df = pd.DataFrame({"timestamp": ts, "lateral_error": lat, "heading_error": hed})
        M = ref.metrics_calculator
        lm, ls = M.calculate_lateral_error(df)
        hm, hs = M.calculate_heading_error(df)
        mrows.append({"timestamp": hx(ts), "lateral_error": hx(lat), "heading_error": hx(hed),
                      "lateral_mean": hx(lm), "lateral_std": hx(ls), "heading_mean": hx(hm),
                      "heading_std": hx(hs), "lap_time": hx(M.calculate_lap_time(df))})
    G["metrics"] = mrows

    # ---- composed closed loops out of reference objects (order: nodes/controller.py:61-63)
    def closed_loop(steer_name, accel_name, n_steps, cx, cy, curve, init):
        path = np.column_stack((np.arange(len(cx)), cx, cy))
        steer = C.get_steering_controller(steer_name)
        ref_shim.FixedDtPID.dt = conf.dt
        if accel_name == "pid":
            accel = C.AccelerationPIDController(conf.kp_accel, conf.ki_accel, conf.kd_accel, conf.TARGET_SPEED)
        else:
            with contextlib.redirect_stdout(io.StringIO()):
                accel = C.LQGAccelerationController(conf.dt)
        st = CS.State(*init)
        ti = 0
        traj, steers, accels, tis = [], [], [], []
        for _ in range(n_steps):
            traj.append([st.x, st.y, st.yaw, st.v, st.r, st.beta])
            s, ti = steer.compute_steering(st, path, ti)
            kappa = curve[ti] if ti < len(curve) else curve[-1]
            a = accel.compute_acceleration(st.v, kappa)
            st.update(a, s)
            steers.append(s); accels.append(a); tis.append(int(ti))
            if ti >= len(cx) - 1:
                break
        return {"traj": hx(traj), "steer": hx(steers), "accel": hx(accels), "ti": tis,
                "final": hx([st.x, st.y, st.yaw, st.v, st.r, st.beta])}

    cx, cy = paths["ellipse720"]
    # analytic curvature of the ellipse a=40, b=20
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    G["loop_curve_recipe"] = "kappa = a*b / ((a sin th)^2 + (b cos th)^2)^1.5, a=40, b=20"
    init = (40.0, 0.0, math.pi / 2, 0.0, 0.0, 0.0)
    G["loop/pp_pid_dynamic"] = dict(closed_loop("pure_pursuit", "pid", 400, cx, cy, curve, init), init=hx(init))
    G["loop/stanley_lqg_dynamic"] = dict(closed_loop("stanley", "lqg", 400, cx, cy, curve, init), init=hx(init))
    G["loop/stanley_pid_dynamic"] = dict(closed_loop("stanley", "pid", 400, cx, cy, curve, init), init=hx(init))

    # ---- MPC cost / dynamics definitions through the eager cvxpy stand-in
    from oracle import laws
    rng = np.random.default_rng(6006)
    mpc_rows = []
    cx, cy = paths["ellipse400"]
    for case in range(12):
        N = 10 if case < 8 else 30
        dt = 0.05
        i0 = int(rng.integers(0, 300))
        npts = 60 if case % 4 else 6                # short path exercises the k >= len(ref) clamp (:541-542)
        path = np.column_stack((np.arange(npts), cx[i0:i0 + npts], cy[i0:i0 + npts]))
        x0, y0 = float(path[0, 1] + rng.normal(0, 0.3)), float(path[0, 2] + rng.normal(0, 0.3))
        yaw0, v0 = float(rng.uniform(-3, 3)), float(rng.uniform(0.5, 8))
        u = np.column_stack((rng.uniform(-1.5, 1.5, N), rng.uniform(-conf.MAX_STEER, conf.MAX_STEER, N)))
        # state trajectory under the reference's linearised dynamics (:522-527)
        X = np.zeros((N + 1, 4))
        X[0] = [x0, y0, yaw0, v0]
        for k in range(N):
            X[k + 1, 0] = X[k, 0] + v0 * np.cos(yaw0) * dt
            X[k + 1, 1] = X[k, 1] + v0 * np.sin(yaw0) * dt
            X[k + 1, 2] = X[k, 2] + v0 / conf.WB * u[k, 1] * dt
            X[k + 1, 3] = X[k, 3] + u[k, 0] * dt
        eager.seed_values = [X.copy(), u.copy()]
        eager.captured = None
        ctl = C.MPCController(N=N, dt=dt)
        a_ret, s_ret = ctl.compute_control(CS.State(x=x0, y=y0, yaw=yaw0, v=v0), path)
        cap = eager.captured
        eq_res = max(r for kind, r in cap["constraints"] if kind == "eq")
        ineq = max(r for kind, r in cap["constraints"] if kind != "eq")
        mpc_rows.append({"N": N, "dt": hx(dt), "state": hx([x0, y0, yaw0, v0]), "ref_x": hx(path[:, 1]),
                         "ref_y": hx(path[:, 2]), "u": hx(u), "cost": hx(cap["cost"]),
                         "max_eq_residual": hx(eq_res), "max_ineq_violation": hx(ineq),
                         "ret_accel": hx(a_ret), "ret_steer": hx(s_ret)})
        assert eq_res < 1e-9, eq_res
    G["mpc_cost"] = mpc_rows

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(G, f, indent=0, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
