"""The oracle (oracle/laws.py, oracle/loop.py) against golden vectors frozen from the reference's
own unmodified code (tests/golden/reference_kat.json, made by oracle/gen_golden.py).

Bar: indices exact; floats BIT-EXACT on the generating platform's libm (this container, where the
driver runs the ``not gpu`` suite).  ``ULPS`` of slack is allowed only so that a host whose numpy
SIMD sin/cos/hypot kernels differ in the last place does not fail the suite; it is 0 by default
and can be raised with BGR_GOLDEN_ULPS.
"""
import math
import os

import numpy as np
import pytest

from conftest import unhex
from oracle import laws, loop

ULPS = int(os.environ.get("BGR_GOLDEN_ULPS", "0"))


def same(a, b):
    a, b = float(a), float(b)
    if a == b or (math.isnan(a) and math.isnan(b)):
        return True
    if ULPS == 0:
        return False
    return abs(a - b) <= ULPS * max(np.spacing(abs(a)), np.spacing(abs(b)))


def assert_same(a, b, what=""):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    assert a.shape == b.shape, what
    for i, (u, v) in enumerate(zip(a, b)):
        assert same(u, v), f"{what}[{i}]: oracle {u!r} ({float(u).hex()}) != reference {v!r} ({float(v).hex()})"


def ellipse(n, a, b):
    th = np.linspace(0, 2 * np.pi, n, endpoint=False)
    return a * np.cos(th), b * np.sin(th)


PATHS = {"ellipse720": (720, 40.0, 20.0), "ellipse400": (400, 30.0, 15.0)}


def test_vehicle_config(golden):
    g = {k: unhex(v) for k, v in golden["vehicle_config"].items()}
    c = laws.CONF
    for k, v in g.items():
        assert same(getattr(c, k), v), k


def test_normalize_angle(golden):
    g = golden["normalize_angle"]
    for a, want in zip(unhex(g["in"]), unhex(g["out"])):
        assert same(laws.normalize_angle(a), want), a
    assert laws.normalize_angle(math.pi) == -math.pi      # [-pi, pi): SURVEY a8


@pytest.mark.parametrize("pname", list(PATHS))
def test_pure_pursuit(golden, pname):
    cx, cy = ellipse(*PATHS[pname])
    for row in golden[f"pure_pursuit/{pname}"]:
        x, y, yaw, v = unhex(row["state"])
        s, ind = laws.pure_pursuit_steering(x, y, yaw, cx, cy, row["start"], unhex(row["Lf"]))
        assert ind == row["ind"]
        assert same(s, unhex(row["steering"])), row


def test_pure_pursuit_survey_kat(golden):
    """SURVEY.md section 8c starter vectors."""
    cx, cy = ellipse(720, 40.0, 20.0)
    s, ind = laws.pure_pursuit_steering(40.3, -0.2, 1.5, cx, cy, 0, 5.5)
    assert (s, ind) == (0.2328625291477659, 30)
    cx, cy = ellipse(400, 30.0, 15.0)
    s, ind = laws.pure_pursuit_steering(30.0, 0.0, math.pi / 2, cx, cy, 0, 5.5)
    assert (s, ind) == (0.22090553298549134, 23)
    assert laws.find_target_point(30.0, 0.0, cx, cy, 0) == 23
    s, ind = laws.stanley_steering(30.0, 0.0, math.pi / 2, 5.0, cx, cy)
    assert (s, ind) == (0.01570699444132373, 0)
    cx, cy = ellipse(720, 40.0, 20.0)
    s, ind = laws.stanley_steering(40.3, -0.2, 1.5, 5.0, cx, cy)
    assert (s, ind) == (0.10378436919576073, 719)


def test_pure_pursuit_clamp(golden):
    g = golden["pure_pursuit/clamp"]
    cx, cy = ellipse(400, 30.0, 15.0)
    x, y, yaw, v = unhex(g["state"])
    s, ind = laws.pure_pursuit_steering(x, y, yaw, cx[:g["n"]], cy[:g["n"]], g["start"], laws.CONF.LOOKAHEAD_DISTANCE)
    assert ind == g["ind"] == g["n"] - 1
    assert same(s, unhex(g["steering"]))


@pytest.mark.parametrize("pname", list(PATHS))
def test_stanley(golden, pname):
    cx, cy = ellipse(*PATHS[pname])
    for row in golden[f"stanley/{pname}"]:
        x, y, yaw, v = unhex(row["state"])
        s, ind = laws.stanley_steering(x, y, yaw, v, cx, cy, unhex(row["k"]), unhex(row["k_soft"]))
        assert ind == row["ind"]
        assert same(s, unhex(row["steering"])), row


@pytest.mark.parametrize("pname", list(PATHS))
def test_find_target_point(golden, pname):
    cx, cy = ellipse(*PATHS[pname])
    for row in golden[f"find_target_point/{pname}"]:
        x, y, yaw, v = unhex(row["state"])
        assert laws.find_target_point(x, y, cx, cy, row["start"]) == row["ind"]


def test_desired_speed(golden):
    g = golden["compute_breaking"]
    for k, v in zip(unhex(g["kappa"]), unhex(g["v"])):
        assert same(laws.desired_speed(k), v)
    assert laws.desired_speed(0.08) == 6.9167221482221635


def test_pid_sequence(golden):
    g = golden["pid_sequence"]
    kp, ki, kd = unhex(g["gains"])
    st = laws.PIDState()
    for v, k, a, ms in zip(unhex(g["v"]), unhex(g["kappa"]), unhex(g["accel"]), unhex(g["maxspeed"])):
        acc, des = laws.pid_acceleration(st, v, k, kp, ki, kd, unhex(g["dt"]))
        assert same(acc, a) and same(des, ms)


def test_lqg(golden):
    K = laws.lqr_gain(0.05)
    assert_same(K.ravel(), unhex(golden["lqg_K"]), "K")
    np.testing.assert_allclose(K.ravel(), [22.181042932186813, 56.299306123839216], rtol=1e-12)
    g = golden["lqg_sequence"]
    st = laws.LQGState()
    mats = laws.lqg_matrices(unhex(g["dt"]))
    gains = laws.kalman_gain_sequence(unhex(g["dt"]), len(g["v"]))
    for i, (v, k) in enumerate(zip(unhex(g["v"]), unhex(g["kappa"]))):
        x_prev, a_prev = st.x_hat.copy(), st.a_k_prev
        acc, des = laws.lqg_acceleration(st, K, mats, v, k)
        assert same(acc, unhex(g["accel"][i])), i
        assert same(des, unhex(g["v_desired"][i]))
        assert_same(st.x_hat.ravel(), unhex(g["x_hat"][i]), "x_hat")
        assert_same(st.P.ravel(), unhex(g["P"][i]), "P")
        # the shared (data independent) gain table reproduces the per-episode filter
        dt = unhex(g["dt"])
        xp0 = x_prev[0, 0] + dt * a_prev
        xp1 = dt * x_prev[0, 0] + x_prev[1, 0]
        y = (v - des) - xp0
        np.testing.assert_allclose([xp0 + gains[i, 0] * y, xp1 + gains[i, 1] * y], st.x_hat.ravel(),
                                   rtol=1e-12, atol=1e-14)


def test_state_equations(golden):
    for row in golden["state_equations"]:
        out = laws.dynamic_bicycle_step(unhex(row["x"]), *unhex(row["u"]), unhex(row["dt"]))
        assert_same(out, unhex(row["out"]), "state_equations")
    for row in golden["state_equations_lowmu"]:
        out = laws.dynamic_bicycle_step(unhex(row["x"]), *unhex(row["u"]), unhex(row["dt"]),
                                        mu=unhex(row["mu"]), c_f=unhex(row["c_f"]), c_r=unhex(row["c_r"]))
        assert_same(out, unhex(row["out"]), "state_equations_lowmu")
    # SURVEY 8c starter vectors
    assert laws.dynamic_bicycle_step([1, 2, 0.3, 6, 0.2, 0.05], 0.15, 1.0, 0.05) == [
        1.2818118138542136, 2.1028693422366356, 0.31, 6.05, 0.23864705882352943, 0.02888888888888889]
    assert laws.dynamic_bicycle_step([0, 0, 0, 0, 0, 0], 0.3, 1.5, 0.05) == [
        5.000000000000001e-07, 0.0, 0.0, 0.07501000000000001, 0.06621750000000003, 0.0]


def test_state_update(golden):
    g = golden["state_update"]
    out = laws.dynamic_bicycle_step(unhex(g["init"]), unhex(g["delta"]), unhex(g["a"]), laws.CONF.dt)
    assert_same(out, unhex(g["out"]), "State.update")


def test_metrics(golden):
    for row in golden["metrics"]:
        m = laws.lap_metrics(unhex(row["timestamp"]), unhex(row["lateral_error"]), unhex(row["heading_error"]))
        # pandas' mean/std use different summation orders than numpy: allow a few ulps here
        for k in ("lateral_mean", "lateral_std", "heading_mean", "heading_std"):
            np.testing.assert_allclose(m[k], unhex(row[k]), rtol=1e-13, atol=1e-17, err_msg=k)
        assert same(m["lap_time"], unhex(row["lap_time"]))
    m = laws.lap_metrics([0, 1, 2], [.1, .2, .4], [0, 0, 0])
    np.testing.assert_allclose([m["lateral_mean"], m["lateral_std"]],
                               [0.23333333333333336, 0.1527525231651947], rtol=1e-15)


LOOPS = {"loop/pp_pid_dynamic": ("pure_pursuit", "pid"), "loop/stanley_lqg_dynamic": ("stanley", "lqg"),
         "loop/stanley_pid_dynamic": ("stanley", "pid")}


@pytest.mark.parametrize("name", list(LOOPS))
def test_closed_loop_matches_reference_objects(golden, name):
    """The composed oracle loop reproduces a loop built from the reference's own controller /
    State objects (same order as nodes/controller.py:61-63), bit for bit."""
    g = golden[name]
    steer, accel = LOOPS[name]
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    cx, cy = 40.0 * np.cos(th), 20.0 * np.sin(th)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    track = loop.Track(cx, cy, curve, closed=True)
    want_traj = np.asarray(unhex(g["traj"]))
    spec = loop.RolloutSpec(steer=steer, accel=accel, model="dynamic", max_steps=len(want_traj), laps=1)
    out = loop.rollout_episode(track, spec, loop.default_params(), unhex(g["init"]))
    n = out["steps"]
    assert n == len(want_traj)
    assert list(out["ti"]) == g["ti"]
    assert_same(out["traj"], want_traj, "traj")
    assert_same(out["steer"], unhex(g["steer"]), "steer")
    assert_same(out["accel"], unhex(g["accel"]), "accel")
    assert_same(out["state"], unhex(g["final"]), "final")


def test_mpc_cost_definition(golden):
    """Cost / linearised dynamics as evaluated by the reference's own compute_control through the
    eager cvxpy stand-in (oracle/gen_golden.py)."""
    for row in golden["mpc_cost"]:
        x0, y0, yaw0, v0 = unhex(row["state"])
        u = np.asarray(unhex(row["u"]))
        cost, feas = laws.mpc_candidate_cost(x0, y0, yaw0, v0, unhex(row["ref_x"]), unhex(row["ref_y"]),
                                             u, unhex(row["dt"]))
        np.testing.assert_allclose(cost, unhex(row["cost"]), rtol=1e-13)
        assert feas == (unhex(row["max_ineq_violation"]) <= 0.0)
        cb, fb = laws.mpc_costs_batch(x0, y0, yaw0, v0, unhex(row["ref_x"]), unhex(row["ref_y"]),
                                      u[None], unhex(row["dt"]))
        np.testing.assert_allclose(cb[0], cost, rtol=1e-13)
        # return convention (:565-568): (clip(a), -clip(steer)) of the first control
        best, a, s, c = laws.mpc_select(cb, fb, u[None])
        if feas:
            assert same(a, unhex(row["ret_accel"])) and same(s, unhex(row["ret_steer"]))


def test_shadowed_stanley_variant():
    """core/controllers.py:234-324 (dead code, executed from the reference's own source via ast by
    oracle/gen_golden_shadowed.py) vs the oracle restatement: single calls and a 400-step closed loop."""
    import json
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "shadowed_stanley_kat.json")) as f:
        g = json.load(f)
    d = {k: unhex(v) for k, v in g["defaults"].items()}
    assert (d["k"], d["k_soft"], d["alpha"], d["lookahead_distance"]) == (laws.CONF.k_stanley, laws.CONF.k_soft_stanley,
                                                                         laws.SHADOWED_ALPHA, 0.0)
    assert same(d["max_steer_rate"], laws.SHADOWED_RATE)
    cx, cy = ellipse(720, 40.0, 20.0)
    for row in g["single"]:
        x, y, yaw, v = unhex(row["state"])
        s, ind, prev = laws.stanley_shadowed_steering(x, y, yaw, v, cx, cy, unhex(row["prev"]), unhex(row["dt"]),
                                                      unhex(row["k"]), unhex(row["k_soft"]), unhex(row["rate"]),
                                                      unhex(row["alpha"]))
        assert ind == row["ind"]
        assert same(s, unhex(row["steering"])) and same(prev, unhex(row["prev_out"])), row
    # the optional look-ahead advance (:276-281), on the closed ellipse and on open paths ending just ahead of the car
    assert sum(r["n_path"] < 720 for r in g["single_lookahead"]) >= 16
    for row in g["single_lookahead"]:
        x, y, yaw, v = unhex(row["state"])
        n = row["n_path"]
        s, ind, prev = laws.stanley_shadowed_steering(x, y, yaw, v, cx[:n], cy[:n], unhex(row["prev"]), unhex(row["dt"]),
                                                      unhex(row["k"]), unhex(row["k_soft"]),
                                                      lookahead_distance=unhex(row["lookahead"]))
        assert ind == row["ind"]
        assert same(s, unhex(row["steering"])) and same(prev, unhex(row["prev_out"])), row
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    L = g["loop"]
    want = np.asarray(unhex(L["traj"]))
    p = loop.default_params()
    p[loop.P_K_STANLEY], p[loop.P_K_SOFT] = laws.CONF.k_stanley, laws.CONF.k_soft_stanley
    spec = loop.RolloutSpec(steer="stanley_shadowed", accel="pid", model="dynamic", max_steps=len(want))
    out = loop.rollout_episode(loop.Track(cx, cy, curve, closed=True), spec, p, unhex(L["init"]))
    assert list(out["ti"]) == L["ti"]
    assert_same(out["traj"], want, "traj")
    assert_same(out["steer"], unhex(L["steer"]), "steer")
    assert_same(out["accel"], unhex(L["accel"]), "accel")


@pytest.mark.parametrize("name", ["ellipse_h40_s3", "ellipse_h12_s8_lf3", "open_quarter_h30_s2"])
def test_replanning_regime_matches_reference_nodes(name):
    """nodes/planner.py:17-102 + nodes/controller.py:28-94, run unmodified by oracle/gen_golden_replan.py around a
    window-planner stub and the reference's own State.update, vs the oracle's restated re-planning loop: the planner's
    index, the controller's index, steering, acceleration and the state, bit for bit, every tick."""
    import json
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "replan_kat.json")) as f:
        g = json.load(f)[name]
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    cx, cy = 40.0 * np.cos(th), 20.0 * np.sin(th)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    if not g["closed"]:
        cx, cy, curve = cx[:200], cy[:200], curve[:200]
    p = loop.default_params()
    p[loop.P_LF] = unhex(g["lf"])
    spec = loop.RolloutSpec(steer="pure_pursuit", accel="pid", model="dynamic", max_steps=g["steps"],
                            replan_horizon=g["horizon"], replan_stride=g["stride"], replan_after=3)
    out = loop.rollout_episode(loop.Track(cx, cy, curve, closed=g["closed"]), spec, p, unhex(g["init"]))
    n = out["steps"]                     # (the open path ends with PATH_END before the reference run was cut)
    assert n == g["steps"] or (not g["closed"] and out["status"] & loop.ST_PATH_END)
    assert n >= 100
    assert list(out["ti"]) == g["ti"][:n]
    assert_same(out["traj"], np.asarray(unhex(g["traj"]))[:n], "traj")
    assert_same(out["steer"], unhex(g["steer"])[:n], "steer")
    assert_same(out["accel"], unhex(g["accel"])[:n], "accel")
    # the regime itself: with the stock look-ahead the planner re-plans on every tick; with a coarse path and a short
    # controller look-ahead it holds a path for several ticks
    replans = np.diff([0] + g["replans"])
    assert replans.max() == 1 and (replans.min() == 1) == (name == "ellipse_h40_s3")


def test_mpc_qp_optimum_is_the_minimum_of_the_reference_cost():
    """oracle ``mpc_qp_optimum``: the exact optimum of the QP core/controllers.py:503-552 builds (two bounded
    least-squares chains).  Checked from first principles: it is feasible, no candidate of either bank and no
    feasible perturbation of it is cheaper under the cost the golden vectors pin (``mpc_candidate_cost``), and the
    v >= 0 constraint (:536) is honoured when it binds."""
    import bgr_pathplanning_control_b200 as bgr
    C = laws.CONF
    H, dt = 30, 0.05
    cx, cy = ellipse(720, 40.0, 20.0)
    rng = np.random.default_rng(7)
    banks = [bgr.candidate_bank(512, H).astype(np.float64),
             bgr.candidate_bank(1024, H, kind="qp_family").astype(np.float64)]
    for case in range(12):
        i = int(rng.integers(0, 720))
        x0, y0 = cx[i] + rng.normal(0, 0.3), cy[i] + rng.normal(0, 0.3)
        yaw0, v0 = float(rng.uniform(-3.2, 3.2)), float(rng.uniform(0.0, 9.0))
        ref = (i + np.arange(H)) % 720
        J, a, s_ = laws.mpc_qp_optimum(x0, y0, yaw0, v0, cx[ref], cy[ref], H, dt)
        assert np.all(np.abs(a) <= C.MAX_ACCEL + 1e-12) and np.all(np.abs(s_) <= C.MAX_STEER + 1e-12)
        assert np.all(v0 + dt * np.cumsum(a) >= -1e-12)
        for bank in banks:
            c, f = laws.mpc_costs_batch(x0, y0, yaw0, v0, cx[ref], cy[ref], bank, dt)
            assert np.all(c[f] >= J * (1 - 1e-9) - 1e-9)
        u = np.stack([a, s_], axis=1)
        for _ in range(200):                       # feasible perturbations never lower the cost (convexity + optimality)
            du = rng.normal(0, 0.05, u.shape)
            w = np.clip(u + du, [-C.MAX_ACCEL, -C.MAX_STEER], [C.MAX_ACCEL, C.MAX_STEER])
            cw, fw = laws.mpc_candidate_cost(x0, y0, yaw0, v0, cx[ref], cy[ref], w, dt)
            assert (not fw) or cw >= J * (1 - 1e-9) - 1e-9
    # the qp_family bank contains the optimum up to its grid: within 0.5 % here (0.2 % over config 4's states)
    J, _, _ = laws.mpc_qp_optimum(40.0, 0.2, 1.4, 4.0, cx[:H], cy[:H], H, dt)
    c, f = laws.mpc_costs_batch(40.0, 0.2, 1.4, 4.0, cx[:H], cy[:H], bgr.candidate_bank(4096, H, kind="qp_family"), dt)
    assert 0.0 <= c[f].min() / J - 1.0 < 5e-3
    # :536 binding: a target speed below zero pulls v through 0; the optimum stops AT 0 and beats every feasible candidate
    J, a, s_ = laws.mpc_qp_optimum(0.0, 0.0, 0.1, 0.3, cx[:H], cy[:H], H, dt, target_speed=-2.0)
    v = 0.3 + dt * np.cumsum(a)
    assert v.min() >= -1e-9 and v.min() < 1e-6
    c, f = laws.mpc_costs_batch(0.0, 0.0, 0.1, 0.3, cx[:H], cy[:H], banks[0], dt, target_speed=-2.0)
    assert f.any() and np.all(c[f] >= J * (1 - 1e-7))

