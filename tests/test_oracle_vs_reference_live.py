"""The oracle against the reference ITSELF, imported live from /root/reference (build container only; skipped where
the tree is absent, e.g. on the GPU box -- nothing here is a ``gpu`` test).

The frozen vectors of tests/golden/ pin the oracle on a fixed set of inputs; this file re-draws inputs at random
(hypothesis-style seeds) and calls the reference's own unmodified functions next to the restatement, so a
restatement that only happened to match the frozen cases would be caught: laws one call at a time, the vehicle model,
the metrics, closed loops driven by the reference's own objects, and the re-planning regime driven by the reference's
own Planner and Controller nodes.  Bit-exact (same interpreter, same libm).
"""
import math

import numpy as np
import pytest

from oracle import laws, loop, ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


def random_path(rng):
    n = int(rng.integers(30, 500))
    th = np.linspace(0, 2 * np.pi, n, endpoint=False)
    r = rng.uniform(15, 40) + sum(rng.uniform(-1.5, 1.5) * np.cos(k * th + rng.uniform(0, 6.28)) for k in range(2, 5))
    return r * np.cos(th), rng.uniform(0.5, 1.0) * r * np.sin(th)


def make_state(ref, x, y, yaw, v, r=0.0, beta=0.0):
    return ref.car_state.State(x=x, y=y, yaw=yaw, v=v, r=r, beta=beta)


@pytest.mark.parametrize("seed", range(12))
def test_steering_laws_single_calls(ref, seed):
    rng = np.random.default_rng(seed)
    cx, cy = random_path(rng)
    path = np.column_stack([np.arange(len(cx)), cx, cy])
    for _ in range(40):
        j = int(rng.integers(0, len(cx)))
        x, y = cx[j] + rng.normal(0, 1.5), cy[j] + rng.normal(0, 1.5)
        yaw, v = rng.uniform(-7, 7), rng.uniform(0, 12)
        ti = int(rng.integers(0, len(cx)))
        lf = rng.uniform(1.0, 9.0)
        st = make_state(ref, x, y, yaw, v)
        want_s, want_i = ref.controllers.PurePursuitController(look_ahead_distance=lf).compute_steering(st, path, ti)
        got_s, got_i = laws.pure_pursuit_steering(x, y, yaw, cx, cy, ti, lf)
        assert got_i == want_i and float(got_s) == float(want_s)
        k, ks = rng.uniform(0.2, 2.0), rng.uniform(0.01, 1.0)
        want_s, want_i = ref.controllers.StanleyController(k=k, k_soft=ks).compute_steering(st, path, ti)
        got_s, got_i = laws.stanley_steering(x, y, yaw, v, cx, cy, k, ks)
        assert got_i == int(want_i) and float(got_s) == float(want_s)
        assert laws.find_target_point(x, y, cx, cy, ti) == ref.controllers.find_target_point(st, cx, cy, ti)
        a = rng.uniform(-50, 50)
        assert laws.normalize_angle(a) == ref.controllers.normalize_angle(a)


@pytest.mark.parametrize("seed", range(6))
def test_vehicle_model_and_metrics(ref, seed):
    rng = np.random.default_rng(100 + seed)
    model = ref.car_state.DynamicBicycleModel()
    conf = ref.conf
    for _ in range(60):
        s6 = [rng.normal(0, 30), rng.normal(0, 30), rng.uniform(-10, 10), rng.uniform(-1, 14), rng.normal(0, 0.5),
              rng.normal(0, 0.1)]
        delta, a = rng.uniform(-0.6, 0.6), rng.uniform(-2.5, 2.0)
        want = model.state_equations(list(s6), [delta, a], conf.dt)
        got = laws.dynamic_bicycle_step(s6, delta, a, conf.dt)
        assert [float(u) for u in got] == [float(u) for u in want]
    import pandas as pd
    n = int(rng.integers(2, 400))
    lat, hed = rng.normal(0.05, 0.2, n), rng.normal(0, 0.05, n)
    t = np.arange(n) * conf.dt
    df = pd.DataFrame({"timestamp": t, "lateral_error": lat, "heading_error": hed})
    mc = ref.metrics_calculator
    got = laws.lap_metrics(t, lat, hed)
    assert (got["lateral_mean"], got["lateral_std"]) == tuple(float(u) for u in mc.calculate_lateral_error(df))
    assert (got["heading_mean"], got["heading_std"]) == tuple(float(u) for u in mc.calculate_heading_error(df))
    assert got["lap_time"] == float(mc.calculate_lap_time(df))


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("steer,accel", [("pure_pursuit", "pid"), ("stanley", "lqg")])
def test_closed_loop_with_reference_objects(ref, seed, steer, accel):
    """The reference's own controller objects and State.update stepped in the order of nodes/controller.py:61-63."""
    rng = np.random.default_rng(200 + seed)
    cx, cy = random_path(rng)
    curve = rng.uniform(0.0, 0.08, len(cx))
    conf = ref.conf
    ref.FixedDtPID.dt = conf.dt
    C = ref.controllers
    lf = float(rng.uniform(3.0, 7.0))
    sc = C.PurePursuitController(look_ahead_distance=lf) if steer == "pure_pursuit" else C.StanleyController()
    ac = (C.AccelerationPIDController(kp=conf.kp_accel, ki=conf.ki_accel, kd=conf.kd_accel, setpoint=conf.TARGET_SPEED)
          if accel == "pid" else C.LQGAccelerationController(dt=conf.dt))
    yaw0 = math.atan2(cy[1] - cy[0], cx[1] - cx[0])
    init = [cx[0] + rng.normal(0, 0.2), cy[0] + rng.normal(0, 0.2), yaw0, 5.0, 0.0, 0.0]
    car = make_state(ref, *init)
    path = np.column_stack([np.arange(len(cx)), cx, cy])
    S, ti = 250, 0
    want = []
    for _ in range(S):
        want.append([car.x, car.y, car.yaw, car.v, car.r, car.beta])
        steering, ti = sc.compute_steering(car, path, ti)
        kappa = curve[ti] if ti < len(curve) else curve[-1]
        acceleration = ac.compute_acceleration(car.v, kappa)
        car.update(acceleration, steering)
        if steer == "pure_pursuit" and ti >= len(cx) - 1:
            break
    p = loop.default_params()
    p[loop.P_LF] = lf
    spec = loop.RolloutSpec(steer=steer, accel=accel, model="dynamic", max_steps=len(want))
    out = loop.rollout_episode(loop.Track(cx, cy, curve, closed=True), spec, p, init)
    n = min(out["steps"], len(want))
    assert n >= min(50, len(want))
    got, ref_traj = np.asarray(out["traj"])[:n], np.asarray(want)[:n]
    assert np.array_equal(got, ref_traj), np.abs(got - ref_traj).max()


@pytest.mark.parametrize("seed", range(4))
def test_replanning_regime_with_reference_nodes(seed):
    """Planner.update + Controller.update (unmodified) around the window-planner stub, random path / horizon / stride."""
    from oracle import gen_golden_replan as gg
    ref, planner_mod, controller_mod, sent = gg.load_nodes()
    rng = np.random.default_rng(300 + seed)
    cx, cy = random_path(rng)
    curve = rng.uniform(0.0, 0.08, len(cx))
    horizon, stride = int(rng.integers(3, 50)), int(rng.integers(1, 7))
    lf = float(rng.uniform(2.0, 8.0))
    yaw0 = math.atan2(cy[1] - cy[0], cx[1] - cx[0])
    init = [cx[0] + rng.normal(0, 0.2), cy[0] + rng.normal(0, 0.2), yaw0, 5.0, 0.0, 0.0]
    closed = bool(seed % 2 == 0)
    g = gg.run_case(ref, planner_mod, controller_mod, sent, cx, cy, curve, closed, horizon, stride, init, 200, lf=lf)
    p = loop.default_params()
    p[loop.P_LF] = lf
    spec = loop.RolloutSpec(steer="pure_pursuit", accel="pid", model="dynamic", max_steps=200, replan_horizon=horizon,
                            replan_stride=stride, replan_after=3)
    out = loop.rollout_episode(loop.Track(cx, cy, curve, closed=closed), spec, p, init)
    n = out["steps"]
    assert n == 200 or (not closed and out["status"] & loop.ST_PATH_END)
    want = np.array([[float.fromhex(v) for v in row] for row in g["traj"]])[:n]
    assert list(out["ti"]) == g["ti"][:n]
    assert np.array_equal(np.asarray(out["traj"]), want)
