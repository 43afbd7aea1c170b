"""The drop-in classes (bgr_pathplanning_control_b200.controllers / car_state: the reference's names and
call conventions, evaluated by the fp64 audit kernels through the C ABI ``bgr_step_host``) against the golden
vectors frozen from the reference's own unmodified code (tests/golden/reference_kat.json).

These tests read like the reference's own tests would: build a controller, call ``compute_steering`` /
``compute_acceleration`` / ``State.update``, compare with what the reference returned for the same inputs.
Bar: indices exact; floats within a few ulps of fp64 (the device libm's sin / cos / atan2 / exp / hypot are
1-2 ulp functions and differ from glibc's in the last place) -- TOL below.
"""
import math

import numpy as np
import pytest

import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from conftest import GOLDEN_PATHS, Obj, ellipse_path, unhex

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-13, 2e-15      # fp64 kernels vs the reference's fp64 numpy: last-place libm differences only


def close(a, b, rtol=RTOL, atol=ATOL):
    return abs(float(a) - float(b)) <= atol + rtol * abs(float(b))


def path3(cx, cy):
    return np.column_stack([np.arange(len(cx)), cx, cy])      # nodes/controller.py:56


@pytest.mark.parametrize("pname", list(GOLDEN_PATHS))
def test_pure_pursuit_compute_steering(golden, cuda_device, pname):
    cx, cy = ellipse_path(*GOLDEN_PATHS[pname])
    path = path3(cx, cy)
    for row in golden[f"pure_pursuit/{pname}"]:
        x, y, yaw, v = unhex(row["state"])
        ctrl = bgr.PurePursuitController(look_ahead_distance=unhex(row["Lf"]))
        steering, ind = ctrl.compute_steering(Obj(x=x, y=y, yaw=yaw, v=v), path, row["start"])
        assert ind == row["ind"] and isinstance(ind, int)
        assert isinstance(steering, np.float64)
        assert close(steering, unhex(row["steering"])), (row, steering)


def test_pure_pursuit_clamps_to_the_last_index(golden, cuda_device):
    g = golden["pure_pursuit/clamp"]
    cx, cy = ellipse_path(400, 30.0, 15.0)
    x, y, yaw, v = unhex(g["state"])
    steering, ind = bgr.PurePursuitController().compute_steering(
        Obj(x=x, y=y, yaw=yaw, v=v), path3(cx[:g["n"]], cy[:g["n"]]), g["start"])
    assert ind == g["ind"] == g["n"] - 1
    assert close(steering, unhex(g["steering"]))


@pytest.mark.parametrize("pname", list(GOLDEN_PATHS))
def test_stanley_compute_steering(golden, cuda_device, pname):
    cx, cy = ellipse_path(*GOLDEN_PATHS[pname])
    path = path3(cx, cy)
    for row in golden[f"stanley/{pname}"]:
        x, y, yaw, v = unhex(row["state"])
        ctrl = bgr.StanleyController(k=unhex(row["k"]), k_soft=unhex(row["k_soft"]))
        steering, ind = ctrl.compute_steering(Obj(x=x, y=y, yaw=yaw, v=v), path, 0)
        assert int(ind) == row["ind"]
        assert close(steering, unhex(row["steering"]), atol=1e-14), (row, steering)


@pytest.mark.parametrize("pname", list(GOLDEN_PATHS))
def test_find_target_point(golden, cuda_device, pname):
    cx, cy = ellipse_path(*GOLDEN_PATHS[pname])
    for row in golden[f"find_target_point/{pname}"]:
        x, y, yaw, v = unhex(row["state"])
        assert bgr.find_target_point(Obj(x=x, y=y, yaw=yaw, v=v), cx, cy, row["start"]) == row["ind"]


def test_survey_starter_vectors(cuda_device):
    """SURVEY.md section 8c (produced from the unmodified reference)."""
    cx, cy = ellipse_path(720, 40.0, 20.0)
    st = Obj(x=40.3, y=-0.2, yaw=1.5, v=5.0)
    s, ind = bgr.PurePursuitController().compute_steering(st, path3(cx, cy), 0)
    assert ind == 30 and close(s, 0.2328625291477659)
    s, ind = bgr.get_steering_controller("stanley").compute_steering(st, path3(cx, cy), 0)
    assert ind == 719 and close(s, 0.10378436919576073, atol=1e-14)
    cx, cy = ellipse_path(400, 30.0, 15.0)
    st = Obj(x=30.0, y=0.0, yaw=math.pi / 2, v=5.0)
    s, ind = bgr.get_steering_controller("pure_pursuit").compute_steering(st, path3(cx, cy), 0)
    assert ind == 23 and close(s, 0.22090553298549134)
    s, ind = bgr.StanleyController().compute_steering(st, path3(cx, cy), 0)
    assert ind == 0 and close(s, 0.01570699444132373, atol=1e-14)
    assert bgr.find_target_point(st, cx, cy, 0) == 23
    assert bgr.normalize_angle(3.5) == -2.7831853071795862
    assert bgr.normalize_angle(math.pi) == -math.pi and bgr.normalize_angle(-math.pi) == -math.pi


def test_compute_breaking(golden, cuda_device):
    g = golden["compute_breaking"]
    ctrl = bgr.AccelerationPIDController(0.35, 0.4, 0.3, bgr.conf.TARGET_SPEED)
    for k, v in zip(unhex(g["kappa"]), unhex(g["v"])):
        assert close(ctrl.compute_breaking(k), v)
    assert close(ctrl.compute_breaking(0.08), 6.9167221482221635)


def test_pid_compute_acceleration_sequence(golden, cuda_device):
    """One controller instance = one episode: the integral and last input live in the instance."""
    g = golden["pid_sequence"]
    kp, ki, kd = unhex(g["gains"])
    assert unhex(g["dt"]) == bgr.conf.dt
    ctrl = bgr.AccelerationPIDController(kp, ki, kd, bgr.conf.TARGET_SPEED)
    for v, k, a, ms in zip(unhex(g["v"]), unhex(g["kappa"]), unhex(g["accel"]), unhex(g["maxspeed"])):
        acc = ctrl.compute_acceleration(v, k)
        assert close(acc, a, rtol=1e-12, atol=1e-13), (acc, a)
        assert close(ctrl.maxspeed, ms)                      # read by nodes/controller.py:71,84


def test_lqg_compute_acceleration_sequence(golden, cuda_device):
    g = golden["lqg_sequence"]
    ctrl = bgr.LQGAccelerationController(unhex(g["dt"]))
    np.testing.assert_allclose(ctrl.K.ravel(), unhex(golden["lqg_K"]), rtol=1e-12)
    for i, (v, k) in enumerate(zip(unhex(g["v"]), unhex(g["kappa"]))):
        acc = ctrl.compute_acceleration(v, k)
        assert close(acc, unhex(g["accel"][i]), rtol=1e-10, atol=1e-11), (i, acc)
        assert close(ctrl.v_desired, unhex(g["v_desired"][i]))
        np.testing.assert_allclose(ctrl.x_hat.ravel(), np.ravel(unhex(g["x_hat"][i])), rtol=1e-10, atol=1e-12)


def test_state_update_is_the_dynamic_bicycle(golden, cuda_device):
    """State.update(a, delta): argument order (a, delta), model input [delta, a] (car_state.py:118-125)."""
    g = golden["state_update"]
    x, y, yaw, v, r, beta = unhex(g["init"])
    st = bgr.State(x=x, y=y, yaw=yaw, v=v, r=r, beta=beta)
    st.update(unhex(g["a"]), unhex(g["delta"]))
    got = [st.x, st.y, st.yaw, st.v, st.r, st.beta]
    for a, b in zip(got, unhex(g["out"])):
        assert close(a, b), (got, unhex(g["out"]))
    for a, b in zip([st.rear_x, st.rear_y, st.front_x, st.front_y], unhex(g["axles"])):
        assert close(a, b)
    # State(0,0,0,5).update(0.5, 0.1)
    st = bgr.State(0.0, 0.0, 0.0, 5.0)
    st.update(0.5, 0.1)
    assert close(st.x, 0.25) and close(st.v, 5.025) and close(st.r, 0.04235294117647059) and st.beta == 0.0


def test_state_equations(golden, cuda_device):
    """DynamicBicycleModel.state_equations rows (default dt only: State.update always uses conf.dt)."""
    lib = abi.load()
    import ctypes as C
    from bgr_pathplanning_control_b200 import controllers, engine

    rows = [(r, {}) for r in golden["state_equations"]] + \
           [(r, {"mu": unhex(r["mu"]), "c_f": unhex(r["c_f"]), "c_r": unhex(r["c_r"])})
            for r in golden["state_equations_lowmu"]]
    n = len(rows)
    params = engine.default_params(n, np.float64)
    state = np.zeros((n, abi.N_STATE))
    cmd = np.zeros((n, abi.N_CMD))
    want = np.zeros((n, 6))
    dts = set()
    for i, (r, over) in enumerate(rows):
        state[i] = unhex(r["x"])
        cmd[i, abi.CMD_STEER], cmd[i, abi.CMD_ACCEL] = unhex(r["u"])      # u = [delta, a]
        want[i] = unhex(r["out"])
        dts.add(unhex(r["dt"]))
        if over:
            params[i, abi.P_MU], params[i, abi.P_C_F], params[i, abi.P_C_R] = over["mu"], over["c_f"], over["c_r"]
    for dt in sorted(dts):
        sel = np.array([unhex(r["dt"]) == dt for r, _ in rows])
        st = np.ascontiguousarray(state[sel])
        pr = np.ascontiguousarray(params[sel])
        cm = np.ascontiguousarray(cmd[sel])
        cfg = engine.RolloutConfig(model="dynamic", precision="fp64", dt=dt).to_struct()
        abi.check(lib.bgr_step_host(controllers._dummy_track().handle, C.byref(cfg), abi.PHASE_MODEL, int(sel.sum()),
                                    pr.ctypes.data, st.ctypes.data, None, None, cm.ctypes.data, None, None))
        np.testing.assert_allclose(st, want[sel], rtol=1e-13, atol=1e-15)
    # SURVEY 8c starter vectors
    st = np.array([[1, 2, 0.3, 6, 0.2, 0.05], [0, 0, 0, 0, 0, 0]], dtype=np.float64)
    cm = np.zeros((2, abi.N_CMD))
    cm[0, :2] = 0.15, 1.0
    cm[1, :2] = 0.3, 1.5
    cfg = engine.RolloutConfig(model="dynamic", precision="fp64", dt=0.05).to_struct()
    abi.check(lib.bgr_step_host(controllers._dummy_track().handle, C.byref(cfg), abi.PHASE_MODEL, 2,
                                engine.default_params(2, np.float64).ctypes.data, st.ctypes.data, None, None,
                                cm.ctypes.data, None, None))
    np.testing.assert_allclose(st[0], [1.2818118138542136, 2.1028693422366356, 0.31, 6.05, 0.23864705882352943,
                                       0.02888888888888889], rtol=1e-14)
    np.testing.assert_allclose(st[1], [5.000000000000001e-07, 0.0, 0.0, 0.07501000000000001, 0.06621750000000003, 0.0],
                               rtol=1e-14, atol=0)


LOOPS = {"loop/pp_pid_dynamic": ("pure_pursuit", "pid"), "loop/stanley_lqg_dynamic": ("stanley", "lqg"),
         "loop/stanley_pid_dynamic": ("stanley", "pid")}


def golden_loop_track():
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    cx, cy = 40.0 * np.cos(th), 20.0 * np.sin(th)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    return cx, cy, curve


@pytest.mark.parametrize("name", list(LOOPS))
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_closed_loop_against_the_reference_objects(golden, cuda_device, name, precision):
    """400 closed-loop steps driven by the reference's OWN controller and State objects (frozen by
    oracle/gen_golden.py in the order of nodes/controller.py:61-63) vs one ``bgr_rollout`` call.
    fp64 audit kernels: ~1e-10; fp32 production kernels: the north-star tolerances (1e-3 m, 1e-4 rad)."""
    g = golden[name]
    steer, accel = LOOPS[name]
    cx, cy, curve = golden_loop_track()
    track = bgr.Track(cx, cy, curve, closed=True)
    want = np.asarray(unhex(g["traj"]))
    n = len(want)
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model="dynamic", precision=precision, max_steps=n, laps=1)
    params = bgr.default_params(1)
    init = np.asarray(unhex(g["init"]), dtype=np.float64).reshape(1, 6)
    res = bgr.rollout(track, cfg, params, init, traj=True)
    assert res.status[0, abi.S_STEPS] == n
    tr = res.traj[0]
    assert list(tr[:, abi.T_TARGET_IND].astype(int)) == g["ti"]
    # the batched parameter rows are float32 (kp = 0.35f ... differ from the reference's doubles by 1e-8
    # relative): that, not the kernel arithmetic, sets the fp64 tolerance
    pos_tol, yaw_tol, u_tol = (5e-6, 5e-7, 5e-6) if precision == "fp64" else (1e-3, 1e-4, 2e-4)
    assert np.hypot(tr[:, abi.T_X] - want[:, 0], tr[:, abi.T_Y] - want[:, 1]).max() < pos_tol
    assert np.abs(tr[:, abi.T_YAW] - want[:, 2]).max() < yaw_tol
    assert np.abs(tr[:, abi.T_V] - want[:, 3]).max() < pos_tol * 10
    assert np.abs(tr[:, abi.T_STEER] - np.asarray(unhex(g["steer"]))).max() < u_tol
    assert np.abs(tr[:, abi.T_ACCEL] - np.asarray(unhex(g["accel"]))).max() < u_tol * 10
    final = np.asarray(unhex(g["final"]))
    np.testing.assert_allclose(res.metrics[0, abi.M_X:abi.M_BETA + 1], final, rtol=0, atol=pos_tol * 10)


def test_metrics_calculator_columns(golden, cuda_device):
    """calculate_lateral_error / calculate_heading_error / calculate_lap_time on logged columns
    (preformance_analysis/metrics_calculator.py:1-8; pandas std => ddof = 1)."""
    from bgr_pathplanning_control_b200 import metrics_calculator as mc
    for row in golden["metrics"]:
        data = {"timestamp": np.asarray(unhex(row["timestamp"])), "lateral_error": np.asarray(unhex(row["lateral_error"])),
                "heading_error": np.asarray(unhex(row["heading_error"]))}
        m, s = mc.calculate_lateral_error(data)
        assert close(m, unhex(row["lateral_mean"]), rtol=1e-12) and close(s, unhex(row["lateral_std"]), rtol=1e-12)
        m, s = mc.calculate_heading_error(data)
        assert close(m, unhex(row["heading_mean"]), rtol=1e-12, atol=1e-15)
        assert close(s, unhex(row["heading_std"]), rtol=1e-12)
        assert mc.calculate_lap_time(data) == unhex(row["lap_time"])
    m, s = mc.calculate_lateral_error({"lateral_error": [.1, .2, .4]})
    assert close(m, 0.23333333333333336, rtol=1e-14) and close(s, 0.1527525231651947, rtol=1e-14)


def test_mpc_compute_control_cost_definition(golden, cuda_device):
    """The cost of a given control sequence as the reference's compute_control evaluates it (through the eager
    cvxpy stand-in of oracle/gen_golden.py) vs the fp64 MPC kernel scoring that sequence as a 1-candidate bank;
    and the return convention (acceleration, -steering) of :565-568."""
    for row in golden["mpc_cost"]:
        x0, y0, yaw0, v0 = unhex(row["state"])
        u = np.asarray(unhex(row["u"]))
        H = row["N"]
        rx, ry = np.asarray(unhex(row["ref_x"])), np.asarray(unhex(row["ref_y"]))
        track = bgr.Track(rx, ry, None, closed=False)
        cfg = bgr.MpcConfig(horizon=H, dt=unhex(row["dt"]), precision="fp64")
        bank = u.astype(np.float32).reshape(1, H, 2)
        res = bgr.mpc_evaluate(track, cfg, np.array([[x0, y0, yaw0, v0]]), np.zeros(1, np.int32), bank,
                               return_costs=True)
        feasible = unhex(row["max_ineq_violation"]) <= 0.0
        # the bank is float32 (the ABI's candidate format): re-evaluate the golden's cost at the rounded controls
        from oracle import laws
        want, feas32 = laws.mpc_candidate_cost(x0, y0, yaw0, v0, rx, ry, bank[0].astype(np.float64), unhex(row["dt"]))
        assert abs(want - unhex(row["cost"])) <= 1e-5 * max(1.0, abs(want))       # fp32 rounding of u only
        if feas32:
            assert np.isclose(res.costs[0, 0], want, rtol=1e-6)
            assert res.best_idx[0] == 0
            if feasible:
                assert np.isclose(res.best_u[0, 0], unhex(row["ret_accel"]), atol=1e-6)
                assert np.isclose(res.best_u[0, 1], unhex(row["ret_steer"]), atol=1e-6)
        else:
            assert res.best_idx[0] == -1 and np.isinf(res.costs[0, 0])
            assert res.best_u[0, 0] == 0.0 and res.best_u[0, 1] == 0.0             # :558-562 fallback


def test_shadowed_stanley_variant(cuda_device):
    """SURVEY 8 row a2': the dead-code Stanley body (core/controllers.py:234-324), golden vectors produced by executing
    the reference's own source lines (oracle/gen_golden_shadowed.py).  Single calls through the labelled drop-in class
    (fp64 audit kernel), then the 400-step closed loop through one bgr_rollout in fp64 and in fp32."""
    import json
    import os
    from conftest import GOLDEN_DIR
    with open(os.path.join(GOLDEN_DIR, "shadowed_stanley_kat.json")) as f:
        g = json.load(f)
    cx, cy = ellipse_path(720, 40.0, 20.0)
    path = path3(cx, cy)
    for row in g["single"]:
        x, y, yaw, v = unhex(row["state"])
        ctl = bgr.ShadowedStanleyController(k=unhex(row["k"]), k_soft=unhex(row["k_soft"]),
                                            max_steer_rate=unhex(row["rate"]), alpha=unhex(row["alpha"]))
        ctl.previous_steering = unhex(row["prev"])
        s, ind = ctl.compute_steering(Obj(x=x, y=y, yaw=yaw, v=v), path, 0, unhex(row["dt"]))
        assert int(ind) == row["ind"]
        assert close(s, unhex(row["steering"]), atol=1e-13) and close(ctl.previous_steering, unhex(row["prev_out"]), atol=1e-13)
    # the optional look-ahead advance (:276-281): closed ellipse and open paths that end just ahead of the car
    advanced = 0
    for row in g["single_lookahead"]:
        x, y, yaw, v = unhex(row["state"])
        n = row["n_path"]
        ctl = bgr.ShadowedStanleyController(k=unhex(row["k"]), k_soft=unhex(row["k_soft"]),
                                            lookahead_distance=unhex(row["lookahead"]))
        ctl.previous_steering = unhex(row["prev"])
        s, ind = ctl.compute_steering(Obj(x=x, y=y, yaw=yaw, v=v), path[:n], 0, unhex(row["dt"]))
        assert int(ind) == row["ind"], row
        assert close(s, unhex(row["steering"]), atol=1e-13) and close(ctl.previous_steering, unhex(row["prev_out"]), atol=1e-13)
        plain = bgr.ShadowedStanleyController(k=unhex(row["k"]), k_soft=unhex(row["k_soft"]))
        advanced += int(plain.compute_steering(Obj(x=x, y=y, yaw=yaw, v=v), path[:n], 0, unhex(row["dt"]))[1] != row["ind"])
    assert advanced >= 48          # the vectors do exercise the advance
    # ... and the same advance inside a rollout: fp64 and fp32 builds against the oracle loop
    from oracle import loop
    ta = bgr.tracks.stadium_oval(256)
    trk = bgr.Track.from_arrays(ta)
    p1 = bgr.default_params(1)
    i1 = bgr.initial_states(ta, 1, lateral_offset=[0.2], v0=4.0)
    ref = loop.rollout_episode(loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones),
                               loop.RolloutSpec(steer="stanley_shadowed", accel="pid", model="kinematic", max_steps=300,
                                                shadowed_lookahead=2.5), p1[0].astype(np.float64), i1[0])
    for precision, pos_tol in (("fp64", 1e-9), ("fp32", 1e-3)):
        cfg = bgr.RolloutConfig(steer="stanley", accel="pid", model="kinematic", precision=precision, max_steps=300,
                                stanley_shadowed=True, stanley_lookahead=2.5)
        tr = bgr.rollout(trk, cfg, p1, i1, traj=True).traj[0]
        assert np.array_equal(tr[:, abi.T_TARGET_IND].astype(int), ref["ti"]), precision
        assert np.array_equal(tr[:, abi.T_NEAREST_IND].astype(int), ref["near"]), precision
        assert np.hypot(tr[:, abi.T_X] - ref["traj"][:, 0], tr[:, abi.T_Y] - ref["traj"][:, 1]).max() < pos_tol
        assert np.abs(tr[:, abi.T_E_CT] - ref["e_ct"]).max() < max(pos_tol, 1e-9)
    _, _, curve = golden_loop_track()
    L = g["loop"]
    want = np.asarray(unhex(L["traj"]))
    track = bgr.Track(cx, cy, curve, closed=True)
    params = bgr.default_params(1)
    params[0, abi.P_K_STANLEY], params[0, abi.P_K_SOFT] = bgr.conf.k_stanley, bgr.conf.k_soft_stanley
    init = np.asarray(unhex(L["init"]), dtype=np.float64).reshape(1, 6)
    for precision, pos_tol, yaw_tol in (("fp64", 5e-6, 5e-7), ("fp32", 1e-3, 1e-4)):
        cfg = bgr.RolloutConfig(steer="stanley", accel="pid", model="dynamic", precision=precision, max_steps=len(want),
                                stanley_shadowed=True)
        res = bgr.rollout(track, cfg, params, init, traj=True)
        tr = res.traj[0]
        assert list(tr[:, abi.T_TARGET_IND].astype(int)) == L["ti"]
        assert np.hypot(tr[:, abi.T_X] - want[:, 0], tr[:, abi.T_Y] - want[:, 1]).max() < pos_tol
        assert np.abs(tr[:, abi.T_YAW] - want[:, 2]).max() < yaw_tol
        assert np.abs(tr[:, abi.T_STEER] - np.asarray(unhex(L["steer"]))).max() < 20 * yaw_tol
