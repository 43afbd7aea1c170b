"""The callers either side of the hot path: the ``Controller`` node with the reference's
``update(data)`` dict protocol, the batched facade behind the same protocol, planner-path ingestion with the
``cy = -path[:, 2]`` flip, and the CSV wire format that ``metrics_calculator`` consumes."""
import numpy as np
import pytest

import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from bgr_pathplanning_control_b200 import metrics_calculator as mc
from conftest import Obj
from oracle import laws, loop

pytestmark = pytest.mark.gpu


def planner_path(ta):
    """What fsd-path-planning hands the Planner node: (M, 4) = [s, x, y, curvature] in the SIM frame (y flipped)."""
    ds = np.hypot(np.diff(ta.x, append=ta.x[0]), np.diff(ta.y, append=ta.y[0]))
    s = np.concatenate([[0.0], np.cumsum(ds)[:-1]])
    return np.column_stack([s, ta.x, -ta.y, ta.kappa])


def test_controller_node_update_protocol(cuda_device):
    """nodes/controller.py:28-94 driven for 40 control periods; laws vs the oracle restatement, dict keys, States."""
    ta = bgr.tracks.stadium_oval(256)
    path = planner_path(ta)
    cx, cy, curve = bgr.path_from_planner(path)
    assert np.array_equal(cy, ta.y) and np.array_equal(cx, ta.x)                  # nodes/planner.py:68,88
    sent = []
    node = bgr.Controller(actuator=lambda di, ai: sent.append((di, ai)))
    data = {"path": path, "cx": cx, "cy": cy, "curve": curve, "target_ind": 0}
    pid = laws.PIDState()
    state = [float(ta.x[0]), float(ta.y[0]) + 0.2, 0.05, 2.0, 0.0, 0.0]
    ti = 0
    for k in range(40):
        data["car_state"] = Obj(x=state[0], y=state[1], yaw=state[2], v=state[3], v_linear=1, v_angular=2,
                                a_angular=Obj(x_val=0.1, y_val=0.2), a_linear=Obj(x_val=0.3, y_val=0.4),
                                timestamp=0.05 * k)
        data["current_time"] = 0.05 * k
        data["target_ind"] = ti             # what a Planner node would have published (nodes/planner.py:101)
        node.update(data)
        assert data["target_ind"] == ti     # ... and the controller leaves it alone, like the reference's
        want_s, ti = laws.pure_pursuit_steering(state[0], state[1], state[2], cx, cy, ti, bgr.conf.LOOKAHEAD_DISTANCE)
        kap = curve[ti] if ti < len(curve) else curve[-1]
        want_a, want_v = laws.pid_acceleration(pid, state[3], kap, bgr.conf.kp_accel, bgr.conf.ki_accel,
                                               bgr.conf.kd_accel, bgr.conf.dt)
        assert node.target_ind == ti
        assert abs(data["steering"] - want_s) < 1e-12 and abs(data["acceleration"] - want_a) < 1e-11
        assert abs(data["v_log"] - want_v) < 1e-12
        assert sent[-1] == (-data["steering"], data["acceleration"])              # sign flip on the way out (:90)
        st, th, br = data["controls"]
        assert abs(st - (-want_s / bgr.conf.MAX_STEER)) < 1e-12 and 0.0 <= th <= 1.0 and br == 0.0
        state = laws.kinematic_bicycle_step(state, want_s, want_a, bgr.conf.dt)
    hist = data["states"]
    assert hist is node.states and len(hist.x) == 40 and len(hist.steering) == 40 and len(hist.v_log) == 40
    assert hist.a_longitudinal[0] == 0.1 and hist.a_lateral[0] == 0.2     # a_linear=car_state.a_angular (:74, sic)
    assert hist.y[0] == -(float(ta.y[0]) + 0.2)                                   # the logged state is y-flipped (:43-45)
    # missing inputs: warn and return, nothing appended (:37-41)
    node.update({"cx": cx})
    assert len(node.states.x) == 40
    assert bgr.sim_controls(0.1, -1.0)[1:] == (0.0, 0.0)    # brake = clip(-a / MAX_DECEL) with MAX_DECEL < 0 is 0


def test_batched_controller_and_csv_wire_format(cuda_device, tmp_path):
    ta = bgr.tracks.stadium_oval(512)
    cx, cy, curve = bgr.path_from_planner(planner_path(ta))
    E = 12
    node = bgr.BatchedController(steer="pure_pursuit", accel="pid", model="dynamic", max_steps=400, laps=2)
    params = bgr.default_params(E)
    params[:, abi.P_LF] = np.linspace(4.0, 7.0, E)
    data = {"cx": cx, "cy": cy, "curve": curve, "car_state": Obj(x=ta.x[0], y=ta.y[0], yaw=0.0, v=5.0)}
    res = node.run(data, params, closed=True)
    assert data["rollout"] is res and res.metrics.shape == (E, abi.N_METRICS)
    assert np.all(res.status[:, abi.S_STEPS] == 400)
    # identical to calling the engine directly
    track = bgr.Track(cx, cy, curve, closed=True)
    init = np.tile([[ta.x[0], ta.y[0], 0.0, 5.0, 0.0, 0.0]], (E, 1))
    direct = bgr.rollout(track, node.config, params, init, traj=True, aggregate=True)
    assert np.array_equal(direct.metrics, res.metrics) and np.array_equal(direct.status, res.status)
    # States history of one episode (what core/visualization.py plots)
    hist = node.states(res, 3)
    assert len(hist.x) == 400 and hist.x[0] == ta.x[0] and len(hist.steering) == 400
    assert np.allclose(hist.t, np.arange(400) * bgr.conf.dt)
    # CSV in the reference schema -> pandas -> metrics_calculator == what the kernel accumulated
    fn = tmp_path / "simulation_log.csv"
    logger = bgr.SimulationLogger(str(fn))
    frame = logger.log_episode(res, 3)
    logger.log(400 * 0.05, 0, 0, 0, 0, 0, 0, 0.0, 0.0, 0, 0)     # the reference's own row-wise API still works
    loaded = mc.load_simulation_data(str(fn))
    assert list(loaded.columns) == mc.CSV_COLUMNS and len(loaded) == 401
    loaded = loaded.iloc[:400]
    lat_mean, lat_std = mc.calculate_lateral_error(loaded)
    head_mean, head_std = mc.calculate_heading_error(loaded)
    lap = mc.calculate_lap_time(loaded)
    m = res.metrics[3]
    assert abs(lat_mean - m[abi.M_LAT_MEAN]) < 1e-6 and abs(lat_std - m[abi.M_LAT_STD]) < 1e-6
    assert abs(head_mean - m[abi.M_HEAD_MEAN]) < 1e-6 and abs(head_std - m[abi.M_HEAD_STD]) < 1e-6
    assert abs(lap - m[abi.M_LAP_TIME]) < 1e-6
    # and equal to pandas' own mean / std (ddof = 1) on the same frame: the reference's three functions verbatim
    assert abs(lat_std - frame["lateral_error"].std()) < 1e-12 and abs(lat_mean - frame["lateral_error"].mean()) < 1e-12
    rm, rs = mc.calculate_lateral_error(res)
    assert rm.shape == (E,) and rs.shape == (E,)
    # throttle / brake columns follow providers/sim/sim_util.py:197-209 (brake is always 0: MAX_DECEL < 0)
    assert np.all(frame["brake"] == 0.0) and frame["throttle"].between(0.0, 1.0).all()
