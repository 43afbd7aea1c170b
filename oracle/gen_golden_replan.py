"""Golden vectors for the RE-PLANNING regime of the live stack: nodes/planner.py:17-102 feeding nodes/controller.py:28-94.

TEST INFRASTRUCTURE.  Both node classes are imported from /root/reference and run UNMODIFIED, tick by tick, in the
order of main.py:46-48 (Planner, then Controller), around the reference's own ``State.update`` as the vehicle.  What is
stubbed is only what SURVEY.md puts out of scope:

* ``fsd_path_planning.PathPlanner`` -- replaced by the ORACLE-DEFINED window planner (oracle/loop.py:
  window_path_indices): the path planned from a car position is the window of a fixed centreline that starts at the
  waypoint nearest to it, returned in fsd-path-planning's ``(M, 4) = [s, x, y, curvature]`` layout with y NEGATED, so
  that the node's own ``cy = -path[:, 2]`` flip (nodes/planner.py:68,88) restores the frame;
* ``providers.sim.sim_util`` (FSDS RPC glue: ``sim_car_controls``, ``FSDSClientSingleton``) -- replaced by a recorder.

Frozen: per tick the planner's index, the controller's target index (recorded by wrapping the controller's own
``compute_steering``), steering, acceleration and the state; for an ellipse and for an OPEN path (window clipped at
its end).  Run in the build container:  python -m oracle.gen_golden_replan
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "replan_kat.json")


def hx(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [hx(e) for e in v]
    return float(v).hex()


class WindowPlanner:
    """Stand-in for fsd_path_planning.PathPlanner with the call signature nodes/planner.py:38-40 uses."""

    centreline = None          # (cx, cy, curve, closed), set by the driver below
    horizon = 40
    stride = 1
    calls = 0

    def __init__(self, mission):
        self.mission = mission

    def calculate_path_in_global_frame(self, cones, car_position, car_direction, return_intermediate_results=False):
        from oracle import loop
        cx, cy, curve, closed = WindowPlanner.centreline
        WindowPlanner.calls += 1
        j0 = int(np.argmin(np.hypot(cx - car_position[0], cy - car_position[1])))
        idx = loop.window_path_indices(len(cx), closed, j0, WindowPlanner.horizon, WindowPlanner.stride)
        s = np.arange(len(idx), dtype=np.float64)
        path = np.column_stack([s, cx[idx], -cy[idx], curve[idx]])
        empty = np.zeros((0, 2))
        return path, empty, empty, empty, empty, np.zeros(0, int), np.zeros(0, int)


_nodes = None


def load_nodes():
    """(reference namespace, nodes.planner, nodes.controller, list the stubbed simulator call appends to); cached --
    the node modules bind the stub functions at import time."""
    global _nodes
    if _nodes is None:
        _nodes = _load_nodes()
    return _nodes


def _load_nodes():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    fsd = sys.modules["fsd_path_planning"]
    fsd.PathPlanner = WindowPlanner
    fsd.MissionTypes = types.SimpleNamespace(trackdrive="trackdrive")
    sent = []
    sim = types.ModuleType("providers.sim.sim_util")
    sim.sim_car_controls = lambda client, steering, acceleration: sent.append((float(steering), float(acceleration)))
    sim.FSDSClientSingleton = type("FSDSClientSingleton", (), {"instance": classmethod(lambda cls: None)})
    for name in ("providers", "providers.sim"):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = []
            sys.modules[name] = pkg
    sys.modules["providers.sim.sim_util"] = sim
    import importlib
    planner_mod = importlib.import_module("nodes.planner")
    controller_mod = importlib.import_module("nodes.controller")
    return ref, planner_mod, controller_mod, sent


def run_case(ref, planner_mod, controller_mod, sent, cx, cy, curve, closed, horizon, stride, init, steps, lf=None):
    WindowPlanner.centreline = (cx, cy, curve, closed)
    WindowPlanner.horizon, WindowPlanner.stride, WindowPlanner.calls = horizon, stride, 0
    ref.FixedDtPID.dt = ref.conf.dt
    planner, controller = planner_mod.Planner(), controller_mod.Controller()
    if lf is not None:
        controller.steer_ctrl.look_ahead_distance = lf
    seen = []
    inner = controller.steer_ctrl.compute_steering

    def recording(state, path, target_ind):
        out = inner(state, path, target_ind)
        seen.append((int(target_ind), int(out[1])))
        return out

    controller.steer_ctrl.compute_steering = recording
    car = ref.car_state.State(x=init[0], y=init[1], yaw=init[2], v=init[3], r=init[4], beta=init[5])
    cones = [np.zeros((0, 2)) for _ in range(5)]
    data = {}
    rows = {k: [] for k in ("traj", "steer", "accel", "plan_ti", "ti", "replans")}
    del sent[:]
    for k in range(steps):
        data["car_state"] = car
        data["cones"] = (cones, np.array([car.x, car.y]), np.array([np.cos(car.yaw), np.sin(car.yaw)]))
        data["current_time"] = k * ref.conf.dt
        rows["traj"].append([car.x, car.y, car.yaw, car.v, car.r, car.beta])
        planner.update(data)
        controller.update(data)
        plan_ti, ti = seen[-1]
        assert plan_ti == data["target_ind"] == planner.target_ind
        assert sent[-1] == (-float(data["steering"]), float(data["acceleration"]))     # nodes/controller.py:90
        rows["steer"].append(float(data["steering"]))
        rows["accel"].append(float(data["acceleration"]))
        rows["plan_ti"].append(plan_ti)
        rows["ti"].append(ti)
        rows["replans"].append(WindowPlanner.calls)
        car.update(data["acceleration"], data["steering"])                            # the vehicle (car_state.py:118)
    return {"closed": bool(closed), "horizon": horizon, "stride": stride, "init": hx(init), "steps": steps,
            "lf": hx(controller.steer_ctrl.look_ahead_distance),
            "traj": hx(rows["traj"]), "steer": hx(rows["steer"]), "accel": hx(rows["accel"]),
            "plan_ti": rows["plan_ti"], "ti": rows["ti"], "replans": rows["replans"]}


def main():
    ref, planner_mod, controller_mod, sent = load_nodes()
    G = {"_meta": {"generator": "oracle/gen_golden_replan.py",
                   "note": "reference Planner + Controller nodes, unmodified, around a window-planner stub and the "
                           "reference's State.update; replan_after = 3 and the planner look-ahead are the nodes' own"}}
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    cx, cy = 40.0 * np.cos(th), 20.0 * np.sin(th)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    init = [40.2, -0.3, 1.5, 5.0, 0.0, 0.0]
    G["ellipse_h40_s3"] = run_case(ref, planner_mod, controller_mod, sent, cx, cy, curve, True, 40, 3, init, 400)
    # coarse path and a short controller look-ahead: the planner's index stays <= 3 for several ticks between re-plans
    G["ellipse_h12_s8_lf3"] = run_case(ref, planner_mod, controller_mod, sent, cx, cy, curve, True, 12, 8, init, 400,
                                       lf=3.0)
    # open path: a quarter of the ellipse; the windows are clipped at its last waypoint
    n_open = 200
    G["open_quarter_h30_s2"] = run_case(ref, planner_mod, controller_mod, sent, cx[:n_open], cy[:n_open],
                                        curve[:n_open], False, 30, 2, init, 260)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(G, f)
    for k, v in G.items():
        if k != "_meta":
            print(k, "replans", v["replans"][-1], "of", v["steps"], "ticks; plan_ti max", max(v["plan_ti"]),
                  "ti range", min(v["ti"]), max(v["ti"]))
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
