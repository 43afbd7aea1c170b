"""Golden vectors for the SHADOWED (dead-code) Stanley variant, core/controllers.py:234-324.

TEST INFRASTRUCTURE.  ``class StanleyController`` defines ``__init__`` / ``compute_steering`` twice; Python keeps the
last binding, so the rate-limited / low-pass version at :234-324 is unreachable through the class.
To pin its restatement anyway, this script extracts the FIRST ``__init__`` and ``compute_steering`` definitions from
the reference's own source with ``ast`` and executes them, unmodified, inside the reference module's namespace.
Run in the build container (needs /root/reference):  python -m oracle.gen_golden_shadowed
"""
from __future__ import annotations

import ast
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "shadowed_stanley_kat.json")


def hx(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return [hx(e) for e in v]
    return float(v).hex()


def extract_shadowed_class(controllers_module):
    src_path = controllers_module.__file__
    tree = ast.parse(open(src_path).read(), filename=src_path)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "StanleyController")
    first = {}
    for node in cls.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("__init__", "compute_steering") and node.name not in first:
            first[node.name] = node
    assert first["compute_steering"].lineno < 260, "expected the shadowed definition at core/controllers.py:252"
    assert [a.arg for a in first["compute_steering"].args.args] == ["self", "state", "path", "target_ind", "dt"]
    new_cls = ast.ClassDef(name="ShadowedStanley", bases=[], keywords=[], body=[first["__init__"], first["compute_steering"]],
                           decorator_list=[])
    mod = ast.Module(body=[new_cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = dict(controllers_module.__dict__)
    exec(compile(mod, src_path, "exec"), ns)
    return ns["ShadowedStanley"], (first["__init__"].lineno, first["compute_steering"].end_lineno)


def main():
    sys.path.insert(0, os.path.dirname(HERE))
    from oracle import ref_shim
    ref = ref_shim.load_reference()
    C, CS, conf = ref.controllers, ref.car_state, ref.conf
    Shadowed, lines = extract_shadowed_class(C)
    G = {"_meta": {"generator": "oracle/gen_golden_shadowed.py", "source_lines": list(lines),
                   "note": "first (shadowed) __init__/compute_steering of StanleyController, executed unmodified via ast"}}
    ctl = Shadowed()
    G["defaults"] = {"k": hx(ctl.k), "k_soft": hx(ctl.k_soft), "max_steer_rate": hx(ctl.max_steer_rate),
                     "alpha": hx(ctl.alpha), "lookahead_distance": hx(ctl.lookahead_distance)}
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    cx, cy = 40.0 * np.cos(th), 20.0 * np.sin(th)
    path = np.column_stack((np.arange(720), cx, cy))
    rng = np.random.default_rng(2402)
    rows = []
    for _ in range(160):
        i = int(rng.integers(0, 720))
        x, y = cx[i] + rng.normal(0, 0.6), cy[i] + rng.normal(0, 0.6)
        yaw = math.atan2(20.0 * math.cos(th[i]), -40.0 * math.sin(th[i])) + rng.normal(0, 0.4)
        v = float(rng.uniform(0.0, 10.0))
        k, ks = float(rng.uniform(0.2, 2.0)), float(rng.uniform(0.01, 1.0))
        rate, alpha = float(rng.uniform(0.1, 2.0)), float(rng.uniform(0.1, 1.0))
        prev, dt = float(rng.uniform(-0.5, 0.5)), float(rng.choice([0.05, 0.1, 0.02]))
        c = Shadowed(k=k, k_soft=ks, max_steer_rate=rate, alpha=alpha)
        c.previous_steering = prev
        s, ind = c.compute_steering(CS.State(x=x, y=y, yaw=yaw, v=v), path, 0, dt)
        rows.append({"state": hx([x, y, yaw, v]), "k": hx(k), "k_soft": hx(ks), "rate": hx(rate), "alpha": hx(alpha),
                     "prev": hx(prev), "dt": hx(dt), "steering": hx(s), "ind": int(ind), "prev_out": hx(c.previous_steering)})
    G["single"] = rows
    # the optional look-ahead advance (:276-281): a separate generator stream, so the rows above stay what they were
    rng2 = np.random.default_rng(2403)
    rows = []
    for _ in range(96):
        i = int(rng2.integers(0, 720))
        x, y = cx[i] + rng2.normal(0, 0.6), cy[i] + rng2.normal(0, 0.6)
        yaw = math.atan2(20.0 * math.cos(th[i]), -40.0 * math.sin(th[i])) + rng2.normal(0, 0.4)
        v = float(rng2.uniform(0.0, 10.0))
        k, ks = float(rng2.uniform(0.2, 2.0)), float(rng2.uniform(0.01, 1.0))
        prev, dt = float(rng2.uniform(-0.5, 0.5)), float(rng2.choice([0.05, 0.1]))
        la = float(rng2.uniform(0.3, 6.0))
        c = Shadowed(k=k, k_soft=ks, lookahead_distance=la)
        c.previous_steering = prev
        # the last third runs on an OPEN path that ends shortly after the car: the advance stops at len(cx) - 1 (:277)
        pth = path if len(rows) < 64 else path[: min(720, i + 6)]
        s, ind = c.compute_steering(CS.State(x=x, y=y, yaw=yaw, v=v), pth, 0, dt)
        rows.append({"state": hx([x, y, yaw, v]), "k": hx(k), "k_soft": hx(ks), "prev": hx(prev), "dt": hx(dt),
                     "lookahead": hx(la), "n_path": len(pth), "steering": hx(s), "ind": int(ind),
                     "prev_out": hx(c.previous_steering)})
    G["single_lookahead"] = rows
    # closed loop: shadowed Stanley (constructor defaults) + PID + State.update, 400 periods
    ref_shim.FixedDtPID.dt = conf.dt
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    steer = Shadowed()
    accel = C.AccelerationPIDController(conf.kp_accel, conf.ki_accel, conf.kd_accel, conf.TARGET_SPEED)
    init = (40.0, 0.3, math.pi / 2, 5.0, 0.0, 0.0)
    st = CS.State(*init)
    traj, steers, accels, tis = [], [], [], []
    for _ in range(400):
        traj.append([st.x, st.y, st.yaw, st.v, st.r, st.beta])
        s, ti = steer.compute_steering(st, path, 0, conf.dt)
        kappa = curve[ti] if ti < len(curve) else curve[-1]
        a = accel.compute_acceleration(st.v, kappa)
        st.update(a, s)
        steers.append(s); accels.append(a); tis.append(int(ti))
    G["loop"] = {"init": hx(init), "traj": hx(traj), "steer": hx(steers), "accel": hx(accels), "ti": tis,
                 "final": hx([st.x, st.y, st.yaw, st.v, st.r, st.beta])}
    with open(OUT, "w") as f:
        json.dump(G, f, indent=0, separators=(",", ":"))
    print("wrote", OUT, os.path.getsize(OUT), "bytes; source lines", lines)


if __name__ == "__main__":
    main()
