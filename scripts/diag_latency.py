import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import bgr_pathplanning_control_b200 as bgr
ta = bgr.tracks.stadium_oval(1024)
path = np.column_stack([np.arange(ta.n), ta.x, ta.y])
st = type("S", (), dict(x=0.1, y=0.05, yaw=0.01, v=5.0))()
pp, sl = bgr.PurePursuitController(), bgr.StanleyController()
pid = bgr.AccelerationPIDController(0.35, 0.4, 0.3, 6.9)
lqg = bgr.LQGAccelerationController(0.05)
mpc = bgr.get_steering_controller("mpc")
s = bgr.State(0.0, 0.0, 0.0, 5.0)
for name, fn in (("PurePursuit.compute_steering", lambda: pp.compute_steering(st, path, 3)),
                 ("Stanley.compute_steering", lambda: sl.compute_steering(st, path, 0)),
                 ("PID.compute_acceleration", lambda: pid.compute_acceleration(5.0, 0.02)),
                 ("LQG.compute_acceleration", lambda: lqg.compute_acceleration(5.0, 0.02)),
                 ("State.update", lambda: s.update(0.1, 0.01)),
                 ("MPC.compute_control", lambda: mpc.compute_control(st, path[:60]))):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(200): fn()
    print(f"{name:32s} {(time.perf_counter() - t0) / 200 * 1e6:8.1f} us per call")
