import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import bgr_pathplanning_control_b200 as bgr
for n in (40, 1024, 2048, 6000, 12000):
    th = np.linspace(0, 2 * np.pi, n, endpoint=False)
    x, y = 60 * np.cos(th), 35 * np.sin(th)
    t0 = time.perf_counter(); t = bgr.Track(x, y, None, closed=True); dt = time.perf_counter() - t0
    print(n, f"{dt*1e3:.1f} ms", t.info()["rival_total"]); t.close()
ta = bgr.tracks.skidpad(2048)
t0 = time.perf_counter(); t = bgr.Track.from_arrays(ta); print("skidpad 2048", f"{(time.perf_counter()-t0)*1e3:.1f} ms")
