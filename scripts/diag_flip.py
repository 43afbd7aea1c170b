import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from oracle import cport, loop
ta = bgr.tracks.stadium_oval(1024)
track = bgr.Track.from_arrays(ta, cone_reach=4.0)
ot = loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones)
steer, accel, model, laps = "pure_pursuit", "pid", "kinematic", 8
E, S = 768, 2000
params = bgr.sweeps.pp_sweep_params(E, start=32 * 16384 + 4242)
init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3, 0.3, E))
cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=laps, audit=True)
res = bgr.rollout(track, cfg, params, init, traj=True)
spec = loop.RolloutSpec(steer=steer, accel=accel, model=model, max_steps=S, laps=laps)
m, s, tr = cport.rollout(ot, spec, params.astype(np.float64), init, traj=True)
dpos = np.hypot(res.metrics[:, abi.M_X] - m[:, 6], res.metrics[:, abi.M_Y] - m[:, 7])
ties = res.status[:, abi.S_NEAR_TIES]
print("clean n", (ties == 0).sum(), "clean dpos max", dpos[ties == 0].max(), "flagged n", (ties > 0).sum(), "flagged dpos quantiles", np.quantile(dpos[ties > 0], [0.5, 0.9, 0.99, 1.0]))
print("flagged with dpos>1e-4:", ((ties > 0) & (dpos > 1e-4)).sum())
for e in np.argsort(-dpos)[:4]:
    g = res.traj[e]; o = tr[e]
    dti = np.nonzero(g[:, 8] != o[:, 8])[0]
    dn = np.nonzero(g[:, 9] != o[:, 9])[0]
    d = np.hypot(g[:, 0] - o[:, 0], g[:, 1] - o[:, 1])
    print("ep", e, "dpos", dpos[e], "ties", ties[e], "first_tie", res.status[e, abi.S_FIRST_TIE], "ti mismatch steps", dti[:6], len(dti), "near mismatch", dn[:6], len(dn))
    if len(dti):
        k = dti[0]
        print("   at step", k, "gpu ti", g[k, 8], "oracle ti", o[k, 8], "d before", d[max(0, k - 1)], "after 10", d[min(S - 1, k + 10)], "end", d[-1], "Lf", params[e, 0])
        tix = int(o[k, 8]) % ta.n
        for cand in (int(g[k, 8]) % ta.n, int(o[k, 8]) % ta.n):
            print("     dist to wp", cand, np.hypot(ta.x[cand] - o[k, 0], ta.y[cand] - o[k, 1]), "vs Lf", float(params[e, 0]))
