import numpy as np, sys
sys.path.insert(0,'/root/repo')
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from oracle import cport, loop
ta = bgr.tracks.stadium_oval(1024)
track = bgr.Track.from_arrays(ta, cone_reach=4.0)
ot = loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones)
E,S=768,2000
params = bgr.sweeps.stanley_lqg_params(E); params[:, abi.P_MU] = np.maximum(params[:, abi.P_MU], 0.9)
init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3,0.3,E), v0=5.0)
spec = loop.RolloutSpec(steer="stanley", accel="lqg", model="dynamic", max_steps=S, laps=1)
m,s,tr = cport.rollout(ot, spec, params.astype(np.float64), init, traj=True)
for prec in ("fp32","fp64"):
    cfg = bgr.RolloutConfig(steer="stanley", accel="lqg", model="dynamic", max_steps=S, laps=1, audit=True, precision=prec)
    res = bgr.rollout(track, cfg, params, init, traj=True)
    dpos = np.hypot(res.metrics[:, abi.M_X]-m[:,6], res.metrics[:, abi.M_Y]-m[:,7])
    on = (m[:,5] < 1.5)
    print(prec, "on", on.mean())
    for lo,hi in ((0,0.3),(0.3,0.6),(0.6,1.0),(1.0,1.5),(1.5,100)):
        sel=(m[:,5]>=lo)&(m[:,5]<hi)
        if sel.any(): print("  max|e| in [%.1f,%.1f): n=%d dpos median %.2e max %.2e  ties mean %.2f"%(lo,hi,sel.sum(),np.median(dpos[sel]),dpos[sel].max(), res.status[sel, abi.S_NEAR_TIES].mean()))
    # worst clean episode: where does it start diverging
    sel=np.nonzero(on)[0]; w=sel[np.argmax(dpos[sel])]
    d=np.hypot(res.traj[w,:,0]-tr[w,:,0], res.traj[w,:,1]-tr[w,:,1])
    print("  worst on-track ep", w, "params k,ks,mu,cf,cr", params[w,[6,7,8,9,10]], "max|e|", m[w,5], "dpos at steps 100,500,1000,1500,1999:", d[[100,500,1000,1500,1999]])
    print("   beta range", tr[w,:,5].min(), tr[w,:,5].max(), " r range", tr[w,:,4].min(), tr[w,:,4].max(), "v range", tr[w,:,3].min(), tr[w,:,3].max())
    b=np.abs(res.traj[w,:,5]-tr[w,:,5]); print("   |dbeta| at 100,500,1000,1999", b[[100,500,1000,1999]], "steer diff", np.abs(res.traj[w,:,6]-tr[w,:,6])[[100,500,1000,1999]])
    # saturation: fraction of steps with steer at the clip
    print("   steer clipped frac", (np.abs(tr[w,:,6])>=bgr.conf.MAX_STEER-1e-9).mean())
