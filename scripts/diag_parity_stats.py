"""End-pose parity statistics of the fp32 production kernel against the C oracle twin at scale (768 episodes x 2000
steps of the bench sweeps): median / 90 % / max position and heading deviation over well-conditioned, unflagged,
on-track episodes.  Diagnostic for numerics trade-offs (e.g. polynomial vs MUFU sin / cos)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from oracle import cport, loop

ta = bgr.tracks.stadium_oval(1024)
track = bgr.Track.from_arrays(ta, cone_reach=4.0)
E, S = 768, 2000
for steer, accel, model, laps in (("pure_pursuit", "pid", "kinematic", 8), ("stanley", "pid", "kinematic", 1),
                                  ("stanley", "lqg", "dynamic", 1)):
    params = (bgr.sweeps.pp_sweep_params(E, start=32 * 16384 + 4242) if steer == "pure_pursuit"
              else bgr.sweeps.stanley_lqg_params(E))
    init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3, 0.3, E), v0=5.0 if model == "dynamic" else 0.0)
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=laps, audit=True)
    res = bgr.rollout(track, cfg, params, init)
    spec = loop.RolloutSpec(steer=steer, accel=accel, model=model, max_steps=S, laps=laps)
    tr = loop.Track(ta.x, ta.y, ta.kappa, ta.closed, ta.cones)
    m, s, _ = cport.rollout(tr, spec, params.astype(np.float64), init)
    nudged = init.copy(); nudged[:, abi.X] += 1e-7
    m2, _, _ = cport.rollout(tr, spec, params.astype(np.float64), nudged)
    well = np.hypot(m2[:, 6] - m[:, 6], m2[:, 7] - m[:, 7]) < 1e-5
    good = well & (m[:, 5] < 1.5) & (res.status[:, abi.S_NEAR_TIES] == 0) & (res.metrics[:, abi.M_MAX_ABS_LAT] < 1.5)
    dpos = np.hypot(res.metrics[:, abi.M_X] - m[:, 6], res.metrics[:, abi.M_Y] - m[:, 7])[good]
    dyaw = np.abs(res.metrics[:, abi.M_YAW] - m[:, 8])[good]
    print(f"{steer}+{accel}+{model}: {good.sum()} episodes; dpos median {np.median(dpos):.2e} p90 {np.quantile(dpos, .9):.2e} "
          f"max {dpos.max():.2e} m; dyaw median {np.median(dyaw):.2e} max {dyaw.max():.2e} rad")
