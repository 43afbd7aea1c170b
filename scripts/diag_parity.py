"""Diagnostic (not a test): fp32 production kernel vs the C oracle twin at scale, by oracle conditioning bin."""
import sys
import numpy as np
sys.path.insert(0, '/root/repo')
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from oracle import cport, loop
ta = bgr.tracks.stadium_oval(1024)
track = bgr.Track.from_arrays(ta, cone_reach=4.0)
ot = loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones)
for steer, accel, model, laps in [("pure_pursuit", "pid", "kinematic", 8), ("stanley", "lqg", "dynamic", 1),
                                  ("pure_pursuit", "lqg", "dynamic", 8), ("stanley", "pid", "kinematic", 1)]:
    E, S = 768, 2000
    params = (bgr.sweeps.pp_sweep_params(E, start=32 * 16384 + 4242) if steer == "pure_pursuit" else bgr.sweeps.stanley_lqg_params(E))
    init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3, 0.3, E), v0=5.0 if model == 'dynamic' else 0.0)
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=laps, audit=True)
    res = bgr.rollout(track, cfg, params, init)
    spec = loop.RolloutSpec(steer=steer, accel=accel, model=model, max_steps=S, laps=laps)
    p64 = params.astype(np.float64)
    m, s, _ = cport.rollout(ot, spec, p64, init)
    nud = init.copy(); nud[:, 0] += 1e-7
    m2, _, _ = cport.rollout(ot, spec, p64, nud)
    resp = np.hypot(m2[:, 6] - m[:, 6], m2[:, 7] - m[:, 7])
    on = (m[:, 5] < 1.5) & (res.metrics[:, abi.M_MAX_ABS_LAT] < 1.5)
    ties = res.status[:, abi.S_NEAR_TIES]
    cols = [abi.S_FLAGS, abi.S_STEPS, abi.S_TARGET_IND, abi.S_NEAREST_IND, abi.S_LAPS]
    mism = np.any(res.status[:, cols] != s[:, cols], axis=1)
    dpos = np.hypot(res.metrics[:, abi.M_X] - m[:, 6], res.metrics[:, abi.M_Y] - m[:, 7])
    dyaw = np.abs(res.metrics[:, abi.M_YAW] - m[:, 8])
    print(f"== {steer}+{accel}+{model}: on_track {on.mean():.3f}, ties/episode {ties[on].mean() if on.any() else 0:.2f}, "
          f"index mismatches: all {mism.sum()}, on-track {(mism & on).sum()}, on-track & unflagged {(mism & on & (ties == 0)).sum()}")
    edges = [0, 3e-7, 1e-6, 3e-6, 1e-5, 1e-4, 1e-2, 1e9]
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = on & (resp >= lo) & (resp < hi)
        if sel.any():
            print(f"   nudge response [{lo:.0e},{hi:.0e}): n={sel.sum():4d} dpos med {np.median(dpos[sel]):.2e} max {dpos[sel].max():.2e} "
                  f"dyaw max {dyaw[sel].max():.2e} dlat_mean max {np.abs(res.metrics[sel, 0] - m[sel, 0]).max():.2e} "
                  f"dlat_std max {np.abs(res.metrics[sel, 1] - m[sel, 1]).max():.2e} mism {(mism & sel).sum()} unflagged-mism {(mism & sel & (ties == 0)).sum()}")
    fin = np.isfinite(res.metrics[:, 1]) & np.isfinite(m[:, 1])
    print("   aggregate: median lat_std gpu/oracle", np.median(res.metrics[fin, 1]), np.median(m[fin, 1]), "on-track frac gpu/oracle",
          (res.metrics[:, abi.M_MAX_ABS_LAT] < 1.5).mean(), (m[:, 5] < 1.5).mean())
