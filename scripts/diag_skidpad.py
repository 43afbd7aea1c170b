import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
ta = bgr.tracks.skidpad(2048)
track = bgr.Track.from_arrays(ta, cone_reach=4.0)
yaw, guard = track.tables()
print(track.info()); print("guard quantiles", np.quantile(guard, [0, 0.1, 0.5, 0.9, 1.0]))
dev = torch.device("cuda:0")
per = 174592
for steer, accel, model in [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]:
    for noisy in (False, True):
        kw = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, noise_seed=1) if noisy else {}
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, **kw)
        init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0); init[:, abi.YAW] = 0.0
        params = bgr.sweeps.monte_carlo_params(per)
        dp, di = torch.from_numpy(params).to(dev), torch.from_numpy(init).to(dev)
        bgr.rollout(track, cfg, dp, di)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = bgr.rollout(track, cfg, dp, di, time_kernel=True)
        torch.cuda.synchronize()
        st = r.status.cpu().numpy()
        steps = st[:, abi.S_STEPS]
        print(f"{steer}+{accel}+{model} noisy={noisy}: kernel {r.kernel_ms:.2f} ms, steps mean {steps.mean():.0f} max {steps.max()}, "
              f"rate {steps.sum() / r.kernel_ms / 1e6:.2f} Gsteps/s, fallbacks/episode {st[:, abi.S_FALLBACKS].mean():.1f}, "
              f"flags {np.bincount(st[:, abi.S_FLAGS], minlength=9)[[1,2,4,8]]}, max|e| median {np.median(r.metrics.cpu().numpy()[:, abi.M_MAX_ABS_LAT]):.2f}")
