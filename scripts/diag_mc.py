"""Config-5 diagnostics on the GPU: per-kind kernel times of the skidpad Monte Carlo (device tensors), track info,
and the effect of BGR_ATTR_IN_SMEM / wave-sized batches.  usage: python scripts/diag_mc.py [episodes_per_kind]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgr_pathplanning_control_b200 as bgr  # noqa: E402
from bgr_pathplanning_control_b200 import _abi as abi  # noqa: E402

per = int(sys.argv[1]) if len(sys.argv) > 1 else 174592
dev = torch.device("cuda", 0)
ta = bgr.tracks.skidpad(2048)
track = bgr.Track.from_arrays(ta, device=0, cone_reach=4.0)
print("env BGR_ATTR_IN_SMEM =", repr(os.environ.get("BGR_ATTR_IN_SMEM")), "track info", track.info())
noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
kinds = [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]
rr = int(os.environ.get("BGR_MC_ROW_RANK", "0"))
for g0, (steer, accel, model) in enumerate(kinds):
    g = g0 + 3 * rr
    for label, extra in (("noise+cones", noise), ("noise only", {**noise, "hit_radius": 0.0}),
                         ("clean (lean build)", {})):
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, episode_id_base=g * per,
                                dynamic_schedule=os.environ.get("BGR_MC_DYNAMIC", "1" if steer == "stanley" else "0") == "1", **extra)
        init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0)
        init[:, abi.YAW] = 0.0
        p = torch.from_numpy(bgr.sweeps.monte_carlo_params(per, start=g * per)).to(dev)
        s = torch.from_numpy(init).to(dev)
        for _ in range(2):
            r = bgr.rollout(track, cfg, p, s, aggregate=True, time_kernel=True)
        torch.cuda.synchronize()
        st = r.status.cpu().numpy()
        steps = float(r.aggregate[abi.A_STEPS])
        print(f"{steer}+{accel}+{model:9s} {label:20s}: kernel {r.kernel_ms:8.3f} ms  steps {steps:.3e}  -> {steps / r.kernel_ms / 1e-3:.3e} steps/s  "
              f"mean steps/episode {st[:, abi.S_STEPS].mean():.0f} max {st[:, abi.S_STEPS].max()}  fallbacks/episode {st[:, abi.S_FALLBACKS].mean():.2f}")
