"""Static vs dynamic episode scheduling on a heavy-tailed batch (skidpad Monte Carlo, Stanley kinds: mean 130 steps,
maximum 755 / 1 500), many waves long.  usage: python scripts/diag_dynamic.py [episodes]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgr_pathplanning_control_b200 as bgr  # noqa: E402
from bgr_pathplanning_control_b200 import _abi as abi  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
ta = bgr.tracks.skidpad(2048)
track = bgr.Track.from_arrays(ta, device=0, cone_reach=4.0)
noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
for steer, accel, model in (("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic"), ("pure_pursuit", "pid", "kinematic")):
    init = bgr.initial_states(ta, E, index=0, v0=5.0 if model == "dynamic" else 0.0)
    init[:, abi.YAW] = 0.0
    p = torch.from_numpy(bgr.sweeps.monte_carlo_params(E)).to(dev)
    s = torch.from_numpy(init).to(dev)
    for static in (True, False):
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, dynamic_schedule=not static, **noise)
        for _ in range(2):
            r = bgr.rollout(track, cfg, p, s, aggregate=True, time_kernel=True)
        torch.cuda.synchronize()
        steps = float(r.aggregate[abi.A_STEPS])
        print(f"{steer}+{accel}+{model:9s} E={E} {'static ' if static else 'dynamic'}: kernel {r.kernel_ms:8.3f} ms  "
              f"{steps:.3e} steps -> {steps / r.kernel_ms / 1e-3:.3e} steps/s")
