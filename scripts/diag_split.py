import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
ta = bgr.tracks.skidpad(2048)
track = bgr.Track.from_arrays(ta, cone_reach=4.0)
dev = torch.device("cuda:0")
per = 174592
noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
for steer, accel, model in [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]:
    init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0); init[:, abi.YAW] = 0.0
    params = bgr.sweeps.monte_carlo_params(per)
    dp, di = torch.from_numpy(params).to(dev), torch.from_numpy(init).to(dev)
    for split in (0, 192, 0, 192, 512):
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, split_steps=split, **noise)
        bgr.rollout(track, cfg, dp, di)
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = bgr.rollout(track, cfg, dp, di); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = min(ts); print("   all:", " ".join(f"{x:.2f}" for x in ts))
        st = r.status.cpu().numpy()
        steps = st[:, abi.S_STEPS]
        print(f"{steer}+{accel}+{model} split={split}: {ms:.2f} ms, steps mean {steps.mean():.0f}, survivors beyond split {(steps > split).mean() if split else 0:.4f}")
