"""Config-5 diagnostics: distribution of episode lengths and end reasons per controller kind."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgr_pathplanning_control_b200 as bgr  # noqa: E402
from bgr_pathplanning_control_b200 import _abi as abi  # noqa: E402

per = 174592
dev = torch.device("cuda", 0)
ta = bgr.tracks.skidpad(2048)
track = bgr.Track.from_arrays(ta, device=0, cone_reach=4.0)
noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
kinds = [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]
rr = int(os.environ.get("BGR_MC_ROW_RANK", "0"))
for g0, (steer, accel, model) in enumerate(kinds):
    g = g0 + 3 * rr
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, episode_id_base=g * per,
                            dynamic_schedule=steer == "stanley", **noise)
    init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0)
    init[:, abi.YAW] = 0.0
    p = torch.from_numpy(bgr.sweeps.monte_carlo_params(per, start=g * per)).to(dev)
    s = torch.from_numpy(init).to(dev)
    r = bgr.rollout(track, cfg, p, s, aggregate=True, time_kernel=True)
    r = bgr.rollout(track, cfg, p, s, aggregate=True, time_kernel=True)
    print("   kernel ms", r.kernel_ms)
    st = r.status.cpu().numpy()
    steps = st[:, abi.S_STEPS]
    print(steer, accel, model, "steps pct [50 90 95 99 99.9 100]:", np.percentile(steps, [50, 90, 95, 99, 99.9, 100]).astype(int),
          "mean", steps.mean())
    for thr in (140, 160, 200, 256, 400, 800):
        print(f"   share of episodes longer than {thr}: {(steps > thr).mean():.4f}; their share of all steps: {steps[steps > thr].sum() / steps.sum():.3f}")
    fb = st[:, abi.S_FALLBACKS]
    print("   fallbacks: total", int(fb.sum()), "max per episode", int(fb.max()), "episodes with > 100:", int((fb > 100).sum()),
          "steps of those:", st[fb > 100, abi.S_STEPS][:10])
    flags, cnt = np.unique(st[:, abi.S_FLAGS], return_counts=True)
    print("   end flags:", dict(zip(flags.tolist(), cnt.tolist())))
    w = steps[: per // 32 * 32].reshape(-1, 32)
    print(f"   static warps: mean of per-warp max {w.max(1).mean():.0f} vs mean {steps.mean():.0f} -> lane utilisation {steps.mean() / w.max(1).mean():.2f}")
    c = steps[: per // 256 * 256].reshape(-1, 256)
    print(f"   static CTAs: mean of per-CTA max {c.max(1).mean():.0f}")
