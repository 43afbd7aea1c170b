"""Config-5 diagnostics: where the host-buffer path loses time against the device-tensor path (per kind and grouped)."""
import os
import sys
import time

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
jobs = []
for g, (steer, accel, model) in enumerate(kinds):
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, episode_id_base=g * per,
                            dynamic_schedule=steer == "stanley", **noise)
    init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0)
    init[:, abi.YAW] = 0.0
    params = bgr.sweeps.monte_carlo_params(per, start=g * per)
    hp, hi = torch.from_numpy(params).pin_memory().numpy(), torch.from_numpy(init).pin_memory().numpy()
    ho = bgr.RolloutResult(torch.empty((per, abi.N_METRICS), dtype=torch.float32).pin_memory().numpy(),
                           torch.empty((per, abi.N_STATUS), dtype=torch.int32).pin_memory().numpy(), None, None)
    jobs.append((cfg, torch.from_numpy(params).to(dev), torch.from_numpy(init).to(dev), hp, hi, ho))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


for g, j in enumerate(jobs):
    d = timeit(lambda: bgr.rollout(track, j[0], j[1], j[2], aggregate=True))
    h = timeit(lambda: bgr.rollout(track, j[0], j[3], j[4], aggregate=True, out=j[5]))
    print(f"kind {g} {kinds[g]}: device tensors {d:.3f} ms   host buffers {h:.3f} ms")
d = timeit(lambda: bgr.rollout_groups(track, [(j[0], j[1], j[2]) for j in jobs], aggregate=True, order=[2, 1, 0]))
h = timeit(lambda: bgr.rollout_groups(track, [(j[0], j[3], j[4]) for j in jobs], aggregate=True, outs=[j[5] for j in jobs],
                                      order=[2, 1, 0]))
print(f"grouped: device tensors {d:.3f} ms   host buffers {h:.3f} ms")
h = timeit(lambda: bgr.rollout_groups(track, [(j[0], j[3], j[4]) for j in jobs], aggregate=True, outs=[j[5] for j in jobs],
                                      order=[0, 2, 1]))
print(f"grouped, pure pursuit first: host buffers {h:.3f} ms")
# copies alone
t = timeit(lambda: [torch.from_numpy(j[3]).to(dev, non_blocking=True) for j in jobs] + [torch.from_numpy(j[4]).to(dev, non_blocking=True) for j in jobs])
print(f"uploads alone (all kinds): {t:.3f} ms")

# wall-clock intervals of the three host-buffer calls of one grouped step (are they concurrent?)
import threading
from bgr_pathplanning_control_b200 import engine as _eng
marks = {}
orig = _eng._group_worker


def traced(track_, config, params, init_state, aggregate, out):
    t0 = time.perf_counter()
    r = orig(track_, config, params, init_state, aggregate, out)
    marks[config.steer + "+" + config.accel] = (t0, time.perf_counter(), threading.get_ident(), bgr.last_kernel_ms())
    return r


_eng._group_worker = traced
for _ in range(3):
    marks.clear()
    T0 = time.perf_counter()
    bgr.rollout_groups(track, [(j[0], j[3], j[4]) for j in jobs], aggregate=True, outs=[j[5] for j in jobs], order=[2, 1, 0])
    T1 = time.perf_counter()
    print(f"grouped call {1e3 * (T1 - T0):.3f} ms: " + "  ".join(
        f"{k}: [{1e3 * (a - T0):.2f}, {1e3 * (b - T0):.2f}] thread {t % 10000} last kernel {km:.2f} ms" for k, (a, b, t, km) in marks.items()))
