"""Where does the time between two bgr_rollout launches go?  Device tensors, config 2, 12 steps per variant."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bgr_pathplanning_control_b200 as bgr  # noqa: E402
from bgr_pathplanning_control_b200 import _abi as abi  # noqa: E402

dev = torch.device("cuda", 0)
E, S = 1 << 20, 2000
ta = bgr.tracks.stadium_oval(1024)
track = bgr.Track.from_arrays(ta, device=0)
cfg = bgr.RolloutConfig(steer="pure_pursuit", accel="pid", model="kinematic", max_steps=S, laps=8)
p = torch.from_numpy(bgr.sweeps.pp_sweep_params(E)).to(dev)
s = torch.from_numpy(bgr.initial_states(ta, E)).to(dev)
out = bgr.RolloutResult(torch.empty((E, abi.N_METRICS), dtype=torch.float32, device=dev),
                        torch.empty((E, abi.N_STATUS), dtype=torch.int32, device=dev), None,
                        torch.empty(abi.N_AGG, dtype=torch.float64, device=dev))
for label, kw in (("aggregate, out given", dict(aggregate=True, out=out)), ("no aggregate, out given", dict(aggregate=False, out=out)),
                  ("aggregate, fresh outputs", dict(aggregate=True))):
    for _ in range(3):
        bgr.rollout(track, cfg, p, s, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(12):
        r = bgr.rollout(track, cfg, p, s, **kw)
    b.record()
    t_host = (time.perf_counter() - t0) * 1e3 / 12
    torch.cuda.synchronize()
    print(f"{label:28s}: {a.elapsed_time(b) / 12:7.3f} ms per step on the device, {t_host:6.3f} ms of host time per call, "
          f"kernel alone {bgr.last_kernel_ms():7.3f} ms")
