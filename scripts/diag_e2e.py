import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
import torch, torch.distributed as dist
import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E = 1 << 20
ta = bgr.tracks.stadium_oval(1024)
track = bgr.Track.from_arrays(ta, device=local)
for steer, accel, model, laps in (("pure_pursuit", "pid", "kinematic", 8), ("stanley", "lqg", "dynamic", 1), ("stanley", "pid", "dynamic", 1), ("pure_pursuit", "lqg", "kinematic", 8)):
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=2000, laps=laps)
    p = torch.from_numpy(bgr.sweeps.stanley_lqg_params(E)).pin_memory().numpy()
    i = torch.from_numpy(bgr.initial_states(ta, E, v0=5.0)).pin_memory().numpy()
    out = bgr.RolloutResult(torch.empty((E, 12), dtype=torch.float32).pin_memory().numpy(), torch.empty((E, 10), dtype=torch.int32).pin_memory().numpy())
    ts = []
    for k in range(6):
        if world > 1: dist.barrier()
        t0 = time.perf_counter(); bgr.rollout(track, cfg, p, i, aggregate=True, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"rank {rank} {steer}+{accel}+{model}: e2e ms per call", " ".join(f"{t:.1f}" for t in ts), flush=True)
if world > 1: dist.destroy_process_group()
