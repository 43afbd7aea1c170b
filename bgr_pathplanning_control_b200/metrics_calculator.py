"""Lap metrics with the reference's function names (preformance_analysis/metrics_calculator.py:1-8).

The reference computes them with pandas over a logged CSV (simulation_logger.py:8-10).  Here the kernels
accumulate them per episode while stepping, so each function accepts a ``RolloutResult`` (returns per-episode
tensors) -- or, for drop-in use on logged data, any mapping/DataFrame with the reference's columns, which is
forwarded to the GPU as a one-episode reduction.
"""
from __future__ import annotations

import numpy as np

from . import _abi as abi

CSV_COLUMNS = ["timestamp", "x", "y", "velocity", "steering_angle", "throttle", "brake", "lateral_error",
               "heading_error", "yaw_rate", "lateral_accel"]            # simulation_logger.py:8-10


def _is_result(data):
    return hasattr(data, "metrics") and hasattr(data, "status")


def calculate_lateral_error(data):
    """(mean, std ddof=1) of ``lateral_error``   (:1-2)"""
    if _is_result(data):
        return data.metrics[:, abi.M_LAT_MEAN], data.metrics[:, abi.M_LAT_STD]
    return _column_mean_std(data["lateral_error"])


def calculate_heading_error(data):
    """(mean, std ddof=1) of ``heading_error``   (:4-5)"""
    if _is_result(data):
        return data.metrics[:, abi.M_HEAD_MEAN], data.metrics[:, abi.M_HEAD_STD]
    return _column_mean_std(data["heading_error"])


def calculate_lap_time(data):
    """``timestamp.iloc[-1] - timestamp.iloc[0]``   (:7-8)"""
    if _is_result(data):
        return data.metrics[:, abi.M_LAP_TIME]
    ts = data["timestamp"]
    ts = ts.to_numpy() if hasattr(ts, "to_numpy") else np.asarray(ts)
    return ts[-1] - ts[0]


def _column_mean_std(col):
    """Mean / sample std of one logged column on the GPU (torch reduction = tensor plumbing; the column was
    produced by ``trajectory_frame``)."""
    import torch

    a = np.array(col.to_numpy() if hasattr(col, "to_numpy") else col, dtype=np.float64)   # writable copy
    if not torch.cuda.is_available():
        raise RuntimeError("metrics_calculator needs a CUDA device (no CPU fallback)")
    t = torch.as_tensor(a, dtype=torch.float64, device="cuda")
    mean = t.mean()
    std = t.std(unbiased=True) if t.numel() > 1 else torch.tensor(float("nan"))
    return float(mean), float(std)


def trajectory_frame(result, episode=0, dt=0.05, max_accel=1.5, max_decel=-2.0):
    """One episode's trajectory dump as a pandas DataFrame in the reference's CSV schema
    (simulation_logger.py:8-10).  throttle / brake follow the normalisation of
    providers/sim/sim_util.py:197-209 verbatim: throttle = clip(a / MAX_ACCEL, 0, 1) for a >= 0; brake =
    clip(-a / MAX_DECEL, 0, 1) for a < 0 -- with MAX_DECEL negative that quotient is negative, so the
    reference's brake command is always 0; reproduced, not fixed."""
    import pandas as pd

    if result.traj is None:
        raise ValueError("rollout was run without traj=True")
    steps = int(result.status[episode, abi.S_STEPS])
    tr = result.traj[episode, :steps]
    tr = tr.cpu().numpy() if hasattr(tr, "cpu") else np.asarray(tr)
    acc = tr[:, abi.T_ACCEL]
    v = tr[:, abi.T_V]
    r = tr[:, abi.T_R]
    return pd.DataFrame({
        "timestamp": np.arange(steps) * dt,
        "x": tr[:, abi.T_X], "y": tr[:, abi.T_Y], "velocity": v,
        "steering_angle": tr[:, abi.T_STEER],
        "throttle": np.where(acc >= 0, np.clip(acc / max_accel, 0.0, 1.0), 0.0),
        "brake": np.where(acc >= 0, 0.0, np.clip(-acc / max_decel, 0.0, 1.0)),
        "lateral_error": tr[:, abi.T_E_CT], "heading_error": tr[:, abi.T_THETA_E],
        "yaw_rate": r, "lateral_accel": v * r,
    }, columns=CSV_COLUMNS)


def load_simulation_data(filename):
    """preformance_analysis/data_loader.py:3-4."""
    import pandas as pd
    return pd.read_csv(filename)
