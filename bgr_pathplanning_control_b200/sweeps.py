"""Parameter tensors and initial states of the benchmark configurations.

Host numpy only (inputs); drawn in float64 and cast, so the oracle and the GPU see the same rows.
"""
from __future__ import annotations

import numpy as np

from . import _abi as abi
from . import engine
from .vehicle_config import conf


def pp_sweep_params(n_episodes=1 << 20, start=0):
    """Config 2: pure-pursuit look-ahead / PID-gain grid.  Lf in linspace(2, 9, 64) x kp in
    linspace(0.1, 1, 32) x ki in linspace(0, 0.8, 16) x kd in linspace(0, 0.6, 32) = 1 048 576 combos,
    kd fastest.  Rows ``start .. start + n_episodes`` of the flattened grid (wrapping)."""
    return pp_sweep_rows(np.arange(n_episodes, dtype=np.int64) + start)


def pp_sweep_rows(rows):
    """Rows ``rows`` (any integer array, wrapping) of the flattened config-2 grid: parameter row i is
    (Lf[i // 16384], kp[(i // 512) % 32], ki[(i // 32) % 16], kd[i % 32])."""
    idx = np.asarray(rows, dtype=np.int64) % (64 * 32 * 16 * 32)
    n_episodes = len(idx)
    i_kd = idx % 32
    i_ki = (idx // 32) % 16
    i_kp = (idx // (32 * 16)) % 32
    i_lf = idx // (32 * 16 * 32)
    p = engine.default_params(n_episodes, dtype=np.float64)
    p[:, abi.P_LF] = np.linspace(2.0, 9.0, 64)[i_lf]
    p[:, abi.P_KP] = np.linspace(0.1, 1.0, 32)[i_kp]
    p[:, abi.P_KI] = np.linspace(0.0, 0.8, 16)[i_ki]
    p[:, abi.P_KD] = np.linspace(0.0, 0.6, 32)[i_kd]
    return p.astype(np.float32)


def stanley_lqg_params(n_episodes=1 << 22, seed=20261019, start=0):
    """Config 3: k ~ U(0.3, 2), k_soft ~ U(0.01, 1), mu ~ U(0.6, 1.2), c_f / c_r scale ~ U(0.8, 1.2).
    Row i depends only on (seed, i): shards of any size see identical rows."""
    p = engine.default_params(n_episodes, dtype=np.float64)
    u = _uniform_rows(seed, start, n_episodes, 5)
    p[:, abi.P_K_STANLEY] = 0.3 + 1.7 * u[:, 0]
    p[:, abi.P_K_SOFT] = 0.01 + 0.99 * u[:, 1]
    p[:, abi.P_MU] = 0.6 + 0.6 * u[:, 2]
    p[:, abi.P_C_F] = conf.c_f * (0.8 + 0.4 * u[:, 3])
    p[:, abi.P_C_R] = conf.c_r * (0.8 + 0.4 * u[:, 4])
    return p.astype(np.float32)


def stanley_lqg_rows(rows, seed=20261019):
    """Config 3's parameter rows at arbitrary absolute row numbers (a stratified sample of the 4 Mi sweep)."""
    rows = np.asarray(rows, dtype=np.int64)
    out = np.empty((len(rows), abi.N_PARAMS), dtype=np.float32)
    B = 1 << 16
    for b in np.unique(rows // B):
        sel = np.nonzero(rows // B == b)[0]
        blk = stanley_lqg_params(B, seed=seed, start=int(b) * B)
        out[sel] = blk[rows[sel] - int(b) * B]
    return out


def _uniform_rows(seed, start, n, cols):
    """U(0,1) draws addressed by absolute row: block b of 65 536 rows comes from
    ``default_rng([seed, b])`` so that any shard reproduces its rows without drawing the ones before."""
    B = 1 << 16
    out = np.empty((n, cols))
    pos = 0
    while pos < n:
        row = start + pos
        b, off = divmod(row, B)
        take = min(B - off, n - pos)
        blk = np.random.default_rng([seed, b]).random((B, cols))
        out[pos:pos + take] = blk[off:off + take]
        pos += take
    return out


def mpc_states(track, n_episodes=16384, seed=20261020, start=0):
    """Config 4: states sampled along the track: waypoint ~ U, lateral offset ~ N(0, 0.3 m), heading error ~
    N(0, 0.05 rad), v ~ U(2, 7).  Returns (state (E, 4) float64, ref_start (E,) int32)."""
    u = _uniform_rows(seed, start, n_episodes, 6)
    n = len(track.x)
    j = np.minimum((u[:, 0] * n).astype(np.int64), n - 1)
    jn = (j + 1) % n
    yaw = np.arctan2(track.y[jn] - track.y[j], track.x[jn] - track.x[j])
    g = np.sqrt(-2.0 * np.log(1.0 - u[:, 1]))                    # Box-Muller
    z0, z1 = g * np.cos(2 * np.pi * u[:, 2]), g * np.sin(2 * np.pi * u[:, 2])
    off = 0.3 * z0
    s = np.empty((n_episodes, 4))
    s[:, 0] = track.x[j] - np.sin(yaw) * off
    s[:, 1] = track.y[j] + np.cos(yaw) * off
    s[:, 2] = yaw + 0.05 * z1
    s[:, 3] = 2.0 + 5.0 * u[:, 3]
    return s, j.astype(np.int32)


def monte_carlo_params(n_episodes, seed=20261022, start=0):
    """Config 5: mild per-episode spread of look-ahead / Stanley gain around the defaults."""
    p = engine.default_params(n_episodes, dtype=np.float64)
    u = _uniform_rows(seed, start, n_episodes, 3)
    p[:, abi.P_LF] = 3.0 + 3.0 * u[:, 0]
    p[:, abi.P_K_STANLEY] = 0.5 + 1.5 * u[:, 1]
    p[:, abi.P_TARGET_SPEED] = conf.TARGET_SPEED * (0.6 + 0.4 * u[:, 2])
    return p.astype(np.float32)
