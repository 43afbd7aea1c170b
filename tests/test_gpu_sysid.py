"""Batched system identification: ``bgr_identify_dynamics`` against the reference's own call,
``sklearn.linear_model.LinearRegression().fit(np.hstack([X, U]), Y)`` (old/test_runs_system_id.py:84-111, restated
in oracle/laws.py), on synthetic linear systems with known matrices and on logged dynamic-bicycle episodes."""
import numpy as np
import pytest

import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from oracle import laws

pytestmark = pytest.mark.gpu


def test_recovers_known_linear_systems(cuda_device):
    import torch
    rng = np.random.default_rng(12)
    E, S = 40, 300
    traj = np.zeros((E, S, abi.N_TRAJ))
    truth = []
    rows = rng.integers(60, S + 1, E).astype(np.int32)
    for e in range(E):
        A = np.eye(6) * 0.95 + rng.normal(0, 0.03, (6, 6))
        B = rng.normal(0, 0.2, (6, 2))
        c = rng.normal(0, 0.1, 6)
        x = rng.normal(0, 1, 6) * [30, 30, 1, 5, 0.3, 0.05]           # badly scaled columns, like real logs
        for k in range(S):
            u = np.array([np.sin(0.07 * k + e), 0.3 * np.cos(0.11 * k)]) + rng.normal(0, 0.05, 2)
            traj[e, k, :6] = x
            traj[e, k, abi.T_ACCEL], traj[e, k, abi.T_STEER] = u
            x = A @ x + B @ u + c + rng.normal(0, 1e-3, 6)
        truth.append((A, B, c))
    for dev_arrays in (False, True):
        if dev_arrays:
            A_, B_, c_, ok = bgr.identify_dynamics(torch.from_numpy(traj).to(cuda_device), torch.from_numpy(rows).to(cuda_device))
            A_, B_, c_, ok = (t.cpu().numpy() for t in (A_, B_, c_, ok))
        else:
            A_, B_, c_, ok = bgr.identify_dynamics(traj, rows)
        assert ok.all()
        for e in range(E):
            T = rows[e]
            inputs = traj[e, :T][:, [abi.T_ACCEL, abi.T_STEER]]
            wA, wB, wc, C, D = laws.identify_dynamics(traj[e, :T, :6], inputs)
            # Cholesky on column-scaled normal equations vs sklearn's SVD-based lstsq: 1e-6 on O(1) coefficients
            np.testing.assert_allclose(A_[e], wA, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(B_[e], wB, rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(c_[e], wc, rtol=1e-5, atol=1e-5)
            assert np.abs(A_[e] - truth[e][0]).max() < 0.05 and np.abs(B_[e] - truth[e][1]).max() < 0.05
            assert C.shape == (6, 6) and not D.any()


def test_on_logged_episodes_and_rank_deficiency(cuda_device):
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E = 6
    params = bgr.default_params(E)
    params[:, abi.P_LF] = np.linspace(4.0, 7.0, E)
    init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3, 0.3, E), v0=5.0)
    res = bgr.rollout(track, bgr.RolloutConfig(steer="pure_pursuit", accel="pid", model="dynamic", max_steps=500, laps=2),
                      params, init, traj=True)
    A, B, c, ok = bgr.identify_dynamics(res)
    assert ok.all()
    for e in range(E):
        n = res.status[e, abi.S_STEPS]
        wA, wB, wc, _, _ = laws.identify_dynamics(res.traj[e, :n, :6], res.traj[e, :n][:, [abi.T_ACCEL, abi.T_STEER]])
        # real logs are ill conditioned (x, y ~ 30 m next to beta ~ 1e-2): compare the one-step PREDICTIONS
        Z = np.hstack([res.traj[e, :n - 1, :6], res.traj[e, :n - 1][:, [abi.T_ACCEL, abi.T_STEER]]])
        pred = Z @ np.hstack([A[e], B[e]]).T + c[e]
        want = Z @ np.hstack([wA, wB]).T + wc
        assert np.abs(pred - want).max() < 1e-6
        assert np.abs(pred - res.traj[e, 1:n, :6]).max() < 1.0             # a linear fit of a nonlinear log: coarse only
    # the kinematic model never moves yaw_rate / beta: constant columns -> flagged, NaN (LinearRegression would
    # return the minimum-norm solution)
    res = bgr.rollout(track, bgr.RolloutConfig(max_steps=200, laps=2), params, bgr.initial_states(ta, E), traj=True)
    A, B, c, ok = bgr.identify_dynamics(res)
    assert not ok.any() and np.isnan(A).all()
    # too few rows
    A, B, c, ok = bgr.identify_dynamics(np.zeros((2, 5, abi.N_TRAJ)))
    assert not ok.any()
