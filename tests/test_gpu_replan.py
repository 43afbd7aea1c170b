"""Receding-horizon re-planning replay: the regime the live stack drives in --
``Planner.update`` (nodes/planner.py:17-102) re-plans a short local path whenever its own look-ahead index has passed 3
and hands path + index to ``Controller.update`` (nodes/controller.py:56-62).

Pinned twice: (1) against tests/golden/replan_kat.json, frozen from the reference's UNMODIFIED Planner and Controller
nodes (oracle/gen_golden_replan.py) -- the GPU runs the same episodes through ``bgr_rollout``; (2) against the oracle's
restated loop (itself bit-exact on those vectors, tests/test_oracle_golden.py) on sweeps, with noise, on the oval.
"""
import json
import os

import numpy as np
import pytest

import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from conftest import GOLDEN_DIR
from oracle import loop
from test_gpu_rollout import compare

pytestmark = pytest.mark.gpu


def unhex(v):
    return [unhex(e) for e in v] if isinstance(v, list) else float.fromhex(v)


def ellipse_track(closed):
    th = np.linspace(0, 2 * np.pi, 720, endpoint=False)
    cx, cy = 40.0 * np.cos(th), 20.0 * np.sin(th)
    curve = (40.0 * 20.0) / ((40.0 * np.sin(th)) ** 2 + (20.0 * np.cos(th)) ** 2) ** 1.5
    if not closed:
        cx, cy, curve = cx[:200], cy[:200], curve[:200]
    return cx, cy, curve


@pytest.mark.parametrize("name", ["ellipse_h40_s3", "ellipse_h12_s8_lf3", "open_quarter_h30_s2"])
@pytest.mark.parametrize("precision", ["fp64", "fp32"])
def test_reference_nodes_golden_run(cuda_device, name, precision):
    with open(os.path.join(GOLDEN_DIR, "replan_kat.json")) as f:
        g = json.load(f)[name]
    cx, cy, curve = ellipse_track(g["closed"])
    track = bgr.Track(cx, cy, curve, closed=g["closed"])
    cfg = bgr.RolloutConfig(steer="pure_pursuit", accel="pid", model="dynamic", precision=precision,
                            max_steps=g["steps"], replan_horizon=g["horizon"], replan_stride=g["stride"])
    params = bgr.default_params(1, dtype=np.float64 if precision == "fp64" else np.float32)
    params[:, abi.P_LF] = unhex(g["lf"])
    init = np.array([unhex(g["init"])])
    res = bgr.rollout(track, cfg, params.astype(np.float32), init, traj=True)
    n = int(res.status[0, abi.S_STEPS])
    want = np.asarray(unhex(g["traj"]))
    if g["closed"]:
        assert n == g["steps"] and res.status[0, abi.S_FLAGS] == abi.ST_TIMEOUT
    else:   # the open path ends when the target is its last waypoint (the reference run was simply cut later)
        assert res.status[0, abi.S_FLAGS] == abi.ST_PATH_END and 100 <= n < g["steps"]
    tr = res.traj[0, :n]
    got_ti = tr[:, abi.T_TARGET_IND].astype(int)
    if precision == "fp64":
        assert list(got_ti) == g["ti"][:n]                                   # the controller's LOCAL index, exact
    else:
        # fp32 laws: an index may differ only at an exact-distance near-tie of the ORACLE's own run (margin < 2e-5 m);
        # beyond such a step the two runs are different episodes, so the comparison stops there
        bad = np.nonzero(got_ti != np.asarray(g["ti"][:n]))[0]
        if len(bad):
            p64 = loop.default_params()
            p64[loop.P_LF] = unhex(g["lf"])
            spec = loop.RolloutSpec(steer="pure_pursuit", accel="pid", model="dynamic", max_steps=g["steps"],
                                    replan_horizon=g["horizon"], replan_stride=g["stride"], replan_after=3)
            ref = loop.rollout_episode(loop.Track(cx, cy, curve, closed=g["closed"]), spec, p64, unhex(g["init"]))
            k = int(bad[0])
            assert min(ref["pp_margin"][k], ref["near_margin"][k]) < 2e-5, (k, ref["pp_margin"][k])
            n = k
            tr = tr[:n]
            assert n >= 50
    # (the batched API takes fp32 parameter rows: kp = 0.35 etc. are rounded, which the audit build then integrates)
    tol = (2e-5, 2e-6) if precision == "fp64" else (1e-3, 1e-4)
    assert np.hypot(tr[:, abi.T_X] - want[:n, 0], tr[:, abi.T_Y] - want[:n, 1]).max() < tol[0]
    assert np.abs(tr[:, abi.T_YAW] - want[:n, 2]).max() < tol[1]
    assert np.abs(tr[:, abi.T_STEER] - np.asarray(unhex(g["steer"]))[:n]).max() < (1e-6 if precision == "fp64" else 5e-4)
    assert np.abs(tr[:, abi.T_ACCEL] - np.asarray(unhex(g["accel"]))[:n]).max() < (1e-5 if precision == "fp64" else 2e-3)


@pytest.mark.parametrize("precision", ["fp32", "fp64"])
@pytest.mark.parametrize("horizon,stride", [(40, 3), (12, 8), (64, 1)])
def test_sweep_against_the_oracle_loop(cuda_device, precision, horizon, stride):
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E = 96
    rng = np.random.default_rng(77)
    params = bgr.default_params(E)
    params[:, abi.P_LF] = rng.uniform(2.5, 8.0, E)
    params[:, abi.P_KP] = rng.uniform(0.2, 1.0, E)
    params[:, abi.P_TARGET_SPEED] = rng.uniform(4.0, 9.0, E)
    init = bgr.initial_states(ta, E, index=0, lateral_offset=rng.uniform(-0.4, 0.4, E),
                              yaw_offset=rng.uniform(-0.05, 0.05, E), v0=3.0)
    for e in range(E):       # start anywhere on the lap, also right before the seam
        j = int(rng.integers(0, ta.n)) if e % 4 else ta.n - 3
        init[e] = bgr.initial_states(ta, 1, index=j, lateral_offset=np.array([rng.uniform(-0.4, 0.4)]),
                                     yaw_offset=np.array([rng.uniform(-0.05, 0.05)]), v0=3.0)[0]
    cfg = bgr.RolloutConfig(steer="pure_pursuit", accel="pid", model="kinematic", precision=precision, max_steps=700,
                            replan_horizon=horizon, replan_stride=stride, laps_target=0)
    tol = dict(pos_tol=1e-8, yaw_tol=1e-9, e_ct_tol=1e-8) if precision == "fp64" else {}
    res, worst = compare(ta, track, cfg, params, init, list(range(0, E, 5)), **tol)
    # in this mode nobody runs out of path on a closed track; the local index never exceeds the horizon
    assert np.all(res.status[:, abi.S_FLAGS] == abi.ST_TIMEOUT)
    assert res.status[:, abi.S_TARGET_IND].max() <= horizon - 1
    assert res.status[:, abi.S_LAPS].max() >= 1


def test_with_observation_noise_and_lap_stop(cuda_device):
    """Config-5 style noise: the planner plans from the OBSERVED position (nodes/planner.py:19,38) and both scans
    measure distances from it; episodes stop after one completed lap."""
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E = 64
    params = bgr.default_params(E)
    params[:, abi.P_LF] = np.linspace(3.0, 7.0, E)
    init = bgr.initial_states(ta, E, v0=4.0)
    noise = loop.NoiseSpec(seed=4242, sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1)
    cfg = bgr.RolloutConfig(steer="pure_pursuit", accel="pid", model="kinematic", max_steps=900, laps_target=1,
                            replan_horizon=40, replan_stride=3, sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1,
                            noise_seed=4242)
    res, worst = compare(ta, track, cfg, params, init, list(range(0, E, 7)), noise=noise)
    done = res.status[:, abi.S_FLAGS] == abi.ST_LAPS_DONE
    assert done.mean() > 0.9 and np.all(res.status[done, abi.S_LAPS] == 1)


def test_replay_is_rejected_where_it_is_undefined(cuda_device):
    ta = bgr.tracks.stadium_oval(256)
    track = bgr.Track.from_arrays(ta)
    with pytest.raises(abi.BgrError):
        bgr.rollout(track, bgr.RolloutConfig(steer="stanley", replan_horizon=40), bgr.default_params(1),
                    bgr.initial_states(ta, 1))
    with pytest.raises(abi.BgrError):
        bgr.rollout(track, bgr.RolloutConfig(replan_horizon=1), bgr.default_params(1), bgr.initial_states(ta, 1))


def test_replay_at_scale_against_the_c_twin(cuda_device):
    """512 episodes x 2 000 steps of config 2's sweep rows in the re-planning regime: fp32 production kernel (the
    replay runs in the audited full build) vs the plain-C fp64 twin.  As in test_parity_at_scale_against_the_c_twin:
    every index mismatch must be covered by the kernel's near-tie audit channel; well-conditioned unflagged episodes
    end within the north-star tolerances."""
    from oracle import cport
    from test_gpu_rollout import OFF_TRACK_M, POS_TOL, YAW_TOL, LAP_TOL, otrack
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E, S = 512, 2000
    params = bgr.sweeps.pp_sweep_params(E, start=32 * 16384 + 4242)
    init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3, 0.3, E))
    cfg = bgr.RolloutConfig(steer="pure_pursuit", accel="pid", model="kinematic", max_steps=S, replan_horizon=40,
                            replan_stride=3)
    res = bgr.rollout(track, cfg, params, init)
    spec = loop.RolloutSpec(steer="pure_pursuit", accel="pid", model="kinematic", max_steps=S, replan_horizon=40,
                            replan_stride=3, replan_after=3)
    p64 = params.astype(np.float64)
    m, s, _ = cport.rollout(otrack(ta), spec, p64, init)
    nudged = init.copy()
    nudged[:, abi.X] += 1e-7
    m2, _, _ = cport.rollout(otrack(ta), spec, p64, nudged)
    well = np.hypot(m2[:, 6] - m[:, 6], m2[:, 7] - m[:, 7]) < 1e-5
    on_track = (m[:, 5] < OFF_TRACK_M) & (res.metrics[:, abi.M_MAX_ABS_LAT] < OFF_TRACK_M)
    unflagged = res.status[:, abi.S_NEAR_TIES] == 0
    good = on_track & well & unflagged
    assert on_track.mean() > 0.4 and good.sum() > E // 8, (on_track.mean(), well.mean(), unflagged.mean(), good.sum())
    cols = [abi.S_FLAGS, abi.S_STEPS, abi.S_TARGET_IND, abi.S_NEAREST_IND, abi.S_LAPS]
    mism = np.any(res.status[:, cols] != s[:, cols], axis=1)
    assert not np.any(mism & good), f"{(mism & good).sum()} index mismatches without a flagged near-tie"
    dpos = np.hypot(res.metrics[:, abi.M_X] - m[:, 6], res.metrics[:, abi.M_Y] - m[:, 7])
    dyaw = np.abs(res.metrics[:, abi.M_YAW] - m[:, 8])
    assert dpos[good].max() < POS_TOL and dyaw[good].max() < YAW_TOL, (dpos[good].max(), dyaw[good].max())
    assert np.abs(res.metrics[good, abi.M_LAP_TIME] - m[good, 4]).max() < LAP_TOL
    assert np.abs(res.metrics[good, abi.M_LAT_MEAN] - m[good, 0]).max() < 1e-4
    assert res.status[:, abi.S_LAPS].max() >= 2


def test_replay_through_host_buffers_and_two_phase_launches(cuda_device):
    """The replay is plain episode state: the host-buffer entry point (chunked, pipelined) and a two-phase launch
    (cfg.split_steps) return the rows of the single device launch bit for bit."""
    import torch
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E = 3000
    params = bgr.sweeps.pp_sweep_params(E, start=32 * 16384)
    init = bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.3, 0.3, E))
    kw = dict(steer="pure_pursuit", accel="pid", model="kinematic", max_steps=500, laps_target=1, replan_horizon=25,
              replan_stride=4)
    dev = bgr.rollout(track, bgr.RolloutConfig(**kw), torch.from_numpy(params).to(cuda_device),
                      torch.from_numpy(init).to(cuda_device), aggregate=True)
    host = bgr.rollout(track, bgr.RolloutConfig(**kw), params, init, aggregate=True)
    split = bgr.rollout(track, bgr.RolloutConfig(split_steps=120, **kw), params, init, aggregate=True)
    for other in (host, split):
        assert np.array_equal(dev.status.cpu().numpy(), other.status)
        assert np.array_equal(dev.metrics.cpu().numpy().view(np.int32), other.metrics.view(np.int32))
        assert np.array_equal(dev.aggregate.cpu().numpy(), other.aggregate)
