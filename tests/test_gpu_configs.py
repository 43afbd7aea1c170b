"""BASELINE.json configurations 3, 4 and 5 at FULL size on the GPU, checked through size-independent properties
(sharding == whole, determinism, aggregate == fixed-order sum of rows, argmin invariances) plus oracle spot checks
of sampled episodes, and hypothesis-driven single-step parity of the laws against the oracle restatement.
"""
import numpy as np
import pytest

import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi
from oracle import laws, loop

pytestmark = pytest.mark.gpu


def _dev(a, device):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def test_config3_stanley_lqg_dynamic_4M_episodes_sharded(cuda_device):
    """4 Mi episodes of Stanley + LQG + dynamic bicycle, 2 000 steps, as the 4 contiguous shards a 4-GPU job would
    run (here sequentially on one GPU): the concatenation must be bit-identical to sub-ranges run alone
    with the matching episode_id_base, aggregates must combine in rank order, and sampled on-track episodes must
    agree with the fp64 oracle."""
    import torch
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E, G, S = 1 << 22, 4, 2000
    per = E // G
    metrics, status, aggs = [], [], []
    for g in range(G):
        lo, hi = bgr.sharding.shard_range(E, g, G)
        assert (lo, hi) == (g * per, (g + 1) * per)
        cfg = bgr.RolloutConfig(steer="stanley", accel="lqg", model="dynamic", max_steps=S, episode_id_base=lo)
        params = bgr.sweeps.stanley_lqg_params(hi - lo, start=lo)
        init = bgr.initial_states(ta, hi - lo, v0=5.0)   # above the Euler stability limit of the tyre model
        res = bgr.rollout(track, cfg, _dev(params, cuda_device), _dev(init, cuda_device), aggregate=True)
        metrics.append(res.metrics.cpu().numpy())
        status.append(res.status.cpu().numpy())
        aggs.append(res.aggregate.cpu().numpy())
        del res
        torch.cuda.empty_cache()
    m, s = np.concatenate(metrics), np.concatenate(status)
    total = bgr.sharding.combine_aggregates(aggs)
    assert total[abi.A_EPISODES] == E and total[abi.A_STEPS] == s[:, abi.S_STEPS].astype(np.int64).sum()
    assert np.all((s[:, abi.S_FLAGS] == abi.ST_TIMEOUT) | (s[:, abi.S_FLAGS] & abi.ST_NONFINITE))
    finite = ~(s[:, abi.S_FLAGS] & abi.ST_NONFINITE).astype(bool)
    assert finite.mean() > 0.999
    assert np.all(np.isfinite(m[finite]))
    # about half of the sweep (k, k_soft, mu, stiffness spread at up to 8.5 m/s) stays inside the cone line -- the
    # oracle twin gives the same fraction (scripts/diag_parity.py); the rest is Stanley overshooting the R = 15 m bends
    on_track = m[:, abi.M_MAX_ABS_LAT] < 1.5
    assert 0.3 < on_track.mean() < 0.7, on_track.mean()
    # a window that straddles a shard boundary, run alone, reproduces the sharded rows bit for bit
    lo = per - 300
    cfg = bgr.RolloutConfig(steer="stanley", accel="lqg", model="dynamic", max_steps=S, episode_id_base=lo)
    again = bgr.rollout(track, cfg, bgr.sweeps.stanley_lqg_params(600, start=lo), bgr.initial_states(ta, 600, v0=5.0))
    assert np.array_equal(again.metrics, m[lo:lo + 600], equal_nan=True)
    assert np.array_equal(again.status, s[lo:lo + 600])
    # oracle spot check on sampled on-track episodes
    otrack = loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones)
    checked = 0
    for e in np.nonzero(on_track)[0][:: max(1, on_track.sum() // 4)][:4]:
        p = bgr.sweeps.stanley_lqg_params(1, start=int(e))[0].astype(np.float64)
        ref = loop.rollout_episode(otrack, loop.RolloutSpec(steer="stanley", accel="lqg", model="dynamic", max_steps=S),
                                   p, bgr.initial_states(ta, 1, v0=5.0)[0], record=False)
        assert s[e, abi.S_STEPS] == ref["steps"] and s[e, abi.S_LAPS] == ref["laps"]
        assert abs(m[e, abi.M_LAT_MEAN] - ref["metrics"]["lateral_mean"]) < 1e-3
        assert abs(m[e, abi.M_LAT_STD] - ref["metrics"]["lateral_std"]) < 1e-3
        assert abs(m[e, abi.M_HEAD_STD] - ref["metrics"]["heading_std"]) < 1e-3
        assert np.hypot(m[e, abi.M_X] - ref["state"][0], m[e, abi.M_Y] - ref["state"][1]) < 5e-3
        checked += 1
    assert checked >= 3


def test_config4_mpc_16k_episodes_4096_candidates_h30(cuda_device):
    """Full-size MPC candidate evaluation: argmin against the oracle on sampled episodes, and invariances that do
    not need the oracle: the best cost is unchanged by permuting the bank, and no candidate beats it."""
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta)
    E, K, H = 16384, 4096, 30
    states, ref_start = bgr.sweeps.mpc_states(ta, E)
    bank = bgr.candidate_bank(K, H)
    cfg = bgr.MpcConfig(horizon=H, dt=0.05)
    d_state, d_ref, d_bank = (_dev(a, cuda_device) for a in (states, ref_start, bank))
    res = bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
    idx = res.best_idx.cpu().numpy()
    cost = res.best_cost.cpu().numpy()
    u = res.best_u.cpu().numpy()
    assert np.all(idx >= 0) and np.all(np.isfinite(cost)) and np.all(cost >= 0)
    assert np.all(np.abs(u[:, 1]) <= bgr.conf.MAX_STEER + 1e-7)
    assert np.all((u[:, 0] <= bgr.conf.MAX_ACCEL) & (u[:, 0] >= bgr.conf.MAX_DECEL))
    # permuting the candidates permutes the argmin, not the optimum
    perm = np.random.default_rng(11).permutation(K)
    res_p = bgr.mpc_evaluate(track, cfg, d_state, d_ref, _dev(bank[perm], cuda_device))
    assert np.array_equal(res_p.best_cost.cpu().numpy(), cost)
    idx_p = res_p.best_idx.cpu().numpy()
    same_candidate = perm[idx_p] == idx
    assert same_candidate.mean() > 0.999          # the rest are exact fp32 cost ties between two candidates
    assert np.array_equal(res_p.best_u.cpu().numpy()[same_candidate], u[same_candidate])
    # oracle argmin on sampled episodes
    b64 = bank.astype(np.float64)
    for e in range(0, E, E // 24):
        r = (ref_start[e] + np.arange(H)) % ta.n
        c, f = laws.mpc_costs_batch(*states[e], ta.x[r], ta.y[r], b64, 0.05)
        c = np.where(f, c, np.inf)
        best = int(np.argmin(c))
        assert np.isclose(cost[e], c[best], rtol=2e-5)
        if np.partition(c, 1)[1] > c[best] * (1 + 1e-4):
            assert idx[e] == best


def test_config5_skidpad_mixed_controller_monte_carlo(cuda_device):
    """Skidpad figure-eight Monte Carlo with Philox pose noise and perturbed cones, episodes grouped by controller
    kind on the host (homogeneous launches): runs are reproducible bit for bit, the noise stream is keyed by
    (seed, GLOBAL episode id) -- so sharded == whole -- and a different seed gives a different draw."""
    ta = bgr.tracks.skidpad(2048)
    track = bgr.Track.from_arrays(ta, cone_reach=4.0)
    E = 3 * 8192
    kinds = [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]
    per = E // len(kinds)
    noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
    init = bgr.initial_states(ta, per, index=0)
    init[:, abi.YAW] = 0.0
    out = {}
    for g, (steer, accel, model) in enumerate(kinds):
        lo = g * per
        params = bgr.sweeps.monte_carlo_params(per, start=lo)
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=1500, laps=1, episode_id_base=lo,
                                **noise)
        a = bgr.rollout(track, cfg, _dev(params, cuda_device), _dev(init, cuda_device), aggregate=True)
        b = bgr.rollout(track, cfg, _dev(params, cuda_device), _dev(init, cuda_device), aggregate=True)
        am, as_ = a.metrics.cpu().numpy(), a.status.cpu().numpy()
        assert np.array_equal(am, b.metrics.cpu().numpy(), equal_nan=True)
        assert np.array_equal(as_, b.status.cpu().numpy())
        assert np.array_equal(a.aggregate.cpu().numpy(), b.aggregate.cpu().numpy())
        # two half shards with their own episode_id_base == the whole
        h = per // 2
        c1 = bgr.RolloutConfig(**{**cfg.__dict__, "episode_id_base": lo + h})
        half = bgr.rollout(track, c1, params[h:], init[h:])
        assert np.array_equal(half.metrics, am[h:], equal_nan=True) and np.array_equal(half.status, as_[h:])
        # another seed: another draw
        c2 = bgr.RolloutConfig(**{**cfg.__dict__, "noise_seed": 7})
        other = bgr.rollout(track, c2, params[:256], init[:256])
        assert not np.array_equal(other.metrics, am[:256])
        assert np.all(as_[:, abi.S_CONE_HITS] >= 0) and np.all(as_[:, abi.S_CONE_HITS] <= len(ta.cones))
        assert a.aggregate.cpu().numpy()[abi.A_CONE_HITS] == as_[:, abi.S_CONE_HITS].sum()
        # (exact global scans are the last resort: the first search of an episode descends from waypoint 0, and the
        #  crossing is normally answered by the twin triples / rival lists)
        assert as_[:, abi.S_FALLBACKS].min() >= 0
        out[steer + accel] = as_
    # pure pursuit follows the unrolled figure eight to its end in most episodes
    pp = out["pure_pursuitpid"]
    assert (pp[:, abi.S_FLAGS] & abi.ST_PATH_END).astype(bool).mean() > 0.8


def test_single_step_laws_against_the_oracle_hypothesis(cuda_device):
    """Property test (hypothesis): for random states around a random ellipse, one ``compute_steering`` of each
    drop-in law (fp64 audit kernel through bgr_step_host) equals the oracle's restatement of the reference law:
    indices exactly, steering to fp64 round-off."""
    from hypothesis import given, settings, strategies as st, HealthCheck
    from conftest import Obj, ellipse_path

    cx, cy = ellipse_path(360, 35.0, 18.0)
    path = np.column_stack([np.arange(len(cx)), cx, cy])
    pp = bgr.PurePursuitController()
    stan = bgr.StanleyController()

    @settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(th=st.floats(0, 2 * np.pi), off=st.floats(-2.5, 2.5), dyaw=st.floats(-1.0, 1.0), v=st.floats(0.0, 12.0),
           lf=st.floats(1.5, 12.0), start=st.integers(0, 359), k=st.floats(0.1, 3.0), ks=st.floats(0.01, 2.0))
    def check(th, off, dyaw, v, lf, start, k, ks):
        x = (35.0 + off) * np.cos(th)
        y = (18.0 + off) * np.sin(th)
        yaw = np.arctan2(18.0 * np.cos(th), -35.0 * np.sin(th)) + dyaw
        s = Obj(x=x, y=y, yaw=yaw, v=v)
        pp.look_ahead_distance = lf
        got, ind = pp.compute_steering(s, path, start)
        want, wind = laws.pure_pursuit_steering(x, y, yaw, cx, cy, start, lf)
        # exact-distance near-ties of the look-ahead test are the one documented exemption
        d = np.hypot(cx - x, cy - y)
        tie = np.min(np.abs(d[start:wind + 2] - lf)) < 1e-9 if wind + 2 > start else False
        if not tie:
            assert ind == wind
            assert abs(got - want) < 1e-12
        stan.k, stan.k_soft = k, ks
        got, ind = stan.compute_steering(s, path, 0)
        want, wind = laws.stanley_steering(x, y, yaw, v, cx, cy, k, ks)
        part = np.partition(d, 1)
        if part[1] - part[0] > 1e-9:
            assert int(ind) == wind
            assert abs(got - want) < 1e-11

    check()
