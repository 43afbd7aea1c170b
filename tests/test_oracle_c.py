"""The plain-C oracle twin (oracle/c/bgr_oracle.c) against the numpy oracle (oracle/loop.py), which is itself pinned
bit-exact on the reference-generated golden vectors.  Same algorithm and operation order in fp64; the only slack is
the last-place difference between numpy's and glibc's sin / cos / hypot / atan2 / exp kernels."""
import numpy as np
import pytest

from oracle import cport, loop


def _oval():
    import bgr_pathplanning_control_b200 as bgr
    ta = bgr.tracks.stadium_oval(512)
    return ta, loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones)


def _skidpad():
    import bgr_pathplanning_control_b200 as bgr
    ta = bgr.tracks.skidpad(600)
    return ta, loop.Track(ta.x, ta.y, ta.kappa, False, ta.cones)


CASES = [("pure_pursuit", "pid", "kinematic", 3, 0), ("pure_pursuit", "lqg", "dynamic", 2, 0),
         ("stanley", "pid", "dynamic", 1, 1), ("stanley", "lqg", "kinematic", 1, 1)]


@pytest.mark.parametrize("steer,accel,model,laps,laps_target", CASES)
@pytest.mark.parametrize("trackname", ["oval", "skidpad"])
def test_c_twin_equals_numpy_oracle(steer, accel, model, laps, laps_target, trackname):
    import bgr_pathplanning_control_b200 as bgr
    ta, tr = _oval() if trackname == "oval" else _skidpad()
    if trackname == "skidpad":
        laps, laps_target = 1, 0
    E, S = 6, 500
    rng = np.random.default_rng(5)
    params = np.tile(loop.default_params(), (E, 1))
    params[:, loop.P_LF] = rng.uniform(3.0, 7.0, E)
    params[:, loop.P_K_STANLEY] = rng.uniform(0.5, 1.5, E)
    params[:, loop.P_MU] = rng.uniform(0.8, 1.2, E)
    init = bgr.initial_states(ta, E, lateral_offset=rng.uniform(-0.4, 0.4, E), v0=2.0)
    if trackname == "skidpad":
        init[:, 2] = 0.0
    noise = loop.NoiseSpec(seed=99, sigma_xy=0.03, sigma_yaw=0.005, sigma_v=0.05, sigma_cone=0.04)
    spec = loop.RolloutSpec(steer=steer, accel=accel, model=model, max_steps=S, laps=laps, laps_target=laps_target,
                            hit_radius=0.6, noise=noise)
    m, s, t = cport.rollout(tr, spec, params, init, episode_id_base=40, traj=True, threads=2)
    for e in range(E):
        ref = loop.rollout_episode(tr, spec, params[e], init[e], episode_id=40 + e)
        n = ref["steps"]
        assert s[e, 1] == n and s[e, 0] == ref["status"]
        assert s[e, 2] == ref["target_ind"] and s[e, 3] == ref["nearest_ind"]
        assert s[e, 4] == ref["cone_hits"] and s[e, 5] == ref["laps"]
        assert np.array_equal(t[e, :n, 8].astype(int), ref["ti"]) and np.array_equal(t[e, :n, 9].astype(int), ref["near"])
        np.testing.assert_allclose(t[e, :n, :6], ref["traj"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(t[e, :n, 6], ref["steer"], rtol=0, atol=1e-10)
        np.testing.assert_allclose(t[e, :n, 7], ref["accel"], rtol=0, atol=1e-9)
        rm = ref["metrics"]
        np.testing.assert_allclose(m[e, :5], [rm["lateral_mean"], rm["lateral_std"], rm["heading_mean"],
                                              rm["heading_std"], rm["lap_time"]], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(m[e, 6:12], ref["state"], rtol=0, atol=1e-8)


@pytest.mark.parametrize("trackname,horizon,stride", [("oval", 40, 3), ("oval", 12, 8), ("skidpad", 30, 2)])
def test_c_twin_replanning_regime(trackname, horizon, stride):
    """Re-planning replay (oracle/loop.py step 2'): same local-path windows, planner index and controller index."""
    import bgr_pathplanning_control_b200 as bgr
    ta, tr = _oval() if trackname == "oval" else _skidpad()
    E, S = 6, 500
    rng = np.random.default_rng(6)
    params = np.tile(loop.default_params(), (E, 1))
    params[:, loop.P_LF] = rng.uniform(2.5, 7.0, E)
    init = bgr.initial_states(ta, E, lateral_offset=rng.uniform(-0.4, 0.4, E), v0=2.0)
    if trackname == "skidpad":
        init[:, 2] = 0.0
    noise = loop.NoiseSpec(seed=17, sigma_xy=0.03, sigma_yaw=0.005, sigma_v=0.05)
    spec = loop.RolloutSpec(steer="pure_pursuit", accel="pid", model="kinematic", max_steps=S, laps_target=1,
                            noise=noise, replan_horizon=horizon, replan_stride=stride, replan_after=3)
    m, s, t = cport.rollout(tr, spec, params, init, episode_id_base=7, traj=True, threads=2)
    for e in range(E):
        ref = loop.rollout_episode(tr, spec, params[e], init[e], episode_id=7 + e)
        n = ref["steps"]
        assert s[e, 1] == n and s[e, 0] == ref["status"] and s[e, 2] == ref["target_ind"] and s[e, 5] == ref["laps"]
        assert np.array_equal(t[e, :n, 8].astype(int), ref["ti"]) and np.array_equal(t[e, :n, 9].astype(int), ref["near"])
        np.testing.assert_allclose(t[e, :n, :6], ref["traj"], rtol=0, atol=1e-9)
        np.testing.assert_allclose(t[e, :n, 6], ref["steer"], rtol=0, atol=1e-10)


def test_c_twin_edge_cases():
    import bgr_pathplanning_control_b200 as bgr
    ta, tr = _oval()
    spec = loop.RolloutSpec(max_steps=1)
    m, s, _ = cport.rollout(tr, spec, loop.default_params()[None], bgr.initial_states(ta, 1))
    assert s[0, 1] == 1 and np.isnan(m[0, 1]) and m[0, 4] == 0.0           # one row: pandas std is NaN
    m, s, _ = cport.rollout(tr, loop.RolloutSpec(max_steps=10), np.zeros((0, 16)), np.zeros((0, 6)))
    assert m.shape == (0, 12)
