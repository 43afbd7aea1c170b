"""The C ABI driven from plain C (examples/c_abi_demo.c): it compiles against include/bgr_rollout.h with gcc -std=c99,
links only libbgr_rollout.so, fails loudly without a CUDA device, and on a B200 prints the same lap metrics as the
Python host API for the same inputs."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "bgr_pathplanning_control_b200")


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("demo") / "c_abi_demo")
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L" + LIBDIR, "-lbgr_rollout", "-lm",
                    "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def test_demo_compiles_links_and_fails_loudly_without_a_gpu(demo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([demo, "16", "10"], capture_output=True, text=True)
    assert out.returncode == 1 and "bgr_track_create" in out.stderr and "(-2)" in out.stderr


@pytest.mark.gpu
def test_demo_matches_the_python_host_api(demo):
    import bgr_pathplanning_control_b200 as bgr
    from bgr_pathplanning_control_b200 import _abi as abi
    E, S = 4096, 800
    out = subprocess.run([demo, str(E), str(S)], capture_output=True, text=True, check=True).stdout
    ta = bgr.tracks.ellipse(720, 40.0, 20.0)
    track = bgr.Track.from_arrays(ta)
    params = bgr.default_params(E)
    params[:, abi.P_LF] = (3.0 + 5.0 * np.arange(E, dtype=np.float32) / np.float32(E - 1)).astype(np.float32)
    init = np.zeros((E, 6))
    init[:, 0], init[:, 2] = 40.0, np.pi / 2
    res = bgr.rollout(track, bgr.RolloutConfig(max_steps=S, laps=4), params, init, aggregate=True)
    agg = re.search(r"aggregate sum_lat_mean (\S+) sum_lat_std (\S+) sum_lap_time (\S+) finished (\S+)", out)
    got = [float(g) for g in agg.groups()]
    want = [res.aggregate[abi.A_SUM_LAT_MEAN], res.aggregate[abi.A_SUM_LAT_STD], res.aggregate[abi.A_SUM_LAP_TIME],
            res.aggregate[abi.A_FINISHED]]
    # the C program computes the ellipse with libm, numpy with its own kernels: last-place differences in waypoints
    np.testing.assert_allclose(got, want, rtol=1e-6)
    for e, ln in enumerate(re.findall(r"^episode \d+ .*$", out, re.M)):
        vals = dict(zip(ln.split()[2::2], ln.split()[3::2]))
        assert abs(float(vals["lat_mean"]) - res.metrics[e, abi.M_LAT_MEAN]) < 1e-5
        assert abs(float(vals["lap_time"]) - res.metrics[e, abi.M_LAP_TIME]) < 1e-6
        assert int(vals["laps"]) == res.status[e, abi.S_LAPS]
    assert f"episodes {E} steps {float(res.aggregate[abi.A_STEPS]):.0f}" in out
