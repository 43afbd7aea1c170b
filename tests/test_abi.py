"""The C-ABI boundary without a GPU: the library is built, loads, exports every symbol the header
declares, the Python mirror of the constants / struct layouts agrees with the header (checked by
compiling a probe with gcc), argument errors are reported through status codes, and no compute entry
point silently succeeds without a CUDA device."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import bgr_pathplanning_control_b200 as bgr
from bgr_pathplanning_control_b200 import _abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_is_built_and_exports_the_header():
    lib = abi.load()
    declared = abi.header_functions()
    assert sorted(declared) == sorted(abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bgr_abi_version() == abi.BGR_ABI_VERSION
    # only the ABI is exported (built with -fvisibility=hidden)
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(declared) <= exported
    assert not [s for s in exported if not s.startswith("bgr_") and not s.startswith("_")], exported


def test_constants_match_header():
    h = abi.header_constants()
    names = {
        "BGR_ABI_VERSION": abi.BGR_ABI_VERSION, "BGR_OK": abi.BGR_OK, "BGR_ERR_BAD_ARG": abi.BGR_ERR_BAD_ARG,
        "BGR_ERR_CUDA": abi.BGR_ERR_CUDA, "BGR_ERR_TOO_LARGE": abi.BGR_ERR_TOO_LARGE,
        "BGR_STEER_PURE_PURSUIT": abi.STEER_PURE_PURSUIT, "BGR_STEER_STANLEY": abi.STEER_STANLEY,
        "BGR_ACCEL_PID": abi.ACCEL_PID, "BGR_ACCEL_LQG": abi.ACCEL_LQG,
        "BGR_MODEL_KINEMATIC": abi.MODEL_KINEMATIC, "BGR_MODEL_DYNAMIC": abi.MODEL_DYNAMIC,
        "BGR_PRECISION_FP32": abi.PRECISION_FP32, "BGR_PRECISION_FP64": abi.PRECISION_FP64,
        "BGR_N_PARAMS": abi.N_PARAMS, "BGR_N_STATE": abi.N_STATE, "BGR_N_METRICS": abi.N_METRICS,
        "BGR_N_STATUS": abi.N_STATUS, "BGR_N_TRAJ": abi.N_TRAJ, "BGR_N_AGG": abi.N_AGG, "BGR_N_CMD": abi.N_CMD,
        "BGR_N_CTRL": abi.N_CTRL, "BGR_N_OUT": abi.N_OUT, "BGR_MAX_CONES": abi.MAX_CONES,
        "BGR_CONE_GUARD": abi.CONE_GUARD, "BGR_CFG_AUDIT": abi.CFG_AUDIT,
        "BGR_CFG_STANLEY_SHADOWED": abi.CFG_STANLEY_SHADOWED, "BGR_C_PREV_STEER": abi.C_PREV_STEER,
        "BGR_P_LF": abi.P_LF, "BGR_P_KP": abi.P_KP, "BGR_P_KI": abi.P_KI, "BGR_P_KD": abi.P_KD,
        "BGR_P_TARGET_SPEED": abi.P_TARGET_SPEED, "BGR_P_K_EXPO": abi.P_K_EXPO, "BGR_P_K_STANLEY": abi.P_K_STANLEY,
        "BGR_P_K_SOFT": abi.P_K_SOFT, "BGR_P_MU": abi.P_MU, "BGR_P_C_F": abi.P_C_F, "BGR_P_C_R": abi.P_C_R,
        "BGR_P_MASS": abi.P_MASS, "BGR_P_I_Z": abi.P_I_Z,
        "BGR_M_LAT_MEAN": abi.M_LAT_MEAN, "BGR_M_LAT_STD": abi.M_LAT_STD, "BGR_M_HEAD_MEAN": abi.M_HEAD_MEAN,
        "BGR_M_HEAD_STD": abi.M_HEAD_STD, "BGR_M_LAP_TIME": abi.M_LAP_TIME, "BGR_M_MAX_ABS_LAT": abi.M_MAX_ABS_LAT,
        "BGR_M_X": abi.M_X, "BGR_M_BETA": abi.M_BETA,
        "BGR_S_FLAGS": abi.S_FLAGS, "BGR_S_STEPS": abi.S_STEPS, "BGR_S_TARGET_IND": abi.S_TARGET_IND,
        "BGR_S_NEAREST_IND": abi.S_NEAREST_IND, "BGR_S_CONE_HITS": abi.S_CONE_HITS, "BGR_S_LAPS": abi.S_LAPS,
        "BGR_S_NEAR_TIES": abi.S_NEAR_TIES, "BGR_S_FIRST_TIE": abi.S_FIRST_TIE, "BGR_S_FALLBACKS": abi.S_FALLBACKS,
        "BGR_ST_PATH_END": abi.ST_PATH_END, "BGR_ST_LAPS_DONE": abi.ST_LAPS_DONE, "BGR_ST_TIMEOUT": abi.ST_TIMEOUT,
        "BGR_ST_NONFINITE": abi.ST_NONFINITE,
        "BGR_T_X": abi.T_X, "BGR_T_STEER": abi.T_STEER, "BGR_T_ACCEL": abi.T_ACCEL, "BGR_T_TARGET_IND": abi.T_TARGET_IND,
        "BGR_T_NEAREST_IND": abi.T_NEAREST_IND, "BGR_T_E_CT": abi.T_E_CT, "BGR_T_THETA_E": abi.T_THETA_E,
        "BGR_A_EPISODES": abi.A_EPISODES, "BGR_A_STEPS": abi.A_STEPS, "BGR_A_MIN_LAP_TIME": abi.A_MIN_LAP_TIME,
        "BGR_A_CONE_HITS": abi.A_CONE_HITS, "BGR_A_FINISHED": abi.A_FINISHED,
        "BGR_PHASE_STEER": abi.PHASE_STEER, "BGR_PHASE_ACCEL": abi.PHASE_ACCEL, "BGR_PHASE_MODEL": abi.PHASE_MODEL,
        "BGR_PHASE_ERRORS": abi.PHASE_ERRORS, "BGR_PHASE_KAPPA_GIVEN": abi.PHASE_KAPPA_GIVEN,
        "BGR_CMD_STEER": abi.CMD_STEER, "BGR_CMD_ACCEL": abi.CMD_ACCEL, "BGR_CMD_KAPPA": abi.CMD_KAPPA,
        "BGR_C_PID_INTEGRAL": abi.C_PID_INTEGRAL, "BGR_C_LQG_XHAT0": abi.C_LQG_XHAT0, "BGR_C_STEP": abi.C_STEP,
        "BGR_O_STEER": abi.O_STEER, "BGR_O_ACCEL": abi.O_ACCEL, "BGR_O_V_DESIRED": abi.O_V_DESIRED,
        "BGR_O_NEAREST_IND": abi.O_NEAREST_IND,
    }
    for k, v in names.items():
        assert h[k] == v, k


def test_struct_layouts_match_a_c_compiler(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc sees the header == the ctypes mirror."""
    src = tmp_path / "probe.c"
    fields = {
        "bgr_track_desc_t": (abi.TrackDesc, ["h_x", "n", "closed", "h_cones_xy", "n_cones", "cone_reach"]),
        "bgr_track_info_t": (abi.TrackInfo, ["n", "device", "smem_bytes_fp32", "guard_min", "ds_max", "rival_total",
                                               "rival_max", "rival_unlisted"]),
        "bgr_rollout_cfg_t": (abi.RolloutCfg, ["steer_kind", "max_steps", "flags", "dt", "hit_radius", "tie_eps",
                                                "sigma_cone", "noise_seed", "split_steps", "episode_id_base",
                                                "stanley_max_steer_rate", "stanley_alpha"]),
        "bgr_mpc_cfg_t": (abi.MpcCfg, ["horizon", "dt", "target_speed", "q", "r", "max_steer", "flags"]),
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{abi.HEADER_PATH}"', "int main(void){"]
    for cname, (_, fl) in fields.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fl:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    got = dict(ln.split() for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, (ct, fl) in fields.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for f in fl:
            assert int(got[f"{cname}.{f}"]) == getattr(ct, f).offset, f"{cname}.{f}"


def test_header_is_plain_c():
    """The header compiles as C99 and as C++17 with no CUDA / torch types in any signature."""
    for comp, std in (("gcc", "-std=c99"), ("g++", "-std=c++17")):
        subprocess.run([comp, std, "-Wall", "-Werror", "-fsyntax-only", "-x", "c" if comp == "gcc" else "c++",
                        abi.HEADER_PATH], check=True)
    text = re.sub(r"/\*.*?\*/", "", open(abi.HEADER_PATH).read(), flags=re.S)   # comments stripped
    assert "torch" not in text and "at::" not in text and "cudaStream_t" not in text and "#include <cuda" not in text


def test_bad_arguments_are_status_codes_not_crashes():
    lib = abi.load()
    assert lib.bgr_track_create(None, 0, None) == abi.BGR_ERR_BAD_ARG
    h = C.c_void_p()
    assert lib.bgr_track_create(None, 0, C.byref(h)) == abi.BGR_ERR_BAD_ARG
    assert b"NULL" in lib.bgr_last_error_string()
    cfg = bgr.RolloutConfig().to_struct()
    assert lib.bgr_rollout(None, C.byref(cfg), 1, None, None, None, None, None, None, None) == abi.BGR_ERR_BAD_ARG
    assert lib.bgr_device_count(None) == abi.BGR_ERR_BAD_ARG
    v = C.c_double()
    assert lib.bgr_lqg_design(-1.0, 0, None, None) == abi.BGR_ERR_BAD_ARG
    with pytest.raises(ValueError):
        bgr.RolloutConfig(steer="lqr").to_struct()
    with pytest.raises(ValueError):
        bgr.get_steering_controller("lqr")       # core/controllers.py:589


def test_lqg_design_is_host_math_and_matches_the_reference_gain():
    """control.dlqr at dt = 0.05 (SURVEY.md 8c KAT; golden value frozen from the reference run with a scipy DARE)."""
    K, g = bgr.lqg_design(0.05, 50)
    np.testing.assert_allclose(K, [22.181042932186813, 56.299306123839216], rtol=1e-12)
    assert g.shape == (50, 2) and np.all(g[:, 0] > 0) and np.all(g[:, 0] < 1)


def test_lqg_controller_exposes_the_reference_instance_attributes():
    """anything that inspects ``ctrl.A / B / C / Q / R / K / Q_kf / R_kf / P`` (core/controllers.py:83-109)
    must find them; ``P`` follows the covariance recursion of kalman_filter_update (:145, :160), one update per
    compute_acceleration call (:126).  Host math behind the ABI (bgr_lqg_covariance): no GPU needed."""
    dt = 0.05
    A = np.array([[1, 0], [dt, 1]]); Cm = np.array([[1, 0]])                       # noqa: E702  (:83-87)
    Qk, Rk = np.diag([0.1, 0.1]), np.array([[0.5]])                                 # :101-104
    P = np.eye(2)                                                                   # :109
    np.testing.assert_array_equal(bgr.lqg_covariance(dt, 0), P)
    for n in range(1, 40):
        P_pred = A @ P @ A.T + Qk                                                   # :145
        S = Cm @ P_pred @ Cm.T + Rk                                                 # :151
        K_kf = P_pred @ Cm.T @ np.linalg.inv(S)                                     # :154
        P = (np.eye(2) - K_kf @ Cm) @ P_pred                                        # :160
        np.testing.assert_allclose(bgr.lqg_covariance(dt, n), P, rtol=1e-13, atol=1e-15)
    ctl = bgr.LQGAccelerationController(dt)                # (constructor: host math only, no device call)
    np.testing.assert_array_equal(ctl.A, A)
    np.testing.assert_array_equal(ctl.B, [[dt], [0]])
    np.testing.assert_array_equal(ctl.C, Cm)
    np.testing.assert_array_equal(ctl.Q, np.diag([10.0, 100.0]))
    np.testing.assert_array_equal(ctl.R, [[0.001]])
    np.testing.assert_array_equal(ctl.Q_kf, Qk)
    np.testing.assert_array_equal(ctl.R_kf, Rk)
    assert ctl.K.shape == (1, 2) and ctl.x_hat.shape == (2, 1) and ctl.a_k_prev == 0.0 and ctl.v_desired == 0.0
    np.testing.assert_array_equal(ctl.P, np.eye(2))        # before the first call
    with pytest.raises(RuntimeError):
        bgr.lqg_covariance(-1.0, 3)


def test_build_id_names_this_build():
    """bench.py quotes ncu evidence only for the build id it was captured from (profiles/<dir>/build_id.txt)."""
    bid = bgr.build_id()
    assert isinstance(bid, str) and len(bid) >= 16 and bid != "unknown"


def test_no_cpu_fallback_without_a_device():
    """On a machine without CUDA every compute entry point fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ta = bgr.tracks.stadium_oval(64)
    with pytest.raises(abi.BgrError) as ei:
        bgr.Track.from_arrays(ta)
    assert ei.value.code == abi.BGR_ERR_CUDA
    with pytest.raises(abi.BgrError):
        bgr.fp32_peak_probe(0)
    with pytest.raises(abi.BgrError):
        bgr.PurePursuitController().compute_steering(
            type("S", (), dict(x=0.0, y=0.0, yaw=0.0, v=1.0))(), np.column_stack([np.arange(4.0)] * 3), 0)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import or execute it."""
    pkg = os.path.join(ROOT, "bgr_pathplanning_control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text.replace(
                    "oracle/loop.py", "").replace("oracle/philox.py", ""), f
    code = "import sys, bgr_pathplanning_control_b200; assert not [m for m in sys.modules if m.split('.')[0] == 'oracle']"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_streaming_groups_keep_group_order_and_their_own_host_threads(monkeypatch):
    """engine.rollout_groups_async (host logic only, no GPU): group g is always queued from lane thread g -- the
    library's copy / compute pipeline is per host thread -- in the requested launch order; ``wait()`` answers in the
    order the groups were given and is idempotent."""
    import threading
    import time
    from bgr_pathplanning_control_b200 import engine

    calls = []

    class Pending:
        def __init__(self, tag):
            self.tag, self.waits = tag, 0

        def wait(self):
            self.waits += 1
            return self.tag

    def fake_async(track, cfg, params, init, *, aggregate=False, out=None):
        calls.append((cfg, threading.current_thread().name, aggregate, out))
        time.sleep(0.01)
        return Pending((cfg, params, init))

    monkeypatch.setattr(engine, "rollout_async", fake_async)
    groups = [("pp", 1, 2), ("st_pid", 3, 4), ("st_lqg", 5, 6)]
    outs = ["o0", "o1", "o2"]
    seen = {}
    for batch in range(3):
        pend = bgr.rollout_groups_async("track", groups, aggregate=True, outs=outs, order=[2, 1, 0])
        assert pend.wait() == [("pp", 1, 2), ("st_pid", 3, 4), ("st_lqg", 5, 6)]
        assert pend.wait() is pend.wait()
        for cfg, thread, aggregate, out in calls[-3:]:
            assert aggregate is True and out == outs[[g[0] for g in groups].index(cfg)]
            assert seen.setdefault(cfg, thread) == thread          # the same lane thread in every batch
    assert len(set(seen.values())) == 3                            # one thread per group
    assert len(calls) == 9
    with pytest.raises(IndexError):
        bgr.rollout_groups_async("track", groups, order=[3])

    # one group fails: the others are still waited for (their arrays must not stay in flight), then the error surfaces
    waited = []

    class Slow(Pending):
        def wait(self):
            waited.append(self.tag[0])
            return super().wait()

    def failing_async(track, cfg, params, init, *, aggregate=False, out=None):
        if cfg == "st_pid":
            raise RuntimeError("boom")
        return Slow((cfg, params, init))

    monkeypatch.setattr(engine, "rollout_async", failing_async)
    pend = bgr.rollout_groups_async("track", groups)
    for _ in range(2):
        with pytest.raises(RuntimeError, match="boom"):
            pend.wait()
    assert waited == ["pp", "st_lqg"]
