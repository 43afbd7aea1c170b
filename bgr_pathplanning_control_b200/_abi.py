"""ctypes binding of the C ABI in ``include/bgr_rollout.h`` (``libbgr_rollout.so``, built in-tree by
``bgr_pathplanning_control_b200/csrc/Makefile``).

There is NO CPU fallback: if the shared library is missing ``load()`` raises, and every compute entry
point of the library fails with ``BGR_ERR_CUDA`` when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
# BGR_ROLLOUT_LIB: an A/B build of the SAME library with a different compile-time knob (csrc/Makefile VARIANT=...);
# a measurement aid for kernel work, not a second backend -- every build is sm_100a CUDA behind the same ABI
LIB_PATH = os.environ.get("BGR_ROLLOUT_LIB") or os.path.join(HERE, "libbgr_rollout.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "bgr_rollout.h")

# ---- constants restated from include/bgr_rollout.h (tests/test_abi.py checks them against the header) ----
BGR_ABI_VERSION = 3
BGR_OK, BGR_ERR_BAD_ARG, BGR_ERR_CUDA, BGR_ERR_TOO_LARGE, BGR_ERR_ALLOC, BGR_ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
STEER_PURE_PURSUIT, STEER_STANLEY = 0, 1
ACCEL_PID, ACCEL_LQG = 0, 1
MODEL_KINEMATIC, MODEL_DYNAMIC = 0, 1
PRECISION_FP32, PRECISION_FP64 = 0, 1
(P_LF, P_KP, P_KI, P_KD, P_TARGET_SPEED, P_K_EXPO, P_K_STANLEY, P_K_SOFT,
 P_MU, P_C_F, P_C_R, P_MASS, P_I_Z) = range(13)
N_PARAMS = 16
X, Y, YAW, V, R, BETA = range(6)
N_STATE = 6
(M_LAT_MEAN, M_LAT_STD, M_HEAD_MEAN, M_HEAD_STD, M_LAP_TIME, M_MAX_ABS_LAT,
 M_X, M_Y, M_YAW, M_V, M_R, M_BETA) = range(12)
N_METRICS = 12
(S_FLAGS, S_STEPS, S_TARGET_IND, S_NEAREST_IND, S_CONE_HITS, S_LAPS, S_NEAR_TIES, S_FIRST_TIE,
 S_FALLBACKS) = range(9)
N_STATUS = 10
ST_PATH_END, ST_LAPS_DONE, ST_TIMEOUT, ST_NONFINITE = 1, 2, 4, 8
(T_X, T_Y, T_YAW, T_V, T_R, T_BETA, T_STEER, T_ACCEL, T_TARGET_IND, T_NEAREST_IND, T_E_CT,
 T_THETA_E) = range(12)
N_TRAJ = 12
(A_EPISODES, A_STEPS, A_SUM_LAT_MEAN, A_SUM_LAT_STD, A_SUM_LAP_TIME, A_MIN_LAP_TIME, A_CONE_HITS,
 A_FINISHED) = range(8)
N_AGG = 8
CONE_GUARD = 2.0
MAX_CONES = 512
CFG_AUDIT = 1
CFG_STANLEY_SHADOWED = 2
CFG_DYNAMIC_SCHEDULE = 4
CFG_KEEP_ORDER = 8
MPC_EXPLICIT_ROLLOUT = 1
PHASE_STEER, PHASE_ACCEL, PHASE_MODEL, PHASE_ERRORS, PHASE_KAPPA_GIVEN = 1, 2, 4, 8, 16
CMD_STEER, CMD_ACCEL, CMD_KAPPA = 0, 1, 2
N_CMD = 4
(C_PID_INTEGRAL, C_PID_LAST_INPUT, C_PID_HAS_LAST, C_LQG_XHAT0, C_LQG_XHAT1, C_LQG_A_PREV,
 C_STEP, C_PREV_STEER) = range(8)
N_CTRL = 8
O_STEER, O_ACCEL, O_V_DESIRED, O_KAPPA, O_E_CT, O_THETA_E, O_NEAREST_IND = range(7)
N_OUT = 8
SYSID_COLS = 9

EXPORTS = (
    "bgr_abi_version", "bgr_last_error_string", "bgr_device_count", "bgr_build_id", "bgr_stream_create", "bgr_stream_synchronize",
    "bgr_stream_destroy",
    "bgr_track_create", "bgr_track_destroy", "bgr_track_info", "bgr_track_tables",
    "bgr_rollout", "bgr_rollout_host", "bgr_rollout_host_async", "bgr_step_host", "bgr_lqg_design", "bgr_lqg_covariance",
    "bgr_mpc_eval", "bgr_mpc_eval_host", "bgr_mpc_bank_create", "bgr_mpc_bank_destroy", "bgr_mpc_eval_bank",
    "bgr_mpc_eval_bank_host", "bgr_mpc_eval_bank_host_async", "bgr_identify_dynamics", "bgr_identify_dynamics_host",
    "bgr_fp32_peak_probe", "bgr_fp64_peak_probe", "bgr_last_kernel_ms", "bgr_launch_count",
)

_dp = C.POINTER(C.c_double)


class TrackDesc(C.Structure):
    _fields_ = [("h_x", _dp), ("h_y", _dp), ("h_kappa", _dp), ("n", C.c_int32), ("closed", C.c_int32),
                ("h_cones_xy", _dp), ("n_cones", C.c_int32), ("reserved", C.c_int32), ("cone_reach", C.c_double)]


class TrackInfo(C.Structure):
    _fields_ = [("n", C.c_int32), ("closed", C.c_int32), ("n_cones", C.c_int32), ("device", C.c_int32),
                ("smem_bytes_fp32", C.c_int64), ("smem_bytes_fp64", C.c_int64),
                ("guard_min", C.c_double), ("guard_median", C.c_double), ("ds_min", C.c_double),
                ("ds_max", C.c_double), ("rival_total", C.c_int64), ("rival_max", C.c_int32),
                ("rival_unlisted", C.c_int32), ("twin_count", C.c_int32), ("reserved", C.c_int32)]


class RolloutCfg(C.Structure):
    _fields_ = [("steer_kind", C.c_int32), ("accel_kind", C.c_int32), ("model_kind", C.c_int32),
                ("precision", C.c_int32), ("max_steps", C.c_int32), ("laps", C.c_int32),
                ("laps_target", C.c_int32), ("flags", C.c_int32),
                ("dt", C.c_double), ("wheelbase", C.c_double), ("l_f", C.c_double), ("l_r", C.c_double),
                ("max_steer", C.c_double), ("max_accel", C.c_double), ("max_decel", C.c_double),
                ("hit_radius", C.c_double), ("tie_eps", C.c_double),
                ("sigma_xy", C.c_double), ("sigma_yaw", C.c_double), ("sigma_v", C.c_double),
                ("sigma_cone", C.c_double), ("noise_seed", C.c_uint32), ("split_steps", C.c_uint32),
                ("episode_id_base", C.c_int64), ("stanley_max_steer_rate", C.c_double),
                ("stanley_alpha", C.c_double),
                ("replan_horizon", C.c_int32), ("replan_stride", C.c_int32), ("replan_after", C.c_int32),
                ("reserved2", C.c_int32), ("planner_lookahead", C.c_double), ("stanley_lookahead", C.c_double)]


class MpcCfg(C.Structure):
    _fields_ = [("horizon", C.c_int32), ("precision", C.c_int32), ("dt", C.c_double), ("wheelbase", C.c_double),
                ("target_speed", C.c_double), ("q", C.c_double * 4), ("r", C.c_double * 2),
                ("max_accel", C.c_double), ("max_decel", C.c_double), ("max_steer", C.c_double),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class BgrError(RuntimeError):
    """A C-ABI call returned a negative status."""

    def __init__(self, code, message):
        super().__init__(f"[bgr status {code}] {message}")
        self.code = code


_lib = None


def load():
    """Load ``libbgr_rollout.so`` (once) and declare every prototype.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is not built.  Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C bgr_pathplanning_control_b200/csrc`).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.bgr_abi_version.restype = C.c_int
    lib.bgr_abi_version.argtypes = []
    lib.bgr_last_error_string.restype = C.c_char_p
    lib.bgr_last_error_string.argtypes = []
    lib.bgr_device_count.restype = C.c_int
    lib.bgr_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.bgr_track_create.restype = C.c_int
    lib.bgr_track_create.argtypes = [C.POINTER(TrackDesc), C.c_int, C.POINTER(vp)]
    lib.bgr_track_destroy.restype = None
    lib.bgr_track_destroy.argtypes = [vp]
    lib.bgr_track_info.restype = C.c_int
    lib.bgr_track_info.argtypes = [vp, C.POINTER(TrackInfo)]
    lib.bgr_track_tables.restype = C.c_int
    lib.bgr_track_tables.argtypes = [vp, vp, vp]
    lib.bgr_rollout.restype = C.c_int
    lib.bgr_rollout.argtypes = [vp, C.POINTER(RolloutCfg), i64, vp, vp, vp, vp, vp, vp, vp]
    lib.bgr_rollout_host.restype = C.c_int
    lib.bgr_rollout_host.argtypes = [vp, C.POINTER(RolloutCfg), i64, vp, vp, vp, vp, vp, vp, vp]
    lib.bgr_step_host.restype = C.c_int
    lib.bgr_step_host.argtypes = [vp, C.POINTER(RolloutCfg), i32, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.bgr_lqg_design.restype = C.c_int
    lib.bgr_lqg_design.argtypes = [C.c_double, i32, vp, vp]
    lib.bgr_lqg_covariance.restype = C.c_int
    lib.bgr_lqg_covariance.argtypes = [C.c_double, i32, vp]
    lib.bgr_build_id.restype = C.c_char_p
    lib.bgr_build_id.argtypes = []
    lib.bgr_stream_synchronize.restype = C.c_int
    lib.bgr_stream_synchronize.argtypes = [vp]
    lib.bgr_stream_create.restype = C.c_int
    lib.bgr_stream_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.bgr_stream_destroy.restype = C.c_int
    lib.bgr_stream_destroy.argtypes = [vp]
    lib.bgr_rollout_host_async.restype = C.c_int
    lib.bgr_rollout_host_async.argtypes = [vp, C.POINTER(RolloutCfg), i64, vp, vp, vp, vp, vp, vp]
    lib.bgr_mpc_eval.restype = C.c_int
    lib.bgr_mpc_eval.argtypes = [vp, C.POINTER(MpcCfg), i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.bgr_mpc_eval_host.restype = C.c_int
    lib.bgr_mpc_eval_host.argtypes = [vp, C.POINTER(MpcCfg), i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.bgr_mpc_bank_create.restype = C.c_int
    lib.bgr_mpc_bank_create.argtypes = [C.c_int, vp, i32, i32, C.POINTER(vp)]
    lib.bgr_mpc_bank_destroy.restype = None
    lib.bgr_mpc_bank_destroy.argtypes = [vp]
    lib.bgr_mpc_eval_bank.restype = C.c_int
    lib.bgr_mpc_eval_bank.argtypes = [vp, C.POINTER(MpcCfg), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.bgr_mpc_eval_bank_host.restype = C.c_int
    lib.bgr_mpc_eval_bank_host.argtypes = [vp, C.POINTER(MpcCfg), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.bgr_mpc_eval_bank_host_async.restype = C.c_int
    lib.bgr_mpc_eval_bank_host_async.argtypes = [vp, C.POINTER(MpcCfg), i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.bgr_fp64_peak_probe.restype = C.c_int
    lib.bgr_fp64_peak_probe.argtypes = [C.c_int, vp]
    lib.bgr_identify_dynamics.restype = C.c_int
    lib.bgr_identify_dynamics.argtypes = [i64, i32, vp, vp, vp, vp, vp]
    lib.bgr_identify_dynamics_host.restype = C.c_int
    lib.bgr_identify_dynamics_host.argtypes = [i64, i32, vp, vp, vp, vp, vp]
    lib.bgr_fp32_peak_probe.restype = C.c_int
    lib.bgr_fp32_peak_probe.argtypes = [C.c_int, _dp, _dp]
    lib.bgr_last_kernel_ms.restype = C.c_int
    lib.bgr_last_kernel_ms.argtypes = [_dp]
    lib.bgr_launch_count.restype = i64
    lib.bgr_launch_count.argtypes = []
    if lib.bgr_abi_version() != BGR_ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.bgr_abi_version()} != {BGR_ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(code):
    if code != BGR_OK:
        raise BgrError(code, load().bgr_last_error_string().decode("utf-8", "replace"))


def header_constants():
    """``#define BGR_* <int|float>`` pairs parsed from the header (used by the ABI test)."""
    out = {}
    with open(HEADER_PATH) as f:
        for m in re.finditer(r"^#define\s+(BGR_\w+)\s+\(?(-?[0-9.]+)\)?", f.read(), re.M):
            out[m.group(1)] = float(m.group(2)) if "." in m.group(2) else int(m.group(2))
    return out


def header_functions():
    """Names of every function the header declares."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"^\s*(?:int|void|const char\*|int64_t)\s+(bgr_\w+)\s*\(", text, re.M)))
