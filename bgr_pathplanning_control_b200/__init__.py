"""bgr_pathplanning_control_b200 -- B200-native batched closed-loop rollout engine for the one data-parallel
hot path of grimgrimberg/BGR_PathPlanning_Control (tracking law -> bicycle-model step -> lap metrics).

Layout: ``csrc/`` sm_100a CUDA kernels + the C ABI (``include/bgr_rollout.h``);  ``_abi.py`` ctypes binding;
``engine.py`` batched entry points;  ``controllers.py`` / ``car_state.py`` / ``metrics_calculator.py`` /
``vehicle_config.py`` the host-side mirror of the reference's interface;  ``tracks.py`` / ``sweeps.py``
synthetic inputs;  ``sharding.py`` multi-GPU plumbing;  ``roofline.py`` the frozen flop/byte counts.
Importing the package does not load the CUDA library; the first call does and raises if it is not built.
"""
import os as _os

# More hardware work queues than the default 8: the host-buffer pipelines use four streams per host thread, and streams
# that share a queue wait for each other's event waits (measured: a mixed-controller group of three host threads ran
# its kernels one after the other).  Only effective when set before the process creates its CUDA context; an explicit
# setting of the variable wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _abi, tracks, sweeps, sharding, roofline          # noqa: F401
from .vehicle_config import Vehicle_config, conf               # noqa: F401
from .engine import (Track, RolloutConfig, RolloutResult, MpcConfig, MpcResult, rollout, rollout_groups, mpc_evaluate,   # noqa: F401
                     default_params, initial_states, candidate_bank, lqg_design, lqg_covariance, fp32_peak_probe,
                     last_kernel_ms, launch_count, identify_dynamics, rollout_async, PendingRollout, build_id,
                     release_group_streams, MpcBank, fp64_peak_probe, mpc_evaluate_async, PendingMpc,
                     rollout_groups_async, PendingGroups)
from .controllers import (PurePursuitController, StanleyController, ShadowedStanleyController, AccelerationPIDController,            # noqa: F401
                          LQGAccelerationController, MPCController, get_steering_controller,
                          normalize_angle, find_target_point, compute_control_commands)
from .car_state import State, States                            # noqa: F401
from .nodes import Controller, BatchedController, SimulationLogger, path_from_planner, sim_controls   # noqa: F401
from . import metrics_calculator                                # noqa: F401

__version__ = "0.2.0"
