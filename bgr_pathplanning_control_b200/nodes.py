"""Host-side callers of the hot path, with the reference's node protocol.

* ``Controller``         -- drop-in for nodes/controller.py:14-94: same constructor, same ``update(data)`` dict protocol
                            (reads ``path, cx, cy, curve, car_state, target_ind, current_time``; writes ``states, v_log,
                            acceleration, steering``), the laws evaluated by the CUDA kernels through the scalar drop-in
                            classes.  The simulator call at :90 is replaced by a pluggable ``actuator`` (the FSDS bridge
                            is out of scope); the command the reference would ship -- ``(-steering, acceleration)`` and
                            its throttle / brake normalisation (providers/sim/sim_util.py:186-209) -- is exposed as
                            ``data["controls"]``.
* ``BatchedController``  -- the same protocol for MANY episodes at once: one ``bgr_rollout`` instead of a Python loop, fed
                            from the planner's ``data`` dict; per-episode ``States`` histories come back in the
                            reference's container so core/visualization.py-style plotting code works unchanged.
* ``path_from_planner``  -- centreline ingestion from fsd-path-planning output ``(M, 4) = [s, x, y, curvature]`` with the
                            ``cy = -path[:, 2]`` flip of nodes/planner.py:68,88-89.
* ``SimulationLogger``   -- preformance_analysis/simulation_logger.py:3-20 (same CSV schema), plus ``log_episode`` that
                            writes one GPU episode in that schema so ``metrics_calculator`` / ``plot_suite`` consume it.

Everything numeric goes through the C ABI; there is no CPU fallback here either.
"""
from __future__ import annotations

import csv
import logging

import numpy as np

from . import _abi as abi
from . import engine
from .car_state import State, States
from .controllers import AccelerationPIDController, PurePursuitController
from .metrics_calculator import CSV_COLUMNS, trajectory_frame
from .vehicle_config import conf

log = logging.getLogger("Controller")


def sim_controls(di, ai):
    """The normalised command of providers/sim/sim_util.py:186-209 for steering ``di`` [rad] and acceleration ``ai``:
    ``(steering, throttle, brake)``.  Verbatim, including its quirk: ``brake = clip(-ai / MAX_DECEL, 0, 1)`` with a
    NEGATIVE ``MAX_DECEL`` is always 0."""
    steering = di / conf.MAX_STEER
    if ai >= 0:
        return steering, float(np.clip(ai / conf.MAX_ACCEL, 0.0, 1.0)), 0.0
    return steering, 0.0, float(np.clip(-ai / conf.MAX_DECEL, 0.0, 1.0))


def path_from_planner(path):
    """fsd-path-planning's ``(M, 4)`` array ``[s, x, y, curvature]`` -> ``(cx, cy, curve)`` exactly as
    ``Planner.update`` exposes them (nodes/planner.py:88-89): ``cy`` is NEGATED."""
    path = np.asarray(path, dtype=np.float64)
    if path.ndim != 2 or path.shape[1] < 4:
        raise IndexError("planner path must be (M, 4) = [s, x, y, curvature]")
    return path[:, 1].copy(), -path[:, 2], path[:, 3].copy()


class Controller:
    """nodes/controller.py:14-94."""

    def __init__(self, actuator=None):
        log.info("Initializing Controller with PID and Pure Pursuit controllers")
        self.accel_ctrl = AccelerationPIDController(kp=conf.kp_accel, ki=conf.ki_accel, kd=conf.kd_accel,
                                                    setpoint=conf.TARGET_SPEED)                     # :17-22
        self.steer_ctrl = PurePursuitController(look_ahead_distance=conf.LOOKAHEAD_DISTANCE)        # :23
        self.states = States()                                                                      # :24
        self.actuator = actuator        # callable(steering_cmd, acceleration) standing in for sim_car_controls (:90)
        self.target_ind = 0             # the index compute_steering returned last (a local of update() in the reference,
                                        # logged at :81; kept here for callers that have no Planner node)

    def update(self, data):
        try:
            path = data.get("path")
            cx = data.get("cx")
            cy = data.get("cy")
            curve = data.get("curve")
            car_state = data.get("car_state")
            target_ind = data.get("target_ind")
            if any(x is None for x in [path, cx, cy, car_state]):                                   # :37-41
                log.warning("Missing required data for controller update: "
                            f"path={path is not None}, cx={cx is not None}, "
                            f"cy={cy is not None}, car_state={car_state is not None}")
                return
            state_modifier = State(x=car_state.x, y=-car_state.y, yaw=car_state.yaw, v=car_state.v,   # :43-53
                                   v_linear=getattr(car_state, "v_linear", 0), v_angular=getattr(car_state, "v_angular", 0),
                                   a_angular=getattr(car_state, "a_angular", 0), a_linear=getattr(car_state, "a_linear", 0),
                                   timestamp=getattr(car_state, "timestamp", 0))
            path_track = np.column_stack((np.arange(len(cx)), cx, cy))                              # :56
            # the controller is fed the RAW state, not the y-flipped one (:61)
            steering, target_ind = self.steer_ctrl.compute_steering(car_state, path_track, target_ind or 0)
            curvature = curve[target_ind] if target_ind < len(curve) else curve[-1]                 # :62
            acceleration = self.accel_ctrl.compute_acceleration(car_state.v, curvature)             # :63
            current_time = data.get("current_time", 0)
            self.states.append(current_time, state_modifier, steering if steering else 0.0,         # :66-77
                               acceleration if acceleration else 0.0,
                               self.accel_ctrl.maxspeed if self.accel_ctrl.maxspeed else 0.0,
                               v_linear=getattr(car_state, "v_linear", 0), v_angular=getattr(car_state, "v_angular", 0),
                               a_linear=getattr(car_state, "a_angular", 0), a_angular=getattr(car_state, "a_linear", 0),
                               timestamp=getattr(car_state, "timestamp", 0))
            data["states"] = self.states                                                            # :78
            data["v_log"] = self.accel_ctrl.maxspeed                                                # :84-86
            data["acceleration"] = acceleration
            data["steering"] = steering
            # data["target_ind"] is the PLANNER's to write (nodes/planner.py:101): the reference controller never
            # publishes its own advanced index, so a planner that returns early leaves its stale index in place
            self.target_ind = int(target_ind)
            data["controls"] = sim_controls(-steering, acceleration)                                # :90
            if self.actuator is not None:
                self.actuator(-steering, acceleration)
        except Exception as e:                                                                      # :92-94
            log.error(f"Error in controller update: {str(e)}", exc_info=True)
            raise


class BatchedController:
    """Many closed-loop episodes from one planner ``data`` dict: the batched entry point behind the node protocol.

    ``run(data, params, init_state)`` uploads ``data["cx"], data["cy"], data["curve"]`` once (cached by content),
    runs one ``bgr_rollout`` and returns the ``RolloutResult``; ``states(result, episode)`` rebuilds the reference's
    ``States`` history for one episode from the trajectory dump."""

    def __init__(self, steer="pure_pursuit", accel="pid", model="kinematic", device=0, **config):
        self.config = engine.RolloutConfig(steer=steer, accel=accel, model=model, **config)
        self.device = device
        self._track = None
        self._key = None

    def _track_for(self, cx, cy, curve, closed):
        cx, cy, curve = (np.ascontiguousarray(a, dtype=np.float64) for a in (cx, cy, curve))
        key = (len(cx), bool(closed), cx.tobytes(), cy.tobytes(), curve.tobytes())
        if key != self._key:
            if self._track is not None:
                self._track.close()
            self._track = engine.Track(cx, cy, curve, closed=closed, device=self.device)
            self._key = key
        return self._track

    def run(self, data, params=None, init_state=None, *, n_episodes=None, closed=False, traj=True, aggregate=True):
        cx, cy, curve = data.get("cx"), data.get("cy"), data.get("curve")
        car_state = data.get("car_state")
        if any(x is None for x in [cx, cy, car_state]):
            log.warning("Missing required data for batched controller run")
            return None
        curve = np.zeros(len(cx)) if curve is None else curve
        if params is None:
            params = engine.default_params(n_episodes or 1)
        E = len(params)
        if init_state is None:       # every episode starts from the provider's car state
            init_state = np.tile(np.array([[car_state.x, car_state.y, car_state.yaw, car_state.v,
                                            getattr(car_state, "r", 0.0), getattr(car_state, "beta", 0.0)]],
                                          dtype=np.float64), (E, 1))
        track = self._track_for(cx, cy, curve, closed)
        result = engine.rollout(track, self.config, params, init_state, traj=traj, aggregate=aggregate)
        data["rollout"] = result
        return result

    def states(self, result, episode=0):
        steps = int(result.status[episode, abi.S_STEPS])
        tr = result.traj[episode]
        tr = tr.cpu().numpy() if hasattr(tr, "cpu") else np.asarray(tr)
        return States.from_trajectory(tr, self.config.dt, steps)


class SimulationLogger:
    """preformance_analysis/simulation_logger.py:3-20: same file format, same ``log`` signature."""

    def __init__(self, filename="simulation_log.csv"):
        self.filename = filename
        with open(self.filename, mode="w", newline="") as file:
            csv.writer(file).writerow(CSV_COLUMNS)                                                   # :8-10

    def log(self, timestamp, x, y, velocity, steering_angle, throttle, brake, lateral_error, heading_error, yaw_rate,
            lateral_accel):
        with open(self.filename, mode="a", newline="") as file:                                      # :14-17
            csv.writer(file).writerow([timestamp, x, y, velocity, steering_angle, throttle, brake, lateral_error,
                                       heading_error, yaw_rate, lateral_accel])

    def log_episode(self, result, episode=0, dt=conf.dt):
        """Append every logged row of one GPU episode (trajectory dump) in the reference's column schema."""
        frame = trajectory_frame(result, episode, dt, conf.MAX_ACCEL, conf.MAX_DECEL)
        frame.to_csv(self.filename, mode="a", header=False, index=False, float_format="%.17g")
        return frame

    def close(self):
        pass
