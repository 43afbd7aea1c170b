"""``State`` / ``States`` with the reference's fields and call conventions (core/data/car_state.py:63-154).
``State.update(a, delta)`` advances the dynamic bicycle model (core/data/car_state.py:118-125,178-223) on the
GPU through ``bgr_step_host``; ``States`` is the host-side history container the plots consume.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi as abi
from . import engine
from .vehicle_config import conf


class States:
    """core/data/car_state.py:63-101 -- plain Python lists, same attribute names."""

    def __init__(self):
        self.x, self.y, self.yaw, self.v, self.t = [], [], [], [], []
        self.steering, self.acceleration, self.v_log = [], [], []
        self.v_linear, self.v_angular = [], []
        self.a_longitudinal, self.a_lateral = [], []
        self.timestamp = []

    def append(self, t, state, steering=None, acceleration=None, v_log=None, v_linear=None, v_angular=None,
               a_linear=None, a_angular=None, timestamp=None):
        self.x.append(state.x)
        self.y.append(state.y)
        self.yaw.append(state.yaw)
        self.v.append(state.v)
        self.t.append(t)
        self.v_linear.append(v_linear)
        self.v_angular.append(v_angular)
        self.a_longitudinal.append(a_linear.x_val if a_linear else 0)
        self.a_lateral.append(a_linear.y_val if a_linear else 0)
        self.timestamp.append(timestamp)
        if steering is not None:
            self.steering.append(steering)
        if acceleration is not None:
            self.acceleration.append(acceleration)
        if v_log is not None:
            self.v_log.append(v_log)

    @classmethod
    def from_trajectory(cls, traj, dt=conf.dt, steps=None):
        """Fill a ``States`` from one episode's trajectory dump (steps, BGR_N_TRAJ) so that the reference's
        plotting code (core/visualization.py) works on GPU results unchanged."""
        s = cls()
        traj = np.asarray(traj)
        n = len(traj) if steps is None else int(steps)
        for k in range(n):
            row = traj[k]
            s.x.append(float(row[abi.T_X]))
            s.y.append(float(row[abi.T_Y]))
            s.yaw.append(float(row[abi.T_YAW]))
            s.v.append(float(row[abi.T_V]))
            s.t.append(k * dt)
            s.v_linear.append(None)
            s.v_angular.append(None)
            s.a_longitudinal.append(0)
            s.a_lateral.append(0)
            s.timestamp.append(k * dt)
            s.steering.append(float(row[abi.T_STEER]))
            s.acceleration.append(float(row[abi.T_ACCEL]))
        return s


class State:
    """core/data/car_state.py:103-154."""

    def __init__(self, x=0.0, y=0.0, yaw=0.0, v=0.0, beta=0.0, r=0.0, v_linear=0, v_angular=0, a_angular=0,
                 a_linear=0, timestamp=0):
        self.x = x
        self.y = y
        self.yaw = yaw
        self.v = v
        self.beta = beta
        self.r = r
        self.update_positions()
        self.v_linear = v_linear
        self.v_angular = v_angular
        self.a_angular = a_angular
        self.a_linear = a_linear
        self.timestamp = timestamp

    def update(self, a, delta):
        """One dynamic-bicycle step with ``dt = conf.dt``.  Argument order is (a, delta) as in the reference
        (:118); the model itself takes [delta, a] (:121)."""
        from .controllers import _dummy_track, _step   # late import: controllers imports nothing from here

        cfg = engine.RolloutConfig(model="dynamic", precision="fp64")
        row = np.array([[float(self.x), float(self.y), float(self.yaw), float(self.v), float(self.r),
                         float(self.beta)]])
        cmd = np.zeros((1, abi.N_CMD))
        cmd[0, abi.CMD_STEER], cmd[0, abi.CMD_ACCEL] = float(delta), float(a)
        _step(_dummy_track(), cfg, abi.PHASE_MODEL, engine.default_params(1, np.float64), row, 0, None, cmd)
        self.x, self.y, self.yaw, self.v, self.r, self.beta = (float(v) for v in row[0])
        self.update_positions()

    def update_positions(self):
        """:127-134 (note: the reference uses l_f for the rear axle and l_r for the front one)."""
        self.rear_x = self.x - (conf.l_f * math.cos(self.yaw))
        self.rear_y = self.y - (conf.l_f * math.sin(self.yaw))
        self.front_x = self.x + (conf.l_r * math.cos(self.yaw))
        self.front_y = self.y + (conf.l_r * math.sin(self.yaw))

    def calc_distance(self, point_x, point_y, use_front=True):
        """:136-154 distance from the front (default) or rear axle point to a point."""
        if use_front:
            dx = self.front_x - point_x
            dy = self.front_y - point_y
        else:
            dx = self.rear_x - point_x
            dy = self.rear_y - point_y
        return math.hypot(dx, dy)
