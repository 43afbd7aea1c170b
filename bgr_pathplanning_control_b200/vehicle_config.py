"""Vehicle parameters under the reference's names (``Vehicle_config`` / ``conf``, vehicle_config.py:4-39).

They are the column defaults of the per-episode parameter tensor (``engine.default_params``) and of
``RolloutConfig``; nothing mutates them.  The table below is the single source: (attribute, value, line of the
reference's vehicle_config.py that defines it, what it is).
"""
import math

_TABLE = (
    ("m", 255, 7, "mass [kg]"),
    ("I_z", 1700, 8, "yaw inertia [kg m^2]"),
    ("l_f", 0.9, 9, "centre of gravity -> front axle [m]"),
    ("l_r", 0.9, 10, "centre of gravity -> rear axle [m]"),
    ("c_f", 16000, 11, "front cornering stiffness [N/rad]"),
    ("c_r", 17000, 12, "rear cornering stiffness [N/rad]"),
    ("mu", 1.0, 13, "tyre-road friction coefficient"),
    ("TARGET_SPEED", 25 / 3.6, 14, "cruise speed [m/s]"),
    ("kp_accel", 0.35, 19, "speed PID proportional gain"),
    ("ki_accel", 0.4, 20, "speed PID integral gain"),
    ("kd_accel", 0.3, 21, "speed PID derivative gain"),
    ("k_expo", 0.05, 24, "speed target decay per unit curvature"),
    ("k_stanley", 0.7, 26, "Stanley gain of the shadowed class body (unused by the live law)"),
    ("k_soft_stanley", 0.01, 27, "Stanley softening of the shadowed class body (unused by the live law)"),
    ("MAX_ACCEL", 1.5, 30, "acceleration command ceiling [m/s^2]"),
    ("MAX_DECEL", -2.0, 31, "acceleration command floor [m/s^2]"),
    ("MAX_SPEED", 40 / 3.6, 33, "[m/s]; no law reads it"),
    ("MAX_STEER", math.radians(30), 35, "steering clip [rad] (np.deg2rad(30))"),
    ("LOOKAHEAD_DISTANCE", 5.5, 37, "pure-pursuit look-ahead [m]"),
    ("dt", 0.05, 39, "control period [s]"),
)


class Vehicle_config:
    """Class attributes only, exactly like the reference (its ``@dataclass`` has no annotated fields)."""


for _name, _value, _line, _what in _TABLE:
    setattr(Vehicle_config, _name, _value)
Vehicle_config.WB = Vehicle_config.l_f + Vehicle_config.l_r      # wheelbase, vehicle_config.py:16

conf = Vehicle_config
