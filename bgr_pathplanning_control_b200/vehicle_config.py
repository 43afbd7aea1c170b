"""Vehicle parameters: the same names and values as the reference's ``Vehicle_config``
(vehicle_config.py:4-39).  They are the column defaults of the per-episode parameter tensor and of
``RolloutConfig``; nothing mutates them.
"""
import math


class Vehicle_config:
    # Vehicle parameters                                   vehicle_config.py
    m = 255                    # mass [kg]                               :7
    I_z = 1700                 # yaw inertia [kg m^2]                    :8
    l_f = 0.9                  # CG -> front axle [m]                    :9
    l_r = 0.9                  # CG -> rear axle [m]                     :10
    c_f = 16000                # cornering stiffness front [N/rad]       :11
    c_r = 17000                # cornering stiffness rear [N/rad]        :12
    mu = 1.0                   # friction coefficient                    :13
    TARGET_SPEED = 25 / 3.6    # [m/s]                                   :14
    WB = l_f + l_r             #                                         :16
    kp_accel = 0.35            # PID gains                               :19-21
    ki_accel = 0.4
    kd_accel = 0.3
    k_expo = 0.05              # curvature -> speed decay                :24
    k_stanley = 0.7            # unused by the live Stanley law          :26
    k_soft_stanley = 0.01      # unused by the live Stanley law          :27
    MAX_ACCEL = 1.5            # [m/s^2]                                 :30
    MAX_DECEL = -2.0           # [m/s^2]                                 :31
    MAX_SPEED = 40 / 3.6       # [m/s] (unused by any law)               :33
    MAX_STEER = math.radians(30)   # [rad]  np.deg2rad(30)               :35
    LOOKAHEAD_DISTANCE = 5.5   # [m]                                     :37
    dt = 0.05                  # [s]                                     :39


conf = Vehicle_config
