"""Frozen ALGORITHMIC work counts per unit -- the numerators of the roofline.

Counting rule: each scalar + - * / compare/min/max/abs = 1 flop; each transcendental (sqrt/hypot core, sin,
cos, tan, atan2, exp, floor-mod) = 1 flop (also tallied as T); counted from the reference source lines.
The windowed nearest search is counted at the window the GPU actually visits (W_NEAREST evaluations per
step), not at the reference's O(N) global scan, so useless work is not rewarded.
"""

FLOPS = {
    # steering laws
    "pp_scan_iter": 7,           # core/controllers.py:213-216
    "pp_law": 12,                # :225-229
    "stanley_law": 26,           # :361-381 (without the nearest search)
    "nearest_eval": 7,           # one candidate of the nearest search (:353-356)
    # longitudinal laws
    "pid": 17,                   # :32-47, :58-73 + restated simple-pid
    "lqg": 21,                   # :117-138 with the shared gain sequence
    # vehicle models
    "kinematic": 12,             # oracle-defined
    "dynamic": 38,               # core/data/car_state.py:203-220
    # metrics
    "errors": 16,                # :371-376 on the true state
    "accumulate": 8,             # n, sum, sum of squares, max
    "mpc_candidate_step": 29,    # :522-527 + :545
}

PP_SCAN_ITERS = 2.7              # 1 + v dt / ds at ds = 0.209 m, v = 6.9 m/s
W_NEAREST = 3.7                  # local descent: the start, ~1.7 forward moves, one stop test


def flops_per_step(steer="pure_pursuit", accel="pid", model="kinematic"):
    f = FLOPS
    total = f["errors"] + f["accumulate"] + W_NEAREST * f["nearest_eval"]
    if steer == "pure_pursuit":
        total += PP_SCAN_ITERS * f["pp_scan_iter"] + f["pp_law"]
    else:
        total += f["stanley_law"]           # shares the nearest search with the metrics (noise-free)
    total += f["pid"] if accel == "pid" else f["lqg"]
    total += f["kinematic"] if model == "kinematic" else f["dynamic"]
    return total


def bytes_per_episode():
    """params 64 B + init 48 B in; metrics 48 B + status 40 B out."""
    return 64 + 48 + 48 + 40
