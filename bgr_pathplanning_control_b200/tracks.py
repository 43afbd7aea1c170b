"""Synthetic centrelines and cone layouts (the fixed inputs of SURVEY.md section 8d).

The reference gets its centreline from fsd-path-planning as ``[s, x, y, curvature]``
(nodes/planner.py:68,88-89); that planner stays on the host and is out of scope, so the benchmark and
the parity tests drive the engine with these generated tracks instead.  Pure numpy, float64.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class TrackArrays:
    """Base-lap centreline ``x, y, kappa`` (float64), whether it closes on itself, and cones (C, 2)."""
    x: np.ndarray
    y: np.ndarray
    kappa: np.ndarray
    closed: bool = True
    cones: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    name: str = "track"

    @property
    def n(self):
        return len(self.x)

    def start_pose(self, index=0):
        """(x, y, yaw) on waypoint ``index`` with the tangent heading."""
        j = (index + 1) % self.n if self.closed else min(index + 1, self.n - 1)
        i = index if j != index else index - 1
        return float(self.x[index]), float(self.y[index]), float(np.arctan2(self.y[j] - self.y[i], self.x[j] - self.x[i]))


def stadium_oval(n=1024, straight=60.0, radius=15.0, cone_offset=1.5, cone_spacing=5.0):
    """Two ``straight`` m straights joined by two semicircles of ``radius`` m, counter-clockwise, uniformly
    resampled in arc length to ``n`` points, starting at the middle of the bottom straight (config 1)."""
    L = 2.0 * straight + 2.0 * np.pi * radius
    s = np.arange(n) * (L / n)
    x, y, kappa, nx, ny = _stadium_at(s, straight, radius)
    # cones: left/right at +-cone_offset every cone_spacing m of arc, + 4 big orange at the start line
    sc = np.arange(0.0, L, cone_spacing)
    cx, cy, _, cnx, cny = _stadium_at(sc, straight, radius)
    left = np.stack([cx + cone_offset * cnx, cy + cone_offset * cny], axis=1)
    right = np.stack([cx - cone_offset * cnx, cy - cone_offset * cny], axis=1)
    orange = np.array([[0.3, cone_offset + 0.3], [-0.3, cone_offset + 0.3], [0.3, -cone_offset - 0.3],
                       [-0.3, -cone_offset - 0.3]])
    return TrackArrays(x, y, kappa, True, np.concatenate([left, right, orange]), "stadium_oval")


def _stadium_at(s, straight, radius):
    """Point, curvature and left normal at arc length ``s`` (start: middle of the bottom straight, CCW)."""
    h = straight / 2.0
    a = np.pi * radius
    L = 2.0 * straight + 2.0 * a
    s = np.mod(s, L)
    x = np.empty_like(s)
    y = np.empty_like(s)
    k = np.zeros_like(s)
    th = np.empty_like(s)  # tangent heading
    m = s < h                                        # bottom straight, second half
    x[m], y[m], th[m] = s[m], 0.0, 0.0
    m = (s >= h) & (s < h + a)                       # right semicircle, centre (h, R)
    ang = (s[m] - h) / radius
    x[m], y[m], th[m], k[m] = h + radius * np.sin(ang), radius - radius * np.cos(ang), ang, 1.0 / radius
    m = (s >= h + a) & (s < h + a + straight)        # top straight, heading pi
    x[m], y[m], th[m] = h - (s[m] - h - a), 2.0 * radius, np.pi
    m = (s >= h + a + straight) & (s < h + 2 * a + straight)   # left semicircle, centre (-h, R)
    ang = (s[m] - h - a - straight) / radius
    x[m], y[m], th[m], k[m] = -h - radius * np.sin(ang), radius + radius * np.cos(ang), np.pi + ang, 1.0 / radius
    m = s >= h + 2 * a + straight                    # bottom straight, first half
    x[m], y[m], th[m] = -h + (s[m] - h - 2 * a - straight), 0.0, 0.0
    return x, y, k, -np.sin(th), np.cos(th)


def ellipse(n=720, a=40.0, b=20.0):
    """The ellipse of the starter known-answer vectors; curvature analytic."""
    th = np.linspace(0.0, 2.0 * np.pi, n, endpoint=False)
    x, y = a * np.cos(th), b * np.sin(th)
    kappa = a * b / (a * a * np.sin(th) ** 2 + b * b * np.cos(th) ** 2) ** 1.5
    return TrackArrays(x, y, kappa, True, np.zeros((0, 2)), f"ellipse{n}")


def skidpad(n=2048, radius=9.125, entry=15.0, inner_cones=16, outer_cones=13):
    """Formula-Student skidpad as ONE unrolled open path (config 5): ``entry`` m straight, two laps of the
    right circle (clockwise), two laps of the left circle (counter-clockwise), ``entry`` m exit; circle
    centres 2 * radius apart.  The path crosses itself at the origin on every lap, which is exactly the
    case the nearest-waypoint exactness guard has to catch."""
    lap = 2.0 * np.pi * radius
    L = 2.0 * entry + 4.0 * lap
    s = np.arange(n) * (L / (n - 1))
    x = np.empty(n)
    y = np.empty(n)
    k = np.zeros(n)
    m = s < entry
    x[m], y[m] = s[m] - entry, 0.0
    m = (s >= entry) & (s < entry + 2 * lap)          # right circle, centre (0, -radius), clockwise
    ang = (s[m] - entry) / radius
    x[m], y[m], k[m] = radius * np.sin(ang), -radius + radius * np.cos(ang), -1.0 / radius
    m = (s >= entry + 2 * lap) & (s < entry + 4 * lap)  # left circle, centre (0, +radius), counter-clockwise
    ang = (s[m] - entry - 2 * lap) / radius
    x[m], y[m], k[m] = radius * np.sin(ang), radius - radius * np.cos(ang), 1.0 / radius
    m = s >= entry + 4 * lap
    x[m], y[m] = s[m] - entry - 4 * lap, 0.0
    cones = []
    for cy in (-radius, radius):
        for cnt, r in ((inner_cones, radius - 1.5), (outer_cones, radius + 1.5)):
            # outer ring: leave the gap where the two circles meet (FS-layout-like)
            ang = np.linspace(0, 2 * np.pi, cnt, endpoint=False) + (np.pi / cnt)
            px, py = r * np.cos(ang), cy + r * np.sin(ang)
            keep = np.hypot(px, py - (-cy)) > radius + 1.2 if r > radius else np.ones(cnt, bool)
            cones.append(np.stack([px[keep], py[keep]], axis=1))
    return TrackArrays(x, y, k, False, np.concatenate(cones), "skidpad")
