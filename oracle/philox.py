"""Philox4x32-10 counter-based RNG, numpy restatement.  TEST INFRASTRUCTURE (oracle/__init__.py).

ORACLE-DEFINED (the reference has no RNG on this path).  The CUDA
kernels restate the same integer stream, so observation noise is identical per
(seed, episode id, step) on any number of GPUs.

Algorithm: Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32
with 10 rounds, multipliers 0xD2511F53 / 0xCD9E8D57, Weyl keys 0x9E3779B9 / 0xBB67AE85.
Known-answer vectors from Random123's kat_vectors are checked in tests/test_philox.py.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """All arguments broadcastable uint32 arrays.  Returns four uint32 arrays."""
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(
        *(np.asarray(a, dtype=np.uint32) for a in (c0, c1, c2, c3, k0, k1)))
    c0, c1, c2, c3 = (a.astype(np.uint64) for a in (c0, c1, c2, c3))
    k0 = k0.astype(np.uint32).copy()
    k1 = k1.astype(np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0
            p1 = M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return tuple(a.astype(np.uint32) for a in (c0, c1, c2, c3))


def u01_24(r):
    """uint32 -> uniform in (0, 1): ((r >> 8) + 0.5) * 2^-24, exact in fp32 and fp64."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float64) + 0.5) * (1.0 / 16777216.0)


def gaussians4(c0, c1, c2, c3, k0, k1):
    """Four standard normals from one Philox block via two Box-Muller pairs:
    (z0, z1) from words (0, 1), (z2, z3) from words (2, 3);
    z_even = sqrt(-2 ln u_a) cos(2 pi u_b), z_odd = sqrt(-2 ln u_a) sin(2 pi u_b)."""
    r0, r1, r2, r3 = philox4x32_10(c0, c1, c2, c3, k0, k1)
    ua, ub, uc, ud = u01_24(r0), u01_24(r1), u01_24(r2), u01_24(r3)
    ra = np.sqrt(-2.0 * np.log(ua))
    rc = np.sqrt(-2.0 * np.log(uc))
    return (ra * np.cos(2.0 * np.pi * ub), ra * np.sin(2.0 * np.pi * ub),
            rc * np.cos(2.0 * np.pi * ud), rc * np.sin(2.0 * np.pi * ud))


# Stream ids (counter word 1)
STREAM_OBS = 0      # counter = (step, STREAM_OBS, 0, 0): noise on observed (x, y, yaw, v)
STREAM_CONE = 1     # counter = (cone index, STREAM_CONE, 0, 0): (dx, dy) of that cone, words 0/1
