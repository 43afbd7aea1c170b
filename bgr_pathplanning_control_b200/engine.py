"""Batched entry points: (episodes x params) tensors in, per-episode lap metrics out.

This is the "batched entry point taking (episodes x params) tensors" of the north star.  Everything
numeric happens in the sm_100a kernels behind the C ABI (``_abi.py`` -> ``libbgr_rollout.so``); torch is
used for device memory and streams only.  No CPU fallback: without the library or without a CUDA device
these functions raise.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass, field

import numpy as np

from . import _abi as abi
from .vehicle_config import conf

_STEER = {"pure_pursuit": abi.STEER_PURE_PURSUIT, "stanley": abi.STEER_STANLEY}
_ACCEL = {"pid": abi.ACCEL_PID, "lqg": abi.ACCEL_LQG}
_MODEL = {"kinematic": abi.MODEL_KINEMATIC, "dynamic": abi.MODEL_DYNAMIC}
_PRECISION = {"fp32": abi.PRECISION_FP32, "fp64": abi.PRECISION_FP64}


def _ptr(a):
    """Raw address of a numpy array / torch tensor / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _np(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


class Track:
    """Device-resident centreline handle (``bgr_track_t``): what the reference hands its controllers every
    call as ``path`` (N,3) + ``curve`` (nodes/controller.py:56,62), uploaded once."""

    def __init__(self, x, y, kappa=None, closed=True, cones=None, cone_reach=None, device=0):
        lib = abi.load()
        self.x = _np(x, np.float64)
        self.y = _np(y, np.float64, self.x.shape)
        self.kappa = np.zeros_like(self.x) if kappa is None else _np(kappa, np.float64, self.x.shape)
        self.cones = np.zeros((0, 2)) if cones is None else _np(cones, np.float64).reshape(-1, 2)
        self.closed = bool(closed)
        self.device = int(device)
        self.cone_reach = float(cone_reach) if cone_reach is not None else 4.0
        desc = abi.TrackDesc()
        dp = C.POINTER(C.c_double)
        desc.h_x = self.x.ctypes.data_as(dp)
        desc.h_y = self.y.ctypes.data_as(dp)
        desc.h_kappa = self.kappa.ctypes.data_as(dp)
        desc.n = len(self.x)
        desc.closed = int(self.closed)
        desc.h_cones_xy = self.cones.ctypes.data_as(dp) if len(self.cones) else None
        desc.n_cones = len(self.cones)
        desc.cone_reach = self.cone_reach
        self._h = C.c_void_p()
        self._lib = lib
        abi.check(lib.bgr_track_create(C.byref(desc), self.device, C.byref(self._h)))

    @classmethod
    def from_arrays(cls, t, device=0, cone_reach=None):
        """From a ``tracks.TrackArrays``."""
        return cls(t.x, t.y, t.kappa, t.closed, t.cones, cone_reach, device)

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("track handle is closed")
        return self._h

    @property
    def n(self):
        return len(self.x)

    def info(self):
        out = abi.TrackInfo()
        abi.check(self._lib.bgr_track_info(self.handle, C.byref(out)))
        return {k: getattr(out, k) for k, _ in abi.TrackInfo._fields_}

    def tables(self):
        """(path_yaw, guard_radius): per-waypoint heading (core/controllers.py:361-368) and the exactness
        guard radius of the local nearest search."""
        yaw = np.empty(self.n)
        guard = np.empty(self.n)
        abi.check(self._lib.bgr_track_tables(self.handle, yaw.ctypes.data, guard.ctypes.data))
        return yaw, guard

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bgr_track_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class RolloutConfig:
    """Mirror of ``bgr_rollout_cfg_t``; defaults are the reference's ``Vehicle_config`` values."""
    steer: str = "pure_pursuit"      # PurePursuitController | StanleyController (live def)
    accel: str = "pid"               # AccelerationPIDController | LQGAccelerationController
    model: str = "kinematic"         # oracle-defined kinematic | DynamicBicycleModel
    precision: str = "fp32"          # "fp64" = audit build, reference operation order
    max_steps: int = 2000
    laps: int = 1                    # PP: base lap tiled this many times (monotone target index)
    laps_target: int = 0
    audit: bool = False              # fill the near-tie audit channel (full kernel build)
    dynamic_schedule: bool = False   # full build, > 1 wave of episodes: lanes pull their next episode from a queue
    keep_order: bool = False         # launch in row order (default: rollouts that can end early launch slowest-first)
    dt: float = conf.dt
    wheelbase: float = conf.WB
    l_f: float = conf.l_f
    l_r: float = conf.l_r
    max_steer: float = conf.MAX_STEER
    max_accel: float = conf.MAX_ACCEL
    max_decel: float = conf.MAX_DECEL
    hit_radius: float = 0.0
    tie_eps: float = 0.0
    sigma_xy: float = 0.0
    sigma_yaw: float = 0.0
    sigma_v: float = 0.0
    sigma_cone: float = 0.0
    noise_seed: int = 0
    episode_id_base: int = 0
    split_steps: int = 0             # > 0: two-phase execution (short first launch, survivors re-run packed)
    stanley_shadowed: bool = False   # the dead-code Stanley variant (core/controllers.py:234-324); labelled extra
    stanley_max_steer_rate: float = 0.0   # 0 = reference default deg2rad(30) rad/s
    stanley_alpha: float = 0.0            # 0 = reference default 0.5
    stanley_lookahead: float = 0.0        # shadowed variant's optional look-ahead advance (:276-281); 0 = off
    # receding-horizon re-planning replay (nodes/planner.py:26,62,72-77; pure pursuit only): the controller sees a
    # local path of `replan_horizon` points (every `replan_stride`-th centreline waypoint from the one nearest to the
    # car), re-planned whenever the planner's own look-ahead index has passed `replan_after`.  0 = off.
    replan_horizon: int = 0
    replan_stride: int = 1
    replan_after: int = 3                 # nodes/planner.py:26
    planner_lookahead: float = conf.LOOKAHEAD_DISTANCE   # nodes/planner.py:75

    def to_struct(self):
        c = abi.RolloutCfg()
        try:
            c.steer_kind, c.accel_kind = _STEER[self.steer], _ACCEL[self.accel]
            c.model_kind, c.precision = _MODEL[self.model], _PRECISION[self.precision]
        except KeyError as e:
            raise ValueError(f"Unknown controller / model / precision: {e.args[0]}") from None
        c.max_steps, c.laps, c.laps_target = int(self.max_steps), int(self.laps), int(self.laps_target)
        c.flags = ((abi.CFG_AUDIT if self.audit else 0) | (abi.CFG_STANLEY_SHADOWED if self.stanley_shadowed else 0)
                   | (abi.CFG_DYNAMIC_SCHEDULE if self.dynamic_schedule else 0)
                   | (abi.CFG_KEEP_ORDER if self.keep_order else 0))
        c.stanley_max_steer_rate, c.stanley_alpha = float(self.stanley_max_steer_rate), float(self.stanley_alpha)
        c.stanley_lookahead = float(self.stanley_lookahead)
        for k in ("dt", "wheelbase", "l_f", "l_r", "max_steer", "max_accel", "max_decel", "hit_radius", "tie_eps",
                  "sigma_xy", "sigma_yaw", "sigma_v", "sigma_cone"):
            setattr(c, k, float(getattr(self, k)))
        c.noise_seed = int(self.noise_seed) & 0xFFFFFFFF
        c.split_steps = max(0, int(self.split_steps))
        c.episode_id_base = int(self.episode_id_base)
        c.replan_horizon, c.replan_stride = int(self.replan_horizon), int(self.replan_stride)
        c.replan_after, c.planner_lookahead = int(self.replan_after), float(self.planner_lookahead)
        return c


def default_params(n_episodes=1, dtype=np.float32):
    """(E, 16) parameter rows filled with the reference defaults (vehicle_config.py:4-39; Stanley gains
    from the factory, core/controllers.py:585)."""
    p = np.zeros((n_episodes, abi.N_PARAMS), dtype=dtype)
    p[:, abi.P_LF] = conf.LOOKAHEAD_DISTANCE
    p[:, abi.P_KP], p[:, abi.P_KI], p[:, abi.P_KD] = conf.kp_accel, conf.ki_accel, conf.kd_accel
    p[:, abi.P_TARGET_SPEED] = conf.TARGET_SPEED
    p[:, abi.P_K_EXPO] = conf.k_expo
    p[:, abi.P_K_STANLEY], p[:, abi.P_K_SOFT] = 1.0, 0.1
    p[:, abi.P_MU], p[:, abi.P_C_F], p[:, abi.P_C_R] = conf.mu, conf.c_f, conf.c_r
    p[:, abi.P_MASS], p[:, abi.P_I_Z] = conf.m, conf.I_z
    return p


def initial_states(track, n_episodes=1, index=0, lateral_offset=None, yaw_offset=None, v0=0.0):
    """(E, 6) float64 initial rows ``[x, y, yaw, v, r, beta]`` on waypoint ``index`` of a
    ``tracks.TrackArrays`` / ``Track``, heading along the tangent, optionally shifted sideways."""
    n = len(track.x)
    j = (index + 1) % n
    yaw = float(np.arctan2(track.y[j] - track.y[index], track.x[j] - track.x[index]))
    s = np.zeros((n_episodes, abi.N_STATE))
    off = np.zeros(n_episodes) if lateral_offset is None else np.asarray(lateral_offset, dtype=np.float64)
    s[:, abi.X] = track.x[index] - np.sin(yaw) * off
    s[:, abi.Y] = track.y[index] + np.cos(yaw) * off
    s[:, abi.YAW] = yaw + (0.0 if yaw_offset is None else np.asarray(yaw_offset, dtype=np.float64))
    s[:, abi.V] = v0
    return s


@dataclass
class RolloutResult:
    """Per-episode outputs.  ``metrics`` (E, 12) float32 and ``status`` (E, 10) int32 follow the
    ``BGR_M_*`` / ``BGR_S_*`` columns of the header; tensors live where the inputs lived."""
    metrics: object
    status: object
    traj: object = None
    aggregate: object = None
    kernel_ms: float | None = None

    def _col(self, a, i):
        return a[:, i]

    # the five scalars of preformance_analysis/metrics_calculator.py:1-8, per episode
    @property
    def lateral_error(self):
        return self._col(self.metrics, abi.M_LAT_MEAN), self._col(self.metrics, abi.M_LAT_STD)

    @property
    def heading_error(self):
        return self._col(self.metrics, abi.M_HEAD_MEAN), self._col(self.metrics, abi.M_HEAD_STD)

    @property
    def lap_time(self):
        return self._col(self.metrics, abi.M_LAP_TIME)

    @property
    def cone_hits(self):
        return self._col(self.status, abi.S_CONE_HITS)

    @property
    def steps(self):
        return self._col(self.status, abi.S_STEPS)

    @property
    def flags(self):
        return self._col(self.status, abi.S_FLAGS)

    def final_state(self):
        return self.metrics[:, abi.M_X:abi.M_BETA + 1]


def _torch():
    import torch
    return torch


def rollout(track, config, params, init_state, *, traj=False, aggregate=False, out=None, time_kernel=False):
    """Run ``E`` closed-loop episodes on the track's GPU.

    ``params`` (E, 16) float32 and ``init_state`` (E, 6) float64 are CUDA torch tensors on the track's
    device (the zero-copy path) -- or numpy arrays, in which case the host-buffer entry point
    ``bgr_rollout_host`` does the H2D/D2H copies and numpy arrays come back.
    Asynchronous on the current torch stream for CUDA tensors.
    """
    lib = abi.load()
    cfg = config.to_struct() if isinstance(config, RolloutConfig) else config
    if isinstance(params, np.ndarray) or isinstance(init_state, np.ndarray):
        return _rollout_numpy(lib, track, cfg, params, init_state, traj, aggregate, out)
    torch = _torch()
    E = int(params.shape[0])
    if not (params.is_cuda and init_state.is_cuda):
        raise ValueError("params and init_state must be CUDA tensors (or numpy arrays for the host entry point); "
                         "there is no CPU path")
    if params.device.index != track.device or init_state.device.index != track.device:
        raise ValueError(f"inputs must live on the track's device cuda:{track.device}")
    if params.dtype != torch.float32 or tuple(params.shape) != (E, abi.N_PARAMS) or not params.is_contiguous():
        raise ValueError(f"params must be a contiguous float32 tensor of shape (E, {abi.N_PARAMS})")
    if (init_state.dtype != torch.float64 or tuple(init_state.shape) != (E, abi.N_STATE)
            or not init_state.is_contiguous()):
        raise ValueError(f"init_state must be a contiguous float64 tensor of shape (E, {abi.N_STATE})")
    dev = params.device
    if out is not None:
        metrics, status = out.metrics, out.status
    else:
        metrics = torch.empty((E, abi.N_METRICS), dtype=torch.float32, device=dev)
        status = torch.empty((E, abi.N_STATUS), dtype=torch.int32, device=dev)
    tr = None
    if traj:
        tr = torch.zeros((E, cfg.max_steps, abi.N_TRAJ), dtype=torch.float64, device=dev)
    agg = None
    if aggregate:
        agg = out.aggregate if (out is not None and out.aggregate is not None) else torch.empty(
            abi.N_AGG, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    abi.check(lib.bgr_rollout(track.handle, C.byref(cfg), E, _ptr(params), _ptr(init_state), _ptr(metrics),
                              _ptr(status), _ptr(tr), _ptr(agg), stream))
    ms = None
    if time_kernel and E > 0:
        v = C.c_double()
        abi.check(lib.bgr_last_kernel_ms(C.byref(v)))
        ms = v.value
    return RolloutResult(metrics, status, tr, agg, ms)


_group_streams = {}
_group_pool = None


_worker_local = threading.local()


def _group_worker(track, config, params, init_state, aggregate, out):
    """One group of a host-buffer ``rollout_groups`` call, on a pool thread.  The call is ordered on a stream that
    belongs to the thread: on the legacy default stream (``stream = NULL``) the completion wait of one thread's call
    would sit in front of the next thread's first upload, and the groups would run one after the other."""
    lib = abi.load()
    streams = getattr(_worker_local, "streams", None)
    if streams is None:
        streams = _worker_local.streams = {}
    st = streams.get(track.device)
    if st is None:
        st = C.c_void_p()
        abi.check(lib.bgr_stream_create(int(track.device), C.byref(st)))
        streams[track.device] = st
    cfg = config.to_struct() if isinstance(config, RolloutConfig) else config
    return _rollout_numpy(lib, track, cfg, params, init_state, False, aggregate, out, stream=st)


def _host_group_pool(n):
    global _group_pool
    if _group_pool is None or _group_pool._max_workers < n:
        from concurrent.futures import ThreadPoolExecutor
        if _group_pool is not None:
            _group_pool.shutdown(wait=True)
        _group_pool = ThreadPoolExecutor(max_workers=max(n, 4), thread_name_prefix="bgr-group")
    return _group_pool


def rollout_groups(track, groups, *, aggregate=False, outs=None, order=None):
    """Mixed-controller batches (BASELINE config 5): one homogeneous ``bgr_rollout`` per ``(config, params, init)``
    group, the groups launched on CONCURRENT streams.  A group whose episodes end at very different times (noisy
    Stanley on the skidpad: mean 130 steps, a few at the 1 500-step cap) spends most of its kernel time with a handful
    of warps alive -- time another group's full-occupancy launch can fill.  ``order`` = launch order (default: as
    given; put the latency-bound groups first).  CUDA tensors: one ``RolloutResult`` per group, complete on the
    CURRENT stream (the side streams are joined back into it).  numpy arrays: one host thread per group through
    ``bgr_rollout_host`` (ctypes drops the GIL; the library's copy / compute pipeline is per thread), synchronous."""
    seq = list(order) if order is not None else list(range(len(groups)))
    if any(isinstance(g[1], np.ndarray) or isinstance(g[2], np.ndarray) for g in groups):
        # worker threads PERSIST across calls: the library keeps its host-buffer pipeline (streams, device buffer ring)
        # per host thread, and a fresh thread would have to build it again
        pool = _host_group_pool(len(groups))
        futs = {g: pool.submit(_group_worker, track, groups[g][0], groups[g][1], groups[g][2], aggregate,
                               None if outs is None else outs[g]) for g in seq}
        return [futs[g].result() for g in range(len(groups))]
    torch = _torch()
    dev = torch.device("cuda", track.device)
    main = torch.cuda.current_stream(dev)
    streams = _group_streams.setdefault(track.device, [])
    while len(streams) < len(groups):
        streams.append(torch.cuda.Stream(dev))
    # Result tensors are allocated HERE, on the caller's stream: the caching allocator ties a block to the stream that
    # was current when it was allocated, and a block born on a side stream could be handed out again for the next call's
    # side-stream work while the caller's stream is still reading it.
    if outs is None:
        outs = []
        for cfg, params, _ in groups:
            E = int(params.shape[0])
            outs.append(RolloutResult(torch.empty((E, abi.N_METRICS), dtype=torch.float32, device=dev),
                                      torch.empty((E, abi.N_STATUS), dtype=torch.int32, device=dev), None,
                                      torch.empty(abi.N_AGG, dtype=torch.float64, device=dev) if aggregate else None))
    start = torch.cuda.Event()
    start.record(main)
    results = [None] * len(groups)
    for g in seq:
        cfg, params, init = groups[g]
        streams[g].wait_event(start)
        with torch.cuda.stream(streams[g]):
            results[g] = rollout(track, cfg, params, init, aggregate=aggregate, out=outs[g])
        # the inputs are the caller's too: tell the allocator the side stream is using them
        params.record_stream(streams[g])
        init.record_stream(streams[g])
    for st in streams[: len(groups)]:
        main.wait_stream(st)
    return results


def release_group_streams():
    """Drop the cached side streams of ``rollout_groups`` (one list per device, grown to the largest group count seen)."""
    _group_streams.clear()


def _rollout_numpy(lib, track, cfg, params, init_state, traj, aggregate, out=None, stream=None):
    params = _np(params, np.float32)
    E = params.shape[0]
    params = _np(params, np.float32, (E, abi.N_PARAMS))
    init_state = _np(init_state, np.float64, (E, abi.N_STATE))
    if out is not None:      # caller-owned (e.g. pinned) result buffers
        metrics, status = out.metrics, out.status
        if (metrics.dtype != np.float32 or metrics.shape != (E, abi.N_METRICS) or not metrics.flags.c_contiguous
                or status.dtype != np.int32 or status.shape != (E, abi.N_STATUS) or not status.flags.c_contiguous):
            raise ValueError("out.metrics / out.status must be contiguous (E, 12) float32 / (E, 10) int32 arrays")
    else:
        metrics = np.empty((E, abi.N_METRICS), dtype=np.float32)
        status = np.empty((E, abi.N_STATUS), dtype=np.int32)
    tr = np.zeros((E, cfg.max_steps, abi.N_TRAJ)) if traj else None
    agg = np.empty(abi.N_AGG) if aggregate else None
    abi.check(lib.bgr_rollout_host(track.handle, C.byref(cfg), E, _ptr(params), _ptr(init_state), _ptr(metrics),
                                   _ptr(status), _ptr(tr), _ptr(agg), stream))
    return RolloutResult(metrics, status, tr, agg, None)


_async_local = threading.local()   # per host thread, like the library's pipeline: device -> [cudaStream_t, ...]
_ASYNC_DEPTH = 4


def _async_stream(lib, device):
    """The completion stream of the calling thread's next in-flight host-buffer rollout (round robin over
    ``_ASYNC_DEPTH`` streams per thread and device: each in-flight call completes on its own stream)."""
    pools = getattr(_async_local, "pools", None)
    if pools is None:
        pools = _async_local.pools = {}
        _async_local.next = {}
    pool = pools.setdefault(device, [])
    if len(pool) < _ASYNC_DEPTH:
        h = C.c_void_p()
        abi.check(lib.bgr_stream_create(int(device), C.byref(h)))
        pool.append(h)
        return pool[-1]
    i = _async_local.next.get(device, 0)
    _async_local.next[device] = (i + 1) % _ASYNC_DEPTH
    return pool[i]


class PendingRollout:
    """A host-buffer rollout that has been queued but not waited for (``rollout_async``).  ``wait()`` blocks until the
    result arrays are filled and returns the ``RolloutResult``."""

    def __init__(self, lib, stream, result, keep):
        self._lib, self._stream, self._result, self._keep, self._done = lib, stream, result, keep, False

    def wait(self):
        if not self._done:
            abi.check(self._lib.bgr_stream_synchronize(self._stream))
            self._done = True
            self._keep = None
        return self._result


def rollout_async(track, config, params, init_state, *, aggregate=False, out=None):
    """``rollout`` for HOST (numpy, ideally pinned) buffers without the final synchronise (``bgr_rollout_host_async``):
    returns a ``PendingRollout``.  Calling it for batch i + 1 before ``wait()``-ing batch i lets the library run the
    first upload of the new batch under the last kernel and download of the old one -- the host-buffer path then costs
    the kernel time, not kernel + copies.  The input and output arrays must not be touched until ``wait()`` returns;
    give every in-flight batch its own ``out`` buffers, and keep at most 4 batches in flight per device."""
    lib = abi.load()
    cfg = config.to_struct() if isinstance(config, RolloutConfig) else config
    params = _np(params, np.float32)
    E = params.shape[0]
    params = _np(params, np.float32, (E, abi.N_PARAMS))
    init_state = _np(init_state, np.float64, (E, abi.N_STATE))
    if out is not None:
        metrics, status = out.metrics, out.status
        if (metrics.dtype != np.float32 or metrics.shape != (E, abi.N_METRICS) or not metrics.flags.c_contiguous
                or status.dtype != np.int32 or status.shape != (E, abi.N_STATUS) or not status.flags.c_contiguous):
            raise ValueError("out.metrics / out.status must be contiguous (E, 12) float32 / (E, 10) int32 arrays")
    else:
        metrics = np.empty((E, abi.N_METRICS), dtype=np.float32)
        status = np.empty((E, abi.N_STATUS), dtype=np.int32)
    agg = None
    if aggregate:
        agg = out.aggregate if (out is not None and out.aggregate is not None) else np.empty(abi.N_AGG)
    stream = _async_stream(lib, track.device)
    abi.check(lib.bgr_rollout_host_async(track.handle, C.byref(cfg), E, _ptr(params), _ptr(init_state), _ptr(metrics),
                                         _ptr(status), _ptr(agg), stream))
    return PendingRollout(lib, stream, RolloutResult(metrics, status, None, agg, None), (params, init_state, cfg, track))


_group_lanes = []      # one single-thread executor per group index: group g always runs on ITS thread (and pipeline)


class PendingGroups:
    """The groups of one ``rollout_groups_async`` call.  ``wait()`` returns one ``RolloutResult`` per group (in the
    order the groups were given) once every group's result arrays are filled."""

    def __init__(self, futures):
        self._futures, self._results, self._error = futures, None, None

    def wait(self):
        if self._futures is not None:
            # every group is waited for even when one of them failed: the caller's arrays must not be in flight any
            # more when the error surfaces (the first error is raised once all groups have settled, and again on
            # every later wait())
            results = []
            for f in self._futures:
                try:
                    results.append(f.result().wait())
                except Exception as exc:      # noqa: BLE001
                    results.append(None)
                    self._error = self._error or exc
            self._futures, self._results = None, results
        if self._error is not None:
            raise self._error
        return self._results


def rollout_groups_async(track, groups, *, aggregate=False, outs=None, order=None):
    """``rollout_groups`` for HOST buffers without the final synchronise: every group is QUEUED
    (``bgr_rollout_host_async``) from a host thread of its own -- group g always from the same thread, so it always
    meets the same copy / compute pipeline and device buffer ring inside the library -- and a ``PendingGroups`` is
    returned.  Queue batch i + 1 before ``wait()``-ing batch i: a Monte Carlo batch's uploads then run under the
    previous batch's kernels and its downloads under the next batch's, and the few long episodes at the end of one
    batch's heavy-tailed groups no longer hold the machine alone.  Same rules as ``rollout_async``: do not touch the
    arrays of a batch before its ``wait()`` returns, give every in-flight batch its own ``outs``, at most 3 batches in
    flight (the library's ring of device buffer slots per host thread)."""
    from concurrent.futures import ThreadPoolExecutor
    while len(_group_lanes) < len(groups):
        _group_lanes.append(ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"bgr-lane{len(_group_lanes)}"))
    seq = list(order) if order is not None else list(range(len(groups)))
    futs = {}
    for g in seq:
        cfg, params, init = groups[g]
        futs[g] = _group_lanes[g].submit(rollout_async, track, cfg, params, init, aggregate=aggregate,
                                         out=None if outs is None else outs[g])
    return PendingGroups([futs[g] for g in range(len(groups))])


def build_id():
    """Identity of the loaded library build (sources + compile-time knobs)."""
    return abi.load().bgr_build_id().decode()


def lqg_covariance(dt=conf.dt, n_steps=0):
    """(2, 2) estimate covariance ``P`` of the LQG controller's Kalman filter after ``n_steps`` updates
    (core/controllers.py:109, :145, :160)."""
    P = np.empty(4)
    abi.check(abi.load().bgr_lqg_covariance(float(dt), int(n_steps), _ptr(P)))
    return P.reshape(2, 2)


def last_kernel_ms():
    v = C.c_double()
    abi.check(abi.load().bgr_last_kernel_ms(C.byref(v)))
    return v.value


def launch_count():
    return int(abi.load().bgr_launch_count())


def lqg_design(dt=conf.dt, n_steps=0):
    """(K (2,), kalman_gains (n_steps, 2)): the LQR gain of ``control.dlqr`` (core/controllers.py:97) and
    the data-independent Kalman gain sequence (:145-160) every episode shares."""
    K = np.empty(2)
    g = np.empty((n_steps, 2)) if n_steps > 0 else None
    abi.check(abi.load().bgr_lqg_design(float(dt), int(n_steps), _ptr(K), _ptr(g)))
    return K, g


def fp32_peak_probe(device=0):
    """(TFLOP/s, G lane-FMA/s) of an FFMA-chain microbenchmark: the roofline denominator."""
    t, g = C.c_double(), C.c_double()
    abi.check(abi.load().bgr_fp32_peak_probe(int(device), C.byref(t), C.byref(g)))
    return t.value, g.value


@dataclass
class MpcConfig:
    """Mirror of ``bgr_mpc_cfg_t``; defaults = MPCController as the factory builds it
    (core/controllers.py:483-484, :587)."""
    horizon: int = 10
    dt: float = 0.05
    precision: str = "fp32"
    wheelbase: float = conf.WB
    target_speed: float = conf.TARGET_SPEED
    q: tuple = (1.0, 1.0, 0.5, 0.1)
    r: tuple = (0.1, 0.1)
    max_accel: float = conf.MAX_ACCEL
    max_decel: float = conf.MAX_DECEL
    max_steer: float = conf.MAX_STEER
    explicit_rollout: bool = False   # fp32 stage-by-stage rollout instead of the factored evaluator

    def to_struct(self):
        c = abi.MpcCfg()
        c.flags = abi.MPC_EXPLICIT_ROLLOUT if self.explicit_rollout else 0
        c.horizon, c.precision = int(self.horizon), _PRECISION[self.precision]
        c.dt, c.wheelbase, c.target_speed = float(self.dt), float(self.wheelbase), float(self.target_speed)
        for i in range(4):
            c.q[i] = float(self.q[i])
        for i in range(2):
            c.r[i] = float(self.r[i])
        c.max_accel, c.max_decel, c.max_steer = float(self.max_accel), float(self.max_decel), float(self.max_steer)
        return c


@dataclass
class MpcResult:
    best_u: object        # (E, 2) float32 = (acceleration, steering) as compute_control returns them (:568)
    best_cost: object     # (E,) float32, +inf when no candidate is feasible
    best_idx: object      # (E,) int32, -1 when none is feasible
    costs: object = None  # (E, K) float32 (+inf = infeasible), only on request
    kernel_ms: float | None = None


def candidate_bank(n_candidates=4096, horizon=30, seed=20261021, max_accel=conf.MAX_ACCEL, max_steer=conf.MAX_STEER,
                   kind="uniform", dt=0.05):
    """(K, H, 2) float32 = (accel, steer) candidate control sequences.  Host numpy; the same bank feeds the oracle.

    ``kind="uniform"``: uniform draws smoothed by a 3-tap moving average, candidate 0 all zeros (SURVEY.md section 8d
    config 4) -- a generic sample of the control box.  Measured against the exact optimum of the reference's QP
    (oracle ``mpc_qp_optimum``) its best of 4 096 candidates costs 19 % more in the median (tests/test_gpu_mpc.py).

    ``kind="qp_family"``: the bank the drop-in ``MPCController`` uses.  The reference's QP (core/controllers.py:503-552)
    separates into a steering chain (yaw) and an acceleration chain (v); the optimum of each chain depends on ONE
    scalar besides the horizon -- the initial heading (times v0 / WB) and the initial speed error.  The bank is the
    outer product of sqrt(K) acceleration profiles x sqrt(K) steering profiles, each the exact bounded-least-squares
    optimum of its chain for one grid value of that scalar (64 speed errors in +-8 m/s; 16 headings in +-3.2 rad x
    4 speeds): the cheapest candidate is then within 0.2 % of the QP optimum over config 4's states (median 0.00 %).
    The kernels do not care which bank they score."""
    if kind == "qp_family":
        return _qp_family_bank(n_candidates, horizon, dt, max_accel, max_steer)
    if kind != "uniform":
        raise ValueError(f"Unknown candidate bank kind: {kind}")
    rng = np.random.default_rng(seed)
    u = np.empty((n_candidates, horizon, 2))
    u[..., 0] = rng.uniform(-max_accel, max_accel, (n_candidates, horizon))
    u[..., 1] = rng.uniform(-max_steer, max_steer, (n_candidates, horizon))
    pad = np.concatenate([u[:, :1], u, u[:, -1:]], axis=1)
    u = (pad[:, :-2] + pad[:, 1:-1] + pad[:, 2:]) / 3.0
    u[0] = 0.0
    return np.ascontiguousarray(u, dtype=np.float32)


def _qp_family_bank(n_candidates, horizon, dt, max_accel, max_steer, q=(1.0, 1.0, 0.5, 0.1), r=(0.1, 0.1)):
    from scipy.optimize import lsq_linear       # input synthesis on the host, like tracks.py / sweeps.py
    side = int(round(np.sqrt(n_candidates)))
    if side * side != n_candidates or side < 8 or side % 4:
        raise ValueError("a qp_family bank needs a square number of candidates with sqrt(K) a multiple of 4 (e.g. 4096)")
    H = int(horizon)
    low = np.tril(np.ones((H, H)))

    def chain(scale, qq, rr, e0, bound):
        a_mat = np.vstack([np.sqrt(qq) * scale * low, np.sqrt(rr) * np.eye(H)])
        b_vec = np.concatenate([-np.sqrt(qq) * e0 * np.ones(H), np.zeros(H)])
        return lsq_linear(a_mat, b_vec, bounds=(-bound, bound), method="bvls", tol=1e-14).x

    acc = np.array([chain(dt, q[3], r[0], e, max_accel) for e in np.linspace(-8.0, 8.0, side)])
    n_y = side // 4
    yaws = np.concatenate([-np.geomspace(0.02, 3.2, n_y // 2)[::-1], np.geomspace(0.02, 3.2, n_y - n_y // 2)])
    ste = np.array([chain(v * dt / conf.WB, q[2], r[1], y, max_steer) for v in (1.0, 2.5, 4.5, 7.0) for y in yaws])
    bank = np.empty((side * side, H, 2))
    bank[:, :, 0] = np.repeat(acc, side, axis=0)
    bank[:, :, 1] = np.tile(ste, (side, 1))
    return np.ascontiguousarray(bank, dtype=np.float32)


class MpcBank:
    """Device-resident candidate bank (``bgr_mpc_bank_t``): upload ``bank`` (K, H, 2) float32 once, then pass the
    handle to ``mpc_evaluate`` in place of the array -- per call only the states travel, and the per-candidate
    statistics of the factored evaluator are computed once per cost configuration instead of once per call."""

    def __init__(self, bank, device=0):
        lib = abi.load()
        self.array = _np(bank, np.float32)
        if self.array.ndim != 3 or self.array.shape[2] != 2:
            raise ValueError("bank must have shape (K, H, 2)")
        self.K, self.H = int(self.array.shape[0]), int(self.array.shape[1])
        self.device = int(device)
        self._lib = lib
        self._h = C.c_void_p()
        abi.check(lib.bgr_mpc_bank_create(self.device, _ptr(self.array), self.K, self.H, C.byref(self._h)))

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("bank handle is closed")
        return self._h

    @property
    def shape(self):
        return self.array.shape

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bgr_mpc_bank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fp64_peak_probe(device=0):
    """TFLOP/s of a DFMA-chain microbenchmark: the measured FP64 CUDA-core peak (MPC evaluator roofline)."""
    t = C.c_double()
    abi.check(abi.load().bgr_fp64_peak_probe(int(device), C.byref(t)))
    return t.value


def mpc_evaluate(track, config, state, ref_start, bank, ref_len=None, *, return_costs=False, time_kernel=False, out=None):
    """Score every candidate control sequence of ``bank`` (K, H, 2) -- an array, or a resident ``MpcBank`` -- for every
    episode state (E, 4) = ``[x, y, yaw, v]`` and pick the cheapest feasible one.  CUDA tensors (zero copy, async) or
    numpy arrays (host entry point)."""
    lib = abi.load()
    cfg = config.to_struct() if isinstance(config, MpcConfig) else config
    if isinstance(bank, MpcBank):
        return _mpc_evaluate_bank(lib, track, cfg, state, ref_start, bank, ref_len, return_costs, time_kernel, out)
    if isinstance(state, np.ndarray):
        state = _np(state, np.float64)
        E = state.shape[0]
        state = _np(state, np.float64, (E, 4))
        ref_start = _np(ref_start, np.int32, (E,))
        ref_len = None if ref_len is None else _np(ref_len, np.int32, (E,))
        bank = _np(bank, np.float32)
        K = bank.shape[0]
        bank = _np(bank, np.float32, (K, cfg.horizon, 2))
        u = np.empty((E, 2), np.float32)
        cost = np.empty(E, np.float32)
        idx = np.empty(E, np.int32)
        costs = np.empty((E, K), np.float32) if return_costs else None
        abi.check(lib.bgr_mpc_eval_host(track.handle, C.byref(cfg), E, _ptr(state), _ptr(ref_start), _ptr(ref_len),
                                        _ptr(bank), K, _ptr(u), _ptr(cost), _ptr(idx), _ptr(costs), None))
        return MpcResult(u, cost, idx, costs)
    torch = _torch()
    E, K = int(state.shape[0]), int(bank.shape[0])
    for name, t, dt, shape in (("state", state, torch.float64, (E, 4)), ("ref_start", ref_start, torch.int32, (E,)),
                               ("bank", bank, torch.float32, (K, cfg.horizon, 2))):
        if not t.is_cuda or t.device.index != track.device or t.dtype != dt or tuple(t.shape) != shape \
                or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {dt} CUDA tensor of shape {shape} on cuda:{track.device}")
    dev = state.device
    u = torch.empty((E, 2), dtype=torch.float32, device=dev)
    cost = torch.empty(E, dtype=torch.float32, device=dev)
    idx = torch.empty(E, dtype=torch.int32, device=dev)
    costs = torch.empty((E, K), dtype=torch.float32, device=dev) if return_costs else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    abi.check(lib.bgr_mpc_eval(track.handle, C.byref(cfg), E, _ptr(state), _ptr(ref_start), _ptr(ref_len), _ptr(bank),
                               K, _ptr(u), _ptr(cost), _ptr(idx), _ptr(costs), stream))
    ms = last_kernel_ms() if time_kernel and E > 0 else None
    return MpcResult(u, cost, idx, costs, ms)


def _mpc_evaluate_bank(lib, track, cfg, state, ref_start, bank, ref_len, return_costs, time_kernel, out=None):
    """``out``: an ``MpcResult`` of (E, 2) float32 / (E,) float32 / (E,) int32 host arrays to fill (pinned buffers make
    the copies asynchronous; a streaming caller reuses one result object per control period)."""
    K = bank.K
    if isinstance(state, np.ndarray):
        state = _np(state, np.float64)
        E = state.shape[0]
        state = _np(state, np.float64, (E, 4))
        ref_start = _np(ref_start, np.int32, (E,))
        ref_len = None if ref_len is None else _np(ref_len, np.int32, (E,))
        if out is not None:
            u, cost, idx = out.best_u, out.best_cost, out.best_idx
            if (u.dtype != np.float32 or u.shape != (E, 2) or cost.dtype != np.float32 or cost.shape != (E,)
                    or idx.dtype != np.int32 or idx.shape != (E,)):
                raise ValueError("out must hold (E, 2) float32, (E,) float32 and (E,) int32 arrays")
        else:
            u = np.empty((E, 2), np.float32)
            cost = np.empty(E, np.float32)
            idx = np.empty(E, np.int32)
        costs = np.empty((E, K), np.float32) if return_costs else None
        abi.check(lib.bgr_mpc_eval_bank_host(track.handle, C.byref(cfg), E, _ptr(state), _ptr(ref_start), _ptr(ref_len),
                                             bank.handle, _ptr(u), _ptr(cost), _ptr(idx), _ptr(costs), None))
        return MpcResult(u, cost, idx, costs)
    torch = _torch()
    E = int(state.shape[0])
    for name, t, dt, shape in (("state", state, torch.float64, (E, 4)), ("ref_start", ref_start, torch.int32, (E,))):
        if not t.is_cuda or t.device.index != track.device or t.dtype != dt or tuple(t.shape) != shape \
                or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous {dt} CUDA tensor of shape {shape} on cuda:{track.device}")
    dev = state.device
    u = torch.empty((E, 2), dtype=torch.float32, device=dev)
    cost = torch.empty(E, dtype=torch.float32, device=dev)
    idx = torch.empty(E, dtype=torch.int32, device=dev)
    costs = torch.empty((E, K), dtype=torch.float32, device=dev) if return_costs else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    abi.check(lib.bgr_mpc_eval_bank(track.handle, C.byref(cfg), E, _ptr(state), _ptr(ref_start), _ptr(ref_len),
                                    bank.handle, _ptr(u), _ptr(cost), _ptr(idx), _ptr(costs), stream))
    ms = last_kernel_ms() if time_kernel and E > 0 else None
    return MpcResult(u, cost, idx, costs, ms)


class PendingMpc:
    """A host-buffer MPC evaluation that has been queued but not waited for (``mpc_evaluate_async``)."""

    def __init__(self, lib, stream, result, keep):
        self._lib, self._stream, self._result, self._keep, self._done = lib, stream, result, keep, False

    def wait(self):
        if not self._done:
            abi.check(self._lib.bgr_stream_synchronize(self._stream))
            self._done = True
            self._keep = None
        return self._result


_mpc_async_streams = {}      # device -> two completion streams, alternating (bgr_mpc_eval_bank_host_async)
_mpc_async_next = {}


def mpc_evaluate_async(track, config, state, ref_start, bank, ref_len=None, *, out=None):
    """``mpc_evaluate`` for HOST (numpy, ideally pinned) states and a resident ``MpcBank`` without the final
    synchronise (``bgr_mpc_eval_bank_host_async``): returns a ``PendingMpc``.  Queue batch i + 1 before ``wait()``-ing
    batch i and its upload runs under batch i's kernels (consecutive calls alternate between two streams and two device
    arenas).  The arrays must not be touched until ``wait()`` returns; give each of the (at most two) batches in flight
    its own ``out``."""
    lib = abi.load()
    if not isinstance(bank, MpcBank):
        raise TypeError("mpc_evaluate_async needs a resident MpcBank")
    cfg = config.to_struct() if isinstance(config, MpcConfig) else config
    state = _np(state, np.float64)
    E = state.shape[0]
    state = _np(state, np.float64, (E, 4))
    ref_start = _np(ref_start, np.int32, (E,))
    ref_len = None if ref_len is None else _np(ref_len, np.int32, (E,))
    if out is not None:
        u, cost, idx = out.best_u, out.best_cost, out.best_idx
        if (u.dtype != np.float32 or u.shape != (E, 2) or cost.dtype != np.float32 or cost.shape != (E,)
                or idx.dtype != np.int32 or idx.shape != (E,)):
            raise ValueError("out must hold (E, 2) float32, (E,) float32 and (E,) int32 arrays")
    else:
        u, cost, idx = np.empty((E, 2), np.float32), np.empty(E, np.float32), np.empty(E, np.int32)
    pool = _mpc_async_streams.setdefault(track.device, [])
    while len(pool) < 2:
        h = C.c_void_p()
        abi.check(lib.bgr_stream_create(int(track.device), C.byref(h)))
        pool.append(h)
    i = _mpc_async_next.get(track.device, 0)
    _mpc_async_next[track.device] = i ^ 1
    stream = pool[i]
    abi.check(lib.bgr_mpc_eval_bank_host_async(track.handle, C.byref(cfg), E, _ptr(state), _ptr(ref_start), _ptr(ref_len),
                                               bank.handle, _ptr(u), _ptr(cost), _ptr(idx), stream))
    return PendingMpc(lib, stream, MpcResult(u, cost, idx, None), (state, ref_start, ref_len, bank))


def identify_dynamics(traj, rows=None):
    """Batched ``identify_dynamics`` (old/test_runs_system_id.py:84-111): for every episode of a trajectory dump
    ``traj`` (E, S, 12) -- or a ``RolloutResult`` run with ``traj=True`` -- the least-squares fit
    ``x[k+1] = A x[k] + B u[k] + c`` over its logged rows.  Returns ``(A (E, 6, 6), B (E, 6, 2), c (E, 6), ok (E,))``;
    ``ok`` is False (and the row NaN) where the log is rank deficient.  C = I and D = 0 as in the reference (:106-107).
    CUDA tensors (zero copy, asynchronous) or numpy arrays (host entry point)."""
    lib = abi.load()
    if hasattr(traj, "traj") and hasattr(traj, "status"):
        res = traj
        if res.traj is None:
            raise ValueError("rollout was run without traj=True")
        traj, rows = res.traj, res.status[:, abi.S_STEPS]
    if isinstance(traj, np.ndarray):
        traj = _np(traj, np.float64)
        E, S = traj.shape[0], traj.shape[1]
        traj = _np(traj, np.float64, (E, S, abi.N_TRAJ))
        rows = np.full(E, S, np.int32) if rows is None else _np(rows, np.int32, (E,))
        AB = np.empty((E, 6, abi.SYSID_COLS))
        ok = np.empty(E, np.int32)
        abi.check(lib.bgr_identify_dynamics_host(E, S, _ptr(traj), _ptr(rows), _ptr(AB), _ptr(ok), None))
        return AB[:, :, :6], AB[:, :, 6:8], AB[:, :, 8], ok.astype(bool)
    torch = _torch()
    E, S = int(traj.shape[0]), int(traj.shape[1])
    if not traj.is_cuda or traj.dtype != torch.float64 or tuple(traj.shape) != (E, S, abi.N_TRAJ) or not traj.is_contiguous():
        raise ValueError("traj must be a contiguous float64 CUDA tensor of shape (E, S, 12)")
    rows = torch.full((E,), S, dtype=torch.int32, device=traj.device) if rows is None else rows.to(torch.int32).contiguous()
    AB = torch.empty((E, 6, abi.SYSID_COLS), dtype=torch.float64, device=traj.device)
    ok = torch.empty(E, dtype=torch.int32, device=traj.device)
    stream = torch.cuda.current_stream(traj.device).cuda_stream
    abi.check(lib.bgr_identify_dynamics(E, S, _ptr(traj), _ptr(rows), _ptr(AB), _ptr(ok), stream))
    return AB[:, :, :6], AB[:, :, 6:8], AB[:, :, 8], ok.bool()
