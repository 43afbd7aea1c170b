#!/usr/bin/env python
"""bench.py -- closed-loop sim-steps/sec of the batched rollout engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-secondary]
                    [--workload pp_sweep|stanley_lqg|stanley_lqg_strong|mpc|skidpad_mc|pp_replan]

A "step" is ONE pass of the hot path over one batch of synthetic input: one ``bgr_rollout`` of E episodes x S
closed-loop control steps (default workload = BASELINE config 2: the pure-pursuit look-ahead / PID-gain sweep,
E = 1 Mi episodes x S = 2 000 steps per GPU on the stadium-oval Trackdrive centreline, kinematic bicycle).
``value`` = sim-steps actually executed by all ranks / max-over-ranks device time, inputs resident in HBM.
``e2e`` = the same through the host-buffer C-ABI call (``bgr_rollout_host_async``: pinned numpy in, numpy out, H2D
and D2H of EVERY step inside the timed region; consecutive steps pipeline, as a streaming caller would run them).
Multi-GPU: one process per GPU (torchrun), episodes sharded, NO data-path collective; the final metric gather to
rank 0 + aggregate reduction (NCCL) is timed separately.

The ONE JSON line carries the config-2 headline and, under ``"secondary"``, the other BASELINE configurations measured
in the same process, each with its own value / e2e / roofline / clocks:
  config3_weak        Stanley + LQG + dynamic bicycle, 1 Mi episodes per GPU
  config3_strong      the same, 4 Mi episodes IN TOTAL sharded over the N ranks (BASELINE's statement of config 3)
  config4_mpc         MPC candidate evaluation, 16 384 episodes x 4 096 candidates x horizon 30 (graph replay AND one
                      API call per step)
  config5_skidpad_mc  skidpad mixed-controller Monte Carlo with the NCCL metric gather timed and checked

``--impl reference`` times the CPU arm: the reference's own Python functions as composed by oracle/loop.py
(the reference is 100 % Python and cannot be installed on the GPU box -- see DESIGN.md), on all host cores.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "closed-loop sim-steps/sec (episodes x steps)"
UNIT = "sim-steps/s"
DTYPE = "f32 laws + double-float (2 x f32) integrators for x, y, yaw, v; f64 metric sums"

WORKLOADS = {
    # name: (steer, accel, model, laps of the unrolled path, BASELINE config it is, ncu summary name)
    "pp_sweep": ("pure_pursuit", "pid", "kinematic", 8,
                 "config 2: pure-pursuit lookahead/gain sweep, 1Mi episodes x 2000 steps, stadium oval N=1024, 1 B200",
                 "pp_pid_kinematic"),
    "stanley_lqg": ("stanley", "lqg", "dynamic", 1,
                    "config 3: Stanley + LQG + dynamic bicycle (tyre slip), episodes sharded over the GPUs",
                    "stanley_lqg_dynamic"),
    "pp_replan": ("pure_pursuit", "pid", "kinematic", 1,
                  "SURVEY 8f.3: config 2's sweep in the live stack's re-planning regime (40-point local paths, every 3rd "
                  "waypoint, re-planned when the planner's index passes 3: nodes/planner.py:26), full kernel build",
                  "pp_replan"),
}
WORKLOADS["stanley_lqg_strong"] = WORKLOADS["stanley_lqg"]
# extra RolloutConfig / RolloutSpec fields per workload
WORKLOAD_EXTRA = {"pp_replan": {"replan_horizon": 40, "replan_stride": 3, "replan_after": 3}}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pp_sweep", choices=list(WORKLOADS) + ["mpc", "skidpad_mc"])
    ap.add_argument("--episodes", type=int, default=1 << 20, help="episodes per GPU per step")
    ap.add_argument("--sim-steps", type=int, default=2000, help="closed-loop control steps per episode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="only the headline workload (default: the other BASELINE configs ride along under 'secondary')")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle loop (reference functions composed; test infrastructure) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    workload, start, count, sim_steps = job
    import bgr_pathplanning_control_b200 as bgr
    from oracle import loop
    steer, accel, model, laps = WORKLOADS[workload][:4]
    ta = bgr.tracks.stadium_oval(1024)
    track = loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones)
    params = (bgr.sweeps.pp_sweep_params(count, start) if steer == "pure_pursuit"
              else bgr.sweeps.stanley_lqg_params(count, start=start))
    init = bgr.initial_states(ta, count, v0=5.0 if model == "dynamic" else 0.0)
    spec = loop.RolloutSpec(steer=steer, accel=accel, model=model, max_steps=sim_steps, laps=laps,
                            **WORKLOAD_EXTRA.get(workload, {}))
    done = 0
    t0 = time.perf_counter()
    for e in range(count):
        out = loop.rollout_episode(track, spec, params[e].astype(np.float64), init[e], record=False)
        done += out["steps"]
    return done, time.perf_counter() - t0


def cpu_sample(workload, sim_steps, seconds, cores=None):
    """Run a bounded sample of the workload's episodes through the oracle loop on ``cores`` processes.
    Returns (steps/s aggregate, cores, description)."""
    import multiprocessing as mp
    cores = cores or len(os.sched_getaffinity(0))
    # calibrate on one short episode (in this process), then size the sample for ~`seconds`
    done, dt = _cpu_worker((workload, 0, 1, min(sim_steps, 400)))
    rate1 = done / dt
    per_core = max(1, int(rate1 * seconds / sim_steps))
    per_core = min(per_core, 64)
    jobs = [(workload, 4099 * c, per_core, sim_steps) for c in range(cores)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    desc = (f"{cores} processes x {per_core} episodes x {sim_steps} steps of the same sweep rows "
            f"(oracle/loop.py: reference functions in the order of nodes/controller.py:61-63), "
            f"{total} sim-steps in {busy:.1f} s (pool wall {wall:.1f} s)")
    return total / busy, cores, desc


def c_twin_sample(workload, sim_steps, episodes_per_core=4):
    """The same sample through the plain-C oracle twin (oracle/c, OpenMP over episodes): a compiled restatement of
    the reference algorithm -- its O(N) hypot argmin per step keeps it at the Python loop's rate."""
    try:
        import bgr_pathplanning_control_b200 as bgr
        from oracle import cport, loop
        if not cport.available() or workload in WORKLOAD_EXTRA:
            return None
        steer, accel, model, laps = WORKLOADS[workload][:4]
        cores = len(os.sched_getaffinity(0))
        E = cores * episodes_per_core
        ta = bgr.tracks.stadium_oval(1024)
        params = (bgr.sweeps.pp_sweep_params(E) if steer == "pure_pursuit" else bgr.sweeps.stanley_lqg_params(E))
        init = bgr.initial_states(ta, E, v0=5.0 if model == "dynamic" else 0.0)
        spec = loop.RolloutSpec(steer=steer, accel=accel, model=model, max_steps=sim_steps, laps=laps)
        t0 = time.perf_counter()
        _, st, _ = cport.rollout(loop.Track(ta.x, ta.y, ta.kappa, True, ta.cones), spec, params.astype(np.float64), init,
                                 threads=cores)
        dt = time.perf_counter() - t0
        return {"value": float(st[:, 1].sum()) / dt, "unit": UNIT, "cores": cores,
                "sample": f"{E} episodes x {sim_steps} steps, OpenMP over episodes, {dt:.1f} s"}
    except Exception as exc:            # the C twin is optional context, never a reason to lose the bench line
        return {"unavailable": str(exc)[:200]}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    """`--impl reference`: rank 0 alone times the CPU path; the other ranks exit 0 without work."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    workload = args.workload if args.workload in WORKLOADS else "pp_sweep"
    per_step = max(2.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(min(args.warmup, 1)):
        cpu_sample(workload, min(args.sim_steps, 200), 0.5)
    vals, descs, t_total = [], [], 0.0
    cores = len(os.sched_getaffinity(0))
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, cores, desc = cpu_sample(workload, args.sim_steps, per_step)
        vals.append(v)
        descs.append(desc)
        t_total += time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[workload][4], "episodes_per_gpu": args.episodes, "sim_steps": args.sim_steps,
                   "note": "CPU arm: each step is a bounded sample of the workload's episodes, all host cores"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": descs[-1],
                         "cpu": cpu_model()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# ncu evidence: read from profiles/ at run time, matched to the loaded library by its build id
# ------------------------------------------------------------------------------------------------
def ncu_evidence(name, lib_build_id):
    """The ncu summary of kernel ``name`` captured from THIS library build, or a record saying there is none.

    scripts/make_profiles.sh stores every capture next to the build id of the library that ran
    (profiles/<dir>/build_id.txt).  Evidence is quoted only when that id equals ``bgr_build_id()`` of the library this
    process loaded; a kernel edit that has not been re-profiled shows up as ``stale`` instead of shipping old numbers."""
    found = []
    for idf in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "build_id.txt"))):
        d = os.path.dirname(idf)
        try:
            bid = open(idf).read().split()[0]
        except (OSError, IndexError):
            continue
        found.append((os.path.relpath(d, ROOT), bid))
        path = os.path.join(d, f"ncu_full_{name}.json")
        if bid != lib_build_id or not os.path.isfile(path):
            continue
        try:
            k = json.load(open(path))[0]
        except (OSError, ValueError, IndexError):
            continue

        def val(key):
            v = k.get(key)
            return v["value"] if isinstance(v, dict) and isinstance(v.get("value"), (int, float)) else None

        mb = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum")
        traffic = None
        if isinstance(rd, dict) and isinstance(wr, dict):
            traffic = rd["value"] * mb.get(rd["unit"], 1.0) + wr["value"] * mb.get(wr["unit"], 1.0)
        ev = {
            "file": os.path.relpath(path, ROOT), "file_mtime": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(os.path.getmtime(path))),
            "build_id": bid, "matches_loaded_library": True,
            "kernel_ms_under_ncu": val("gpu__time_duration.sum"),
            "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "sm_throughput_pct": val("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "pipe_fma_pct": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "pipe_xu_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "pipe_fp64_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": val("launch__registers_per_thread"),
            "smem_bank_conflicts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            "dram_bytes_per_launch": traffic,
            "warp_inst_per_warp_step": (k.get("derived") or {}).get("warp_inst_per_warp_step"),
            "hot_lines": os.path.relpath(os.path.join(d, f"hot_lines_{name}.txt"), ROOT),
        }
        ops = os.path.join(d, f"opcodes_{name}.json")
        if os.path.isfile(ops):
            try:
                o = json.load(open(ops))
                ev["executed_fp32_flops_per_step"] = o.get("fp32_flops_per_lane_step")
                ev["executed_fp32_issue_slots_per_step"] = o.get("fp32_warp_inst_per_warp_step")
            except (OSError, ValueError):
                pass
        return ev
    return {"stale": True, "matches_loaded_library": False, "library_build_id": lib_build_id,
            "profiles_found": [f"{d} ({b})" for d, b in found],
            "note": "no ncu capture of this library build under profiles/: run scripts/gpu_round.sh + "
                    "scripts/make_profiles.sh; evidence of an older build is deliberately not quoted"}


def issue_roofline(ev, sim_steps_per_s, clocks):
    if not ev or ev.get("stale") or not ev.get("warp_inst_per_warp_step"):
        return None
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak = 148 * 4 * mhz * 1e6                                     # warp-instructions per second
    achieved = sim_steps_per_s / 32.0 * ev["warp_inst_per_warp_step"]
    return {"bound": "issue", "achieved": achieved, "peak": peak, "unit": "warp-inst/s", "frac": achieved / peak,
            "inst_per_step_source": ev["file"]}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide plumbing: torch, the rank layout, barriers and max / sum over ranks."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        import bgr_pathplanning_control_b200 as bgr
        from bgr_pathplanning_control_b200 import _abi as abi
        from bgr_pathplanning_control_b200 import roofline as rf
        self.torch, self.dist, self.bgr, self.abi, self.rf = torch, dist, bgr, abi, rf
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.build_id = bgr.build_id()
        self._fp32_peak = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def fp32_peak(self):
        if self._fp32_peak is None:
            self._fp32_peak = self.bgr.fp32_peak_probe(self.local)
        return self._fp32_peak

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def run_rollout_workload(ctx, workload, E, S, steps, warmup, *, strong_total=None, cpu_seconds=0.0, time_gather=True):
    """One rollout workload on this rank's E episodes (``strong_total``: E = this rank's shard of that many).  Returns
    the JSON record on rank 0, None elsewhere."""
    torch, bgr, abi, rf = ctx.torch, ctx.bgr, ctx.abi, ctx.rf
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    steer, accel, model, laps, desc, ncu_name = WORKLOADS[workload]
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta, device=ctx.local)
    if strong_total is not None:
        lo, hi = bgr.sharding.shard_range(strong_total, rank, world)
        E, row0 = hi - lo, lo
    else:
        row0 = rank * E
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=laps, episode_id_base=row0,
                            **WORKLOAD_EXTRA.get(workload, {}))
    # Input sets: each timed step consumes a DIFFERENT batch (rows of the sweep shifted per set), so the inputs
    # touched over the timed region (n_sets x 112 B x E) exceed the 126 MB L2.
    h2d = E * (abi.N_PARAMS * 4 + abi.N_STATE * 8)
    n_sets = 4 if 4 * h2d <= (1 << 30) else 2
    make = bgr.sweeps.pp_sweep_params if steer == "pure_pursuit" else \
        (lambda n, start: bgr.sweeps.stanley_lqg_params(n, start=start))
    if strong_total is not None:     # the 4 Mi-episode sweep itself, this rank's rows (sets: the sweep shifted by s)
        h_params = [torch.from_numpy(make(E, row0 + s * 65537)).pin_memory() for s in range(n_sets)]
    else:
        h_params = [torch.from_numpy(make(E, (rank * n_sets + s) * 65537)).pin_memory() for s in range(n_sets)]
    v0 = 5.0 if model == "dynamic" else 0.0     # the Euler-integrated tyre model is unstable below ~3.2 m/s
    h_init = [torch.from_numpy(bgr.initial_states(ta, E, lateral_offset=np.linspace(-0.2, 0.2, E) * (s % 2), v0=v0))
              .pin_memory() for s in range(n_sets)]
    d_params = [t.to(dev) for t in h_params]
    d_init = [t.to(dev) for t in h_init]
    out = bgr.RolloutResult(torch.empty((E, abi.N_METRICS), dtype=torch.float32, device=dev),
                            torch.empty((E, abi.N_STATUS), dtype=torch.int32, device=dev), None,
                            torch.empty(abi.N_AGG, dtype=torch.float64, device=dev))

    def step(i):
        return bgr.rollout(track, cfg, d_params[i % n_sets], d_init[i % n_sets], aggregate=True, out=out)

    for i in range(warmup):
        step(i)
    ctx.barrier()
    launches0 = bgr.launch_count()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    steps_done = torch.zeros((), dtype=torch.float64, device=dev)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    t_begin.record()
    for i in range(steps):
        ev[i][0].record()
        r = step(i)
        ev[i][1].record()
        steps_done += r.aggregate[abi.A_STEPS]
    t_end.record()
    ctx.barrier()
    launches = bgr.launch_count() - launches0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    kernel_ms_last = bgr.last_kernel_ms()
    ms_total = ctx.reduce(t_begin.elapsed_time(t_end), "MAX")
    my_steps = float(steps_done.item())
    total_sim_steps = ctx.reduce(my_steps, "SUM")
    value = total_sim_steps / (ms_total * 1e-3)

    # ---- final metric gather to rank 0 + aggregate reduction: the ONLY cross-rank traffic (timed separately) ----
    gather_ms = None
    if world > 1 and time_gather:
        n_all = strong_total if strong_total is not None else E * world
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bgr.sharding.gather_results(out.metrics, out.status, out.aggregate, n_all)   # warm-up: NCCL channels
        ctx.barrier()
        g0.record()
        m_all, s_all, agg_all = bgr.sharding.gather_results(out.metrics, out.status, out.aggregate, n_all)
        g1.record()
        ctx.barrier()
        gather_ms = ctx.reduce(g0.elapsed_time(g1), "MAX")
        assert float(agg_all[abi.A_EPISODES]) == n_all
        if rank == 0:
            assert m_all.shape[0] == n_all and s_all.shape[0] == n_all
            assert torch.equal(m_all[:E], out.metrics) and torch.equal(s_all[:E], out.status)
        else:
            assert m_all is None and s_all is None
        del m_all, s_all

    # ---- e2e: host buffers through the C ABI; every step uploads its inputs and downloads its results inside the
    #      timed region; consecutive steps pipeline (bgr_rollout_host_async, two result buffers in flight) ----
    np_params = [t.numpy() for t in h_params]
    np_init = [t.numpy() for t in h_init]
    h_out = [bgr.RolloutResult(torch.empty((E, abi.N_METRICS), dtype=torch.float32).pin_memory().numpy(),
                               torch.empty((E, abi.N_STATUS), dtype=torch.int32).pin_memory().numpy(), None,
                               torch.empty(abi.N_AGG, dtype=torch.float64).pin_memory().numpy()) for _ in range(2)]
    # warm-up through the SAME asynchronous path: four calls walk the library's ring of three device buffer slots (each
    # slot is ~200 MB of cudaMalloc the first time it is used) and create the four completion streams of this thread, so
    # that no allocation or stream creation falls into the timed region (seen once on a 4-GPU box: a timed loop of 10
    # steps took twice its time on one rank, the synchronous loop after it did not)
    for i in range(4):
        bgr.rollout_async(track, cfg, np_params[i % n_sets], np_init[i % n_sets], aggregate=True, out=h_out[i % 2]).wait()
    ctx.barrier()
    e2e_steps = 0.0
    w0 = time.perf_counter()
    pending = None
    marks = [w0]                                                          # host clock when each step's result was read
    for i in range(steps):
        nxt = bgr.rollout_async(track, cfg, np_params[i % n_sets], np_init[i % n_sets], aggregate=True, out=h_out[i % 2])
        if pending is not None:
            e2e_steps += float(pending.wait().aggregate[abi.A_STEPS])     # the result of step i - 1 is read on the host
            marks.append(time.perf_counter())
        pending = nxt
    e2e_steps += float(pending.wait().aggregate[abi.A_STEPS])
    marks.append(time.perf_counter())
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - w0) * 1e3
    e2e_step_ms = [round((b - a) * 1e3, 3) for a, b in zip(marks[:-1], marks[1:])]
    # the same call, one step at a time (upload, run, download, synchronise; nothing in flight across steps)
    w1 = time.perf_counter()
    sync_steps = 0.0
    for i in range(min(steps, 5)):
        r = bgr.rollout(track, cfg, np_params[i % n_sets], np_init[i % n_sets], aggregate=True, out=h_out[0])
        sync_steps += float(r.aggregate[abi.A_STEPS])
    sync_ms = (time.perf_counter() - w1) * 1e3
    if world > 1:
        ctx.dist.barrier()
    e2e_value = ctx.reduce(e2e_steps, "SUM") / (ctx.reduce(e2e_ms, "MAX") * 1e-3)
    sync_value = ctx.reduce(sync_steps, "SUM") / (ctx.reduce(sync_ms, "MAX") * 1e-3)
    e2e_slowest_ms = ctx.reduce(max(e2e_step_ms), "MAX")         # a one-off stall on ANY rank shows here
    d2h = E * (abi.N_METRICS * 4 + abi.N_STATUS * 4) + abi.N_AGG * 8
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None

    peaks, peaks_src = measured_peaks()
    fp32_tflops, _ = ctx.fp32_peak()
    flops_step = rf.flops_per_step(steer, accel, model)
    k_ms = float(np.mean(kernel_ms))
    steps_per_launch = my_steps / steps
    rate = steps_per_launch / (k_ms * 1e-3)                      # sim-steps/s of this rank's kernel
    achieved_tflops = rate * flops_step / 1e12
    bytes_launch = E * rf.bytes_per_episode()
    hbm_gbs = bytes_launch / (k_ms * 1e-3) / 1e9
    ncu = ncu_evidence(ncu_name, ctx.build_id)
    roof = {
        "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_tflops, "unit": "TFLOP/s",
        "frac": achieved_tflops / fp32_tflops, "traffic": ncu.get("dram_bytes_per_launch"),
        "peak_source": "bgr_fp32_peak_probe (FFMA chains, measured in this run); MEASURED_PEAKS.json has no fp32 row",
        "flops_per_sim_step": flops_step, "kernel_ms_avg": k_ms, "kernel_ms_last": kernel_ms_last,
        "step_ms": [round(x, 3) for x in kernel_ms],
        "sim_steps_per_launch": steps_per_launch,
        # what ncu measured for THIS library build (profiles/<dir>/, matched by build id): the kernel is bound by
        # INSTRUCTION ISSUE, not by DRAM or tensor throughput
        "ncu": ncu,
        # the roofline that actually binds: warp-instructions issued per second against the issue peak of the chip
        # (148 SMs x 4 schedulers x 1 instruction per cycle at the SM clock seen under load)
        "issue": issue_roofline(ncu, rate, clocks),
        "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                "bytes_per_launch": bytes_launch, "peak_source": peaks_src},
    }
    if ncu.get("executed_fp32_flops_per_step"):
        ex = rate * ncu["executed_fp32_flops_per_step"] / 1e12
        roof["executed_fp32_tflops"] = ex
        roof["executed_fp32_frac"] = ex / fp32_tflops
    rec = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_total / steps, "higher_is_better": True,
        "scaling": "strong" if strong_total is not None else "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": desc, "episodes_per_gpu": E, "episodes_total": strong_total if strong_total is not None else E * world,
                   "sim_steps": S, "steer": steer, "accel": accel, "model": model, "track": "stadium_oval n=1024",
                   "parallelism": f"episode sharding x{world}", "library_build_id": ctx.build_id,
                   "l2": f"{n_sets} rotating input sets ({n_sets * h2d / 1e6:.0f} MB) > 126 MB L2, no flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "bgr_rollout_host_async (pinned numpy in, numpy out; step i + 1 queued before step i is read)",
                "frac_of_value": e2e_value / value,
                # result-to-result intervals on rank 0's host clock (the first one includes filling the pipeline), and the
                # longest interval any rank saw: a one-off host or driver stall is visible instead of silently averaged in
                "step_ms": e2e_step_ms, "slowest_step_ms_any_rank": e2e_slowest_ms,
                "one_call_at_a_time": {"value": sync_value, "api": "bgr_rollout_host", "frac_of_value": sync_value / value}},
        "gpu_launches": int(launches),
        "roofline": roof,
    }
    if gather_ms is not None:
        rec["gather_ms"] = gather_ms
        rec["gather"] = "per-episode rows to rank 0 (dist.gather) + 64 B aggregate rows all-gathered, summed in rank order"
    if world == 1 and cpu_seconds > 0:
        v, cores, sample = cpu_sample(workload, S, cpu_seconds)
        rec["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                               "cpu": cpu_model(), "c_twin": c_twin_sample(workload, S)}
    return rec


def run_skidpad_mc(ctx, episodes, sim_steps, steps, warmup):
    """BASELINE config 5: skidpad figure-eight, mixed-controller Monte Carlo with Philox pose noise and perturbed cones;
    episodes are grouped by controller kind on the host (three homogeneous launches per step on concurrent streams);
    the final per-episode metric gather to rank 0 + aggregate reduction goes through NCCL."""
    torch, bgr, abi, rf = ctx.torch, ctx.bgr, ctx.abi, ctx.rf
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    kinds = [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]
    per = max(256, (episodes // 2) // len(kinds) // 256 * 256)       # 512 Ki episodes per GPU in total by default
    S = min(sim_steps, 1500)
    ta = bgr.tracks.skidpad(2048)
    track = bgr.Track.from_arrays(ta, device=ctx.local, cone_reach=4.0)
    noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
    for k in os.environ.get("BGR_MC_OFF", "").split(","):      # diagnostic: switch parts of the Monte Carlo off
        if k == "cone_noise":
            noise["sigma_cone"] = 0.0
        elif k == "cones":
            noise["hit_radius"] = 0.0
        elif k == "pose_noise":
            noise.update(sigma_xy=0.0, sigma_yaw=0.0, sigma_v=0.0)
    jobs, dynamic = [], []
    for g, (steer, accel, model) in enumerate(kinds):
        lo = ((rank + int(os.environ.get("BGR_MC_ROW_RANK", "0"))) * len(kinds) + g) * per   # (diagnostic: another rank's rows)
        # diagnostics: BGR_MC_SPLIT=<steps> (two-phase run), BGR_MC_DYNAMIC=1 (dynamic schedule), both for the kinds
        # listed in BGR_MC_KINDS (default: all), e.g. BGR_MC_KINDS=1,2 = the two Stanley kinds
        on = str(g) in os.environ.get("BGR_MC_KINDS", "0,1,2").split(",")
        split = max(0, int(os.environ.get("BGR_MC_SPLIT", "0"))) if on else 0
        # Schedules (rows are bit-identical under all of them): the pure-pursuit episodes end where the path ends, after
        # 746 .. 1 241 steps depending on their target speed -- the library launches them slowest-first on its own.  The
        # noisy Stanley episodes end at the first crossing (99.7 % within 160 steps) or, for one in 400 .. 2 500, run on
        # for up to 1 500 steps: heavy-tailed and unpredictable, the case the dynamic schedule exists for.
        dyn = steer == "stanley"
        if "BGR_MC_DYNAMIC" in os.environ:
            dyn = on and os.environ["BGR_MC_DYNAMIC"] == "1"
        dynamic.append(dyn)
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=1, episode_id_base=lo,
                                split_steps=split, dynamic_schedule=dyn, **noise)
        init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0)
        init[:, abi.YAW] = 0.0
        params = bgr.sweeps.monte_carlo_params(per, start=lo)
        out = bgr.RolloutResult(torch.empty((per, abi.N_METRICS), dtype=torch.float32, device=dev),
                                torch.empty((per, abi.N_STATUS), dtype=torch.int32, device=dev), None,
                                torch.empty(abi.N_AGG, dtype=torch.float64, device=dev))
        jobs.append((cfg, torch.from_numpy(params).to(dev), torch.from_numpy(init).to(dev), out,
                     torch.from_numpy(params).pin_memory().numpy(), torch.from_numpy(init).pin_memory().numpy()))

    # the three kinds run on concurrent streams, the latency-bound Stanley kinds first (engine.rollout_groups);
    # BGR_MC_SERIAL=1 launches them back to back on one stream instead
    serial = os.environ.get("BGR_MC_SERIAL", "0") == "1"

    def step():
        if serial:
            return [bgr.rollout(track, j[0], j[1], j[2], aggregate=True, out=j[3]) for j in jobs]
        return bgr.rollout_groups(track, [(j[0], j[1], j[2]) for j in jobs], aggregate=True,
                                  outs=[j[3] for j in jobs], order=[2, 1, 0])

    for _ in range(warmup):
        step()
    ctx.barrier()
    launches0 = bgr.launch_count()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    done = torch.zeros((), dtype=torch.float64, device=dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    t0.record()
    for _ in range(steps):
        for r in step():
            done += r.aggregate[abi.A_STEPS]
    t1.record()
    ctx.barrier()
    launches = bgr.launch_count() - launches0
    ms_total = ctx.reduce(t0.elapsed_time(t1), "MAX")
    my_steps = float(done.item())
    total_steps = ctx.reduce(my_steps, "SUM")
    gather_ms = None
    if world > 1:
        m = torch.cat([j[3].metrics for j in jobs])
        s = torch.cat([j[3].status for j in jobs])
        agg = torch.from_numpy(bgr.sharding.combine_aggregates([j[3].aggregate.cpu().numpy() for j in jobs])).to(dev)
        n_all = m.shape[0] * world
        bgr.sharding.gather_results(m, s, agg, n_all)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        g0.record()
        m_all, s_all, agg_all = bgr.sharding.gather_results(m, s, agg, n_all)
        g1.record()
        ctx.barrier()
        gather_ms = ctx.reduce(g0.elapsed_time(g1), "MAX")
        # the gathered result is checked, not just timed: episode count, step total, and rank 0's own rows in place
        assert float(agg_all[abi.A_EPISODES]) == n_all
        assert abs(float(agg_all[abi.A_STEPS]) - total_steps / steps) < 0.5
        if rank == 0:
            assert m_all.shape[0] == n_all and torch.equal(s_all[: s.shape[0]], s)
            assert int(s_all[:, abi.S_STEPS].to(torch.int64).sum().item()) == int(round(total_steps / steps))
        del m_all, s_all
    # e2e: host buffers.  (a) STREAMING (the headline e2e): every kind is queued from a host thread of its own through
    # bgr_rollout_host_async, batch i + 1 before batch i is read (engine.rollout_groups_async) -- every batch's inputs
    # cross PCIe from pinned memory and its per-episode rows come back inside the timed region; (b) one synchronous
    # rollout_groups call per batch (one persistent host thread per kind through bgr_rollout_host).
    # (result buffers pinned like the inputs: a download into pageable memory goes through the driver's staging buffer
    #  and blocks the host thread that issued it)
    def host_outs():
        return [bgr.RolloutResult(torch.empty((per, abi.N_METRICS), dtype=torch.float32).pin_memory().numpy(),
                                  torch.empty((per, abi.N_STATUS), dtype=torch.int32).pin_memory().numpy(), None,
                                  torch.empty(abi.N_AGG, dtype=torch.float64).pin_memory().numpy()) for _ in jobs]
    h_sets = [host_outs(), host_outs()]
    h_outs = h_sets[0]
    h_groups = [(j[0], j[4], j[5]) for j in jobs]
    e2e_value = sync_value = None
    if not serial:
        for i in range(4):      # (the lane threads build their copy / compute pipelines, buffer rings and streams on first use)
            bgr.rollout_groups_async(track, h_groups, aggregate=True, outs=h_sets[i % 2], order=[2, 1, 0]).wait()
        ctx.barrier()
        w0 = time.perf_counter()
        e2e_steps, pending = 0.0, None
        for i in range(steps):
            nxt = bgr.rollout_groups_async(track, h_groups, aggregate=True, outs=h_sets[i % 2], order=[2, 1, 0])
            if pending is not None:
                e2e_steps += sum(float(r.aggregate[abi.A_STEPS]) for r in pending.wait())
            pending = nxt
        e2e_steps += sum(float(r.aggregate[abi.A_STEPS]) for r in pending.wait())
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - w0
        e2e_value = ctx.reduce(e2e_steps, "SUM") / ctx.reduce(e2e_s, "MAX")
    for _ in range(2):      # (the worker threads build their copy / compute pipelines on first use)
        if not serial:
            bgr.rollout_groups(track, h_groups, aggregate=True, outs=h_outs, order=[2, 1, 0])
    ctx.barrier()
    w0 = time.perf_counter()
    sync_steps = 0.0
    for _ in range(steps):
        if serial:
            rs = [bgr.rollout(track, j[0], j[4], j[5], aggregate=True, out=o) for j, o in zip(jobs, h_outs)]
        else:
            rs = bgr.rollout_groups(track, h_groups, aggregate=True, outs=h_outs, order=[2, 1, 0])
        sync_steps += sum(float(r.aggregate[abi.A_STEPS]) for r in rs)
    sync_s = time.perf_counter() - w0
    sync_value = ctx.reduce(sync_steps, "SUM") / ctx.reduce(sync_s, "MAX")
    if e2e_value is None:
        e2e_value = sync_value
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    fp32_tflops, _ = ctx.fp32_peak()
    flops = np.mean([rf.flops_per_step(*k) for k in kinds])
    ach = my_steps * flops / (ms_total * 1e-3) / 1e12
    hits = sum(float(j[3].aggregate[abi.A_CONE_HITS]) for j in jobs)
    ncu = ncu_evidence("config5_pp", ctx.build_id)
    rec = {
        "metric": METRIC, "value": total_steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": "config 5: skidpad figure-eight, mixed-controller Monte Carlo with noisy pose / cone "
                               "perturbations (Philox4x32-10), NCCL metric gather", "episodes_per_gpu": per * len(kinds),
                   "kinds": ["+".join(k) for k in kinds], "sim_steps_cap": S, "track": "skidpad n=2048, 45 cones",
                   "cone_hits_last_step_rank0": hits, "parallelism": f"episode sharding x{world}",
                   "launch": "kinds back to back on one stream" if serial else
                             "kinds on three concurrent streams, latency-bound Stanley kinds first",
                   "schedule": ["dynamic (one-wave grid; lanes pull their next episode from a queue)" if d else
                                "static, launched in order of target speed (slowest first)" for d in dynamic],
                   "l2": "state lives in registers; 112 B in + 88 B out per episode"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": per * len(kinds) * 112,
                "d2h_bytes_per_step": per * len(kinds) * 88 + 3 * 64,
                "api": "rollout_groups (one bgr_rollout_host per kind, back to back)" if serial else
                       "rollout_groups_async (one host thread per kind through bgr_rollout_host_async; pinned numpy in, "
                       "numpy out; batch i + 1 queued before batch i is read)",
                "frac_of_value": e2e_value / (total_steps / (ms_total * 1e-3)),
                "one_call_at_a_time": {"value": sync_value, "api": "rollout_groups (one persistent host thread per kind "
                                       "through bgr_rollout_host, synchronous)",
                                       "frac_of_value": sync_value / (total_steps / (ms_total * 1e-3))}},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_tflops, "unit": "TFLOP/s", "frac": ach / fp32_tflops,
                     "traffic": ncu.get("dram_bytes_per_launch"), "ncu": ncu,
                     "note": "full kernel build (cones + noise + audit channel)"},
    }
    if gather_ms is not None:
        rec["gather_ms"] = gather_ms
    return rec


def run_mpc(ctx, steps, warmup, E=16384, K=4096, H=30):
    """BASELINE config 4: 16k episodes x 4096 candidates x horizon 30, as a CUDA-graph replay of the API call AND with
    one API call per step; e2e through the device-resident candidate bank handle."""
    torch, bgr, abi, rf = ctx.torch, ctx.bgr, ctx.abi, ctx.rf
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta, device=ctx.local)
    cfg = bgr.MpcConfig(horizon=H, dt=0.05, explicit_rollout=bool(int(os.environ.get("BGR_MPC_EXPLICIT", "0"))))
    states, ref_start = bgr.sweeps.mpc_states(ta, E, start=rank * E)
    bank_kind = os.environ.get("BGR_MPC_BANK", "qp_family")      # the drop-in MPCController's bank (engine.candidate_bank)
    bank = bgr.candidate_bank(K, H, kind=bank_kind, dt=0.05)
    d_state, d_ref, d_bank = (torch.from_numpy(a).to(dev) for a in (states, ref_start, bank))
    for _ in range(warmup):
        bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
    ctx.barrier()
    kernel_ms = bgr.last_kernel_ms()
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    units_step = float(E) * K * H
    # ---- (a) one API call per step ----
    ctx.barrier()
    launches0 = bgr.launch_count()
    n_call = max(steps, 50)                                   # (a call is ~0.1 ms: enough of them for the clock sampler)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(n_call):
        res = bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
    t1.record()
    ctx.barrier()
    call_launches = bgr.launch_count() - launches0
    ms_call = ctx.reduce(t0.elapsed_time(t1), "MAX") / n_call
    # ---- (b) the same call captured ONCE into a CUDA graph (stream-ordered scratch allocation, the kernels, the free)
    #      and replayed per step: one evaluation is ~0.06 ms on the device behind ~0.09 ms of Python + launch work ----
    graph, launch_mode, ms_graph = None, "one API call per step", None
    if os.environ.get("BGR_MPC_GRAPH", "1") != "0":
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gres = bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
            g.replay()
            torch.cuda.synchronize()
            if not (torch.equal(gres.best_idx, res.best_idx) and torch.equal(gres.best_cost, res.best_cost)):
                raise RuntimeError("graph replay disagrees with the eager call")
            graph, launch_mode = g, "CUDA graph replay of one bgr_mpc_eval call"
        except Exception as exc:             # noqa: BLE001 -- the eager path is always available
            launch_mode = f"one API call per step (graph capture unavailable: {str(exc)[:80]})"
            torch.cuda.synchronize()
    per_call = 1 if cfg.explicit_rollout else 2
    n_rep = max(steps, 200)
    if graph is not None:
        ctx.barrier()
        t0.record()
        for _ in range(n_rep):
            graph.replay()
        t1.record()
        ctx.barrier()
        ms_graph = ctx.reduce(t0.elapsed_time(t1), "MAX") / n_rep
    ms = ms_graph if ms_graph is not None else ms_call
    timed_launches = (per_call + 1) * n_rep if graph is not None else call_launches
    # ---- e2e: host states in, host commands out, per step; the candidate bank stays on the device behind a handle ----
    e2e, e2e_api = None, None
    try:
        hbank = bgr.MpcBank(bank, device=ctx.local)
        # a streaming caller keeps its state / command buffers pinned and reuses them every control period
        p_states = torch.from_numpy(states).pin_memory().numpy()
        p_ref = torch.from_numpy(ref_start).pin_memory().numpy()
        p_out = bgr.MpcResult(torch.empty((E, 2), dtype=torch.float32).pin_memory().numpy(),
                              torch.empty(E, dtype=torch.float32).pin_memory().numpy(),
                              torch.empty(E, dtype=torch.int32).pin_memory().numpy())
        for _ in range(2):
            bgr.mpc_evaluate(track, cfg, p_states, p_ref, hbank, out=p_out)
        assert np.array_equal(p_out.best_idx, res.best_idx.cpu().numpy())
        n_e2e = max(steps, 50)
        w0 = time.perf_counter()
        for _ in range(n_e2e):
            bgr.mpc_evaluate(track, cfg, p_states, p_ref, hbank, out=p_out)
        e2e_sync_s = (time.perf_counter() - w0) / n_e2e
        # ... and the streaming form: two sets of pinned result buffers, batch i + 1 queued before batch i is read
        # (bgr_mpc_eval_bank_host_async): the upload of one batch runs under the kernels of the other
        p_out2 = [p_out, bgr.MpcResult(torch.empty((E, 2), dtype=torch.float32).pin_memory().numpy(),
                                       torch.empty(E, dtype=torch.float32).pin_memory().numpy(),
                                       torch.empty(E, dtype=torch.int32).pin_memory().numpy())]
        p_in2 = [(p_states, p_ref), (torch.from_numpy(states).pin_memory().numpy(), torch.from_numpy(ref_start).pin_memory().numpy())]
        for i in range(4):
            bgr.mpc_evaluate_async(track, cfg, *p_in2[i & 1], hbank, out=p_out2[i & 1]).wait()
        assert np.array_equal(p_out2[1].best_idx, res.best_idx.cpu().numpy())
        w0 = time.perf_counter()
        pend = bgr.mpc_evaluate_async(track, cfg, *p_in2[0], hbank, out=p_out2[0])
        for i in range(1, n_e2e):
            nxt = bgr.mpc_evaluate_async(track, cfg, *p_in2[i & 1], hbank, out=p_out2[i & 1])
            got = pend.wait()
            _ = got.best_u[0, 0]                      # (the step's result is read on the host)
            pend = nxt
        pend.wait()
        e2e_s = (time.perf_counter() - w0) / n_e2e
        e2e_api = ("bgr_mpc_eval_bank_host_async (pinned numpy states in, commands out, resident candidate bank; batch "
                   "i + 1 queued before batch i is read)")
        h2d_step = states.nbytes + ref_start.nbytes
    except AttributeError:
        e2e_sync_s = None
        n_e2e = max(steps, 20)
        w0 = time.perf_counter()
        for _ in range(n_e2e):
            bgr.mpc_evaluate(track, cfg, states, ref_start, bank)
        e2e_s = (time.perf_counter() - w0) / n_e2e
        e2e_api = "bgr_mpc_eval_host (bank uploaded per call)"
        h2d_step = states.nbytes + ref_start.nbytes + bank.nbytes
    e2e = units_step * world / ctx.reduce(e2e_s, "MAX")
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return None
    ncu = ncu_evidence("mpc_factored", ctx.build_id)
    if cfg.explicit_rollout:
        fp32_tflops, _ = ctx.fp32_peak()
        ach = units_step * rf.FLOPS["mpc_candidate_step"] / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "fp32", "achieved": ach, "peak": fp32_tflops, "unit": "TFLOP/s", "frac": ach / fp32_tflops,
                "traffic": None, "kernel_ms_last": kernel_ms}
        dtype = "f32"
    else:
        # factored evaluator: 3 fp64 FMAs + 1 add per (episode, candidate) pair = 7 flops on the FP64 CUDA-core pipe;
        # the peak is MEASURED in this run (DFMA chains, bgr_fp64_peak_probe)
        fp64_tflops = bgr.fp64_peak_probe(ctx.local)
        ach = float(E) * K * 7.0 / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "fp64", "achieved": ach, "peak": fp64_tflops, "unit": "TFLOP/s", "frac": ach / fp64_tflops,
                "traffic": ncu.get("dram_bytes_per_launch"), "kernel_ms_last": kernel_ms, "ncu": ncu,
                "peak_source": "bgr_fp64_peak_probe (DFMA chains, measured in this run)",
                "note": "the explicit 30-stage rollout needs 29 flops per candidate-step; the factored form needs 7 per "
                        "(episode, candidate) PAIR -- candidate-steps/s is the BASELINE metric, this roofline is of the "
                        "kernel that runs"}
        dtype = "f64 (factored cost in fp64; the reported cost is rounded to f32)"
    value = units_step * world / (ms * 1e-3)
    return {
        "metric": "MPC candidate-steps/sec (episodes x candidates x horizon)", "value": value,
        "unit": "candidate-steps/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "timed_evaluations": n_rep if graph is not None else n_call,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": "config 4: MPC candidate-sequence rollout/cost evaluation, 16384 episodes x 4096 "
                               "candidates x horizon 30", "solves_per_s": E * world / (ms * 1e-3),
                   "evaluator": "explicit fp32 stage rollout" if cfg.explicit_rollout else
                   "factored (prefix-sum statistics; see DESIGN.md 3.2)",
                   "launch": launch_mode, "bank": bank_kind, "library_build_id": ctx.build_id},
        "one_call_per_step": {"value": units_step * world / (ms_call * 1e-3), "ms_per_step": ms_call,
                              "launch": "bgr.mpc_evaluate (Python -> ctypes -> three kernels) per step"},
        "clocks": clocks,
        "e2e": {"value": e2e, "unit": "candidate-steps/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": E * 16,
                "api": e2e_api, "frac_of_value": e2e / value,
                "frac_of_one_call_per_step": e2e / (units_step * world / (ms_call * 1e-3)),
                # the same batches one synchronous call at a time (bgr_mpc_eval_bank_host): upload, three kernels and
                # download in sequence -- the latency of ONE control period, which a closed loop cannot overlap
                "synchronous": None if e2e_sync_s is None else
                {"value": units_step * world / e2e_sync_s, "ms_per_call": e2e_sync_s * 1e3,
                 "api": "bgr_mpc_eval_bank_host (rank 0's own clock)"}},
        "gpu_launches": int(timed_launches),
        "roofline": roof,
    }


def run_b200(args):
    ctx = Ctx()
    sec_steps = max(3, min(args.steps, 10))
    sec_warm = max(3, min(args.warmup, 3))
    cpu_s = 0.0 if args.no_cpu_baseline else args.cpu_seconds
    if args.workload == "mpc":
        line = run_mpc(ctx, args.steps, args.warmup)
    elif args.workload == "skidpad_mc":
        line = run_skidpad_mc(ctx, args.episodes, args.sim_steps, args.steps, args.warmup)
    elif args.workload == "stanley_lqg_strong":
        line = run_rollout_workload(ctx, "stanley_lqg_strong", 0, args.sim_steps, args.steps, args.warmup,
                                    strong_total=4 * args.episodes, cpu_seconds=0.0)
    else:
        line = run_rollout_workload(ctx, args.workload, args.episodes, args.sim_steps, args.steps, args.warmup,
                                    cpu_seconds=cpu_s)
        if args.workload == "pp_sweep" and not args.no_secondary:
            # the other BASELINE configurations, measured in the same process on the same box
            sec = {}

            def attempt(name, fn):
                try:
                    r = fn()
                except Exception as exc:      # noqa: BLE001 -- a failing secondary must not cost the headline line
                    ctx.torch.cuda.synchronize()
                    r = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
                if ctx.rank == 0:
                    sec[name] = r

            attempt("config3_weak", lambda: run_rollout_workload(
                ctx, "stanley_lqg", args.episodes, args.sim_steps, sec_steps, sec_warm))
            attempt("config3_strong", lambda: run_rollout_workload(
                ctx, "stanley_lqg_strong", 0, args.sim_steps, max(3, sec_steps // 2), sec_warm,
                strong_total=4 * args.episodes))
            attempt("config4_mpc", lambda: run_mpc(ctx, sec_steps, sec_warm))
            attempt("config5_skidpad_mc", lambda: run_skidpad_mc(ctx, args.episodes, args.sim_steps, sec_steps, sec_warm))
            if line is not None:
                line["secondary"] = sec
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
