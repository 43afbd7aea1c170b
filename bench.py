#!/usr/bin/env python
"""bench.py -- closed-loop sim-steps/sec of the batched rollout engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload pp_sweep|stanley_lqg|mpc|skidpad_mc|pp_replan]

A "step" is ONE pass of the hot path over one batch of synthetic input: one ``bgr_rollout`` of E episodes x S
closed-loop control steps (default workload = BASELINE config 2: the pure-pursuit look-ahead / PID-gain sweep,
E = 1 Mi episodes x S = 2 000 steps per GPU on the stadium-oval Trackdrive centreline, kinematic bicycle).
``value`` = sim-steps actually executed by all ranks / max-over-ranks device time, inputs resident in HBM.
``e2e`` = the same through the host-buffer C-ABI call (``bgr_rollout_host``: pinned numpy in, numpy out, H2D and
D2H inside the timed region).  Multi-GPU: one process per GPU (torchrun), episodes sharded, NO data-path
collective (scaling: weak); the final metric gather + aggregate reduction (NCCL) is timed separately.

``--impl reference`` times the CPU arm: the reference's own Python functions as composed by oracle/loop.py
(the reference is 100 % Python and cannot be installed on the GPU box -- see DESIGN.md), on all host cores.
"""
from __future__ import annotations

import argparse
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

# ncu --set full --clock-control none, one capture of the step kernel per change (profiles/r1_final/*.json)
NCU_EVIDENCE = {
    "pp_sweep": {"file": "profiles/r1_final/ncu_full_pp_pid_kinematic.json", "issue_slots_busy_pct": 86.6,
                 "sm_throughput_pct": 85.2, "pipe_fma_pct": 28.5, "pipe_fmaheavy_cycles_pct": 35.8, "pipe_alu_pct": 51.5,
                 "pipe_xu_pct": 46.7, "pipe_fp64_pct": 6.5, "dram_throughput_pct": 0.12,
                 "warp_inst_per_warp_step": 269.1, "registers": 64, "dram_bytes_per_launch": 169.1e6,
                 "hot_lines": "profiles/r1_final/hot_lines_pp_pid_kinematic.txt"},
    "stanley_lqg": {"file": "profiles/r1_final/ncu_full_stanley_lqg_dynamic.json", "issue_slots_busy_pct": 86.1,
                    "sm_throughput_pct": 84.9, "pipe_fma_pct": 34.2, "pipe_fmaheavy_cycles_pct": 33.5,
                    "pipe_alu_pct": 50.6, "pipe_xu_pct": 47.0, "pipe_fp64_pct": 6.6, "dram_throughput_pct": 0.10,
                    "warp_inst_per_warp_step": 303.1, "registers": 64, "dram_bytes_per_launch": 169.2e6,
                    "hot_lines": "profiles/r1_final/hot_lines_stanley_lqg_dynamic.txt"},
}

WORKLOADS = {
    # name: (steer, accel, model, laps of the unrolled path, BASELINE config it is)
    "pp_sweep": ("pure_pursuit", "pid", "kinematic", 8,
                 "config 2: pure-pursuit lookahead/gain sweep, 1Mi episodes x 2000 steps, stadium oval N=1024, 1 B200"),
    "stanley_lqg": ("stanley", "lqg", "dynamic", 1,
                    "config 3: Stanley + LQG + dynamic bicycle (tyre slip), episodes sharded over the GPUs"),
    "pp_replan": ("pure_pursuit", "pid", "kinematic", 1,
                  "SURVEY 8f.3: config 2's sweep in the live stack's re-planning regime (40-point local paths, every 3rd "
                  "waypoint, re-planned when the planner's index passes 3: nodes/planner.py:26), full kernel build"),
}
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
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the bounded CPU sample")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle loop (reference functions composed; test infrastructure) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    workload, start, count, sim_steps = job
    import bgr_pathplanning_control_b200 as bgr
    from oracle import loop
    steer, accel, model, laps, _ = WORKLOADS[workload]
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
        steer, accel, model, laps, _ = WORKLOADS[workload]
        cores = len(os.sched_getaffinity(0))
        E = cores * episodes_per_core
        ta = bgr.tracks.stadium_oval(1024)
        params = (bgr.sweeps.pp_sweep_params(E) if workload == "pp_sweep" else bgr.sweeps.stanley_lqg_params(E))
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
    vals, descs, t_total, steps_total = [], [], 0.0, 0
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
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
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
# GPU arm
# ------------------------------------------------------------------------------------------------
def issue_roofline(ev, sim_steps_per_s, clocks):
    if not ev:
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


def run_b200(args):
    import torch
    import torch.distributed as dist

    import bgr_pathplanning_control_b200 as bgr
    from bgr_pathplanning_control_b200 import _abi as abi
    from bgr_pathplanning_control_b200 import roofline as rf

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "mpc":
        return run_mpc(args, bgr, abi, rf, torch, dist, rank, world, local, dev, barrier)
    if args.workload == "skidpad_mc":
        return run_skidpad_mc(args, bgr, abi, rf, torch, dist, rank, world, local, dev, barrier)

    steer, accel, model, laps, desc = WORKLOADS[args.workload]
    E, S = args.episodes, args.sim_steps
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta, device=local)
    cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=laps,
                            episode_id_base=rank * E, **WORKLOAD_EXTRA.get(args.workload, {}))
    # Input sets: each timed step consumes a DIFFERENT batch (rows of the sweep grid shifted by rank and set), so
    # the inputs touched over the timed region (n_sets x 112 B x E) exceed the 126 MB L2.
    n_sets = 4
    make = bgr.sweeps.pp_sweep_params if steer == "pure_pursuit" else \
        (lambda n, start: bgr.sweeps.stanley_lqg_params(n, start=start))
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

    for i in range(args.warmup):
        step(i)
    barrier()
    launches0 = bgr.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    steps_done = torch.zeros((), dtype=torch.float64, device=dev)
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for i in range(args.steps):
        ev[i][0].record()
        r = step(i)
        ev[i][1].record()
        steps_done += r.aggregate[abi.A_STEPS]
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = bgr.launch_count() - launches0
    ms_total = t_begin.elapsed_time(t_end)
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    kernel_ms_last = bgr.last_kernel_ms()
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    tot = steps_done.clone().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    total_sim_steps = float(tot.item())
    value = total_sim_steps / (ms_total * 1e-3)

    # ---- final metric gather + aggregate reduction: the ONLY cross-rank traffic (timed separately) ----
    gather_ms = None
    if world > 1:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bgr.sharding.gather_results(out.metrics, out.status, out.aggregate, E * world)   # warm-up: NCCL channels
        barrier()
        g0.record()
        m_all, s_all, agg_all = bgr.sharding.gather_results(out.metrics, out.status, out.aggregate, E * world)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)
        assert m_all.shape[0] == E * world and float(agg_all[abi.A_EPISODES]) == E * world

    # ---- e2e: host buffers through bgr_rollout_host (H2D + kernel + D2H inside the timed region) ----
    np_params = [t.numpy() for t in h_params]
    np_init = [t.numpy() for t in h_init]
    h_out = bgr.RolloutResult(torch.empty((E, abi.N_METRICS), dtype=torch.float32).pin_memory().numpy(),
                              torch.empty((E, abi.N_STATUS), dtype=torch.int32).pin_memory().numpy())
    for i in range(min(args.warmup, 2)):
        bgr.rollout(track, cfg, np_params[i % n_sets], np_init[i % n_sets], aggregate=True, out=h_out)
    barrier()
    e2e_steps = 0.0
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dbg = []
    for i in range(args.steps):
        td = time.perf_counter()
        r = bgr.rollout(track, cfg, np_params[i % n_sets], np_init[i % n_sets], aggregate=True, out=h_out)
        e2e_steps += float(r.aggregate[abi.A_STEPS])
        dbg.append((time.perf_counter() - td) * 1e3)
    if os.environ.get("BGR_BENCH_DEBUG"):
        print(f"[rank {rank}] e2e ms per step: " + " ".join(f"{x:.1f}" for x in dbg), file=sys.stderr, flush=True)
    e1.record()
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    e2e_value = float(tot.item()) / (float(t.item()) * 1e-3)
    h2d = E * (abi.N_PARAMS * 4 + abi.N_STATE * 8)
    d2h = E * (abi.N_METRICS * 4 + abi.N_STATUS * 4) + abi.N_AGG * 8

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        fp32_tflops, fp32_ginst = bgr.fp32_peak_probe(local)
        flops_step = rf.flops_per_step(steer, accel, model)
        k_ms = float(np.mean(kernel_ms))
        steps_per_launch = total_sim_steps / world / args.steps
        achieved_tflops = steps_per_launch * flops_step / (k_ms * 1e-3) / 1e12
        bytes_launch = E * rf.bytes_per_episode()
        hbm_gbs = bytes_launch / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 laws + f64 integrators (x, y, yaw, v) and metric sums", "data": "synthetic",
            "config": {"workload": desc, "episodes_per_gpu": E, "sim_steps": S, "steer": steer, "accel": accel,
                       "model": model, "track": "stadium_oval n=1024", "parallelism": f"episode sharding x{world}",
                       "l2": f"{n_sets} rotating input sets ({n_sets * h2d / 1e6:.0f} MB) > 126 MB L2, no flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "bgr_rollout_host (pinned numpy in, numpy out)"},
            "gpu_launches": int(launches),
            "roofline": {
                "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_tflops, "unit": "TFLOP/s",
                "frac": achieved_tflops / fp32_tflops,
                "traffic": (NCU_EVIDENCE.get(args.workload) or {}).get("dram_bytes_per_launch"),
                "peak_source": "bgr_fp32_peak_probe (FFMA chains, measured in this run); MEASURED_PEAKS.json has no fp32 row",
                "flops_per_sim_step": flops_step, "kernel_ms_avg": k_ms, "kernel_ms_last": kernel_ms_last,
                "sim_steps_per_launch": steps_per_launch,
                # what ncu measured for this kernel build (one --set full capture per change, profiles/r1_final/):
                # the kernel is bound by INSTRUCTION ISSUE, not by DRAM (0.08 %) or tensor (0 %)
                "ncu": NCU_EVIDENCE.get(args.workload),
                # the roofline that actually binds: warp-instructions issued per second against the issue peak of the
                # chip (148 SMs x 4 schedulers x 1 instruction per cycle at the SM clock seen under load).  Live steps/s
                # x the per-step instruction count ncu measured for this build (profiles/r1_final).
                "issue": issue_roofline(NCU_EVIDENCE.get(args.workload), steps_per_launch / (k_ms * 1e-3), clocks),
                "hbm": {"achieved": hbm_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / peaks["hbm_gbs"],
                        "bytes_per_launch": bytes_launch, "peak_source": peaks_src},
            },
        }
        if gather_ms is not None:
            line["gather_ms"] = gather_ms
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_sample(args.workload, S, args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "cpu": cpu_model(), "c_twin": c_twin_sample(args.workload, S)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_skidpad_mc(args, bgr, abi, rf, torch, dist, rank, world, local, dev, barrier):
    """BASELINE config 5 (secondary bench line): skidpad figure-eight, mixed-controller Monte Carlo with Philox pose
    noise and perturbed cones; episodes are grouped by controller kind on the host (three homogeneous launches per
    step); the final per-episode metric gather + aggregate reduction goes through NCCL."""
    kinds = [("pure_pursuit", "pid", "kinematic"), ("stanley", "pid", "kinematic"), ("stanley", "lqg", "dynamic")]
    per = max(256, (args.episodes // 2) // len(kinds) // 256 * 256)       # 512 Ki episodes per GPU in total by default
    S = min(args.sim_steps, 1500)
    ta = bgr.tracks.skidpad(2048)
    track = bgr.Track.from_arrays(ta, device=local, cone_reach=4.0)
    noise = dict(sigma_xy=0.05, sigma_yaw=0.01, sigma_v=0.1, sigma_cone=0.05, noise_seed=20261022, hit_radius=0.5)
    for k in os.environ.get("BGR_MC_OFF", "").split(","):      # diagnostic: switch parts of the Monte Carlo off
        if k == "cone_noise":
            noise["sigma_cone"] = 0.0
        elif k == "cones":
            noise["hit_radius"] = 0.0
        elif k == "pose_noise":
            noise.update(sigma_xy=0.0, sigma_yaw=0.0, sigma_v=0.0)
    jobs = []
    for g, (steer, accel, model) in enumerate(kinds):
        lo = (rank * len(kinds) + g) * per
        # (two-phase execution, cfg.split_steps, does not pay here: measured, the kernel time of the Stanley kinds is the
        # serial latency of their few 1 500-step episodes, ~5.6 us per step for a lone warp, packed or not)
        split = max(0, int(os.environ.get("BGR_MC_SPLIT", "0")))
        cfg = bgr.RolloutConfig(steer=steer, accel=accel, model=model, max_steps=S, laps=1, episode_id_base=lo,
                                split_steps=split, **noise)
        init = bgr.initial_states(ta, per, index=0, v0=5.0 if model == "dynamic" else 0.0)
        init[:, abi.YAW] = 0.0
        params = bgr.sweeps.monte_carlo_params(per, start=lo)
        out = bgr.RolloutResult(torch.empty((per, abi.N_METRICS), dtype=torch.float32, device=dev),
                                torch.empty((per, abi.N_STATUS), dtype=torch.int32, device=dev), None,
                                torch.empty(abi.N_AGG, dtype=torch.float64, device=dev))
        jobs.append((cfg, torch.from_numpy(params).to(dev), torch.from_numpy(init).to(dev), out, params, init))

    # the three kinds run on concurrent streams, the latency-bound Stanley kinds first (engine.rollout_groups);
    # BGR_MC_SERIAL=1 launches them back to back on one stream instead
    serial = os.environ.get("BGR_MC_SERIAL", "0") == "1"

    def step():
        if serial:
            return [bgr.rollout(track, j[0], j[1], j[2], aggregate=True, out=j[3]) for j in jobs]
        return bgr.rollout_groups(track, [(j[0], j[1], j[2]) for j in jobs], aggregate=True,
                                  outs=[j[3] for j in jobs], order=[2, 1, 0])

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = bgr.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    done = torch.zeros((), dtype=torch.float64, device=dev)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        for r in step():
            done += r.aggregate[abi.A_STEPS]
    t1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = bgr.launch_count() - launches0
    t = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    tot = done.clone().reshape(1)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_total, total_steps = float(t.item()), float(tot.item())
    gather_ms = None
    if world > 1:
        m = torch.cat([j[3].metrics for j in jobs])
        s = torch.cat([j[3].status for j in jobs])
        agg = torch.from_numpy(bgr.sharding.combine_aggregates([j[3].aggregate.cpu().numpy() for j in jobs])).to(dev)
        bgr.sharding.gather_results(m, s, agg, m.shape[0] * world)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        m_all, s_all, agg_all = bgr.sharding.gather_results(m, s, agg, m.shape[0] * world)
        g1.record()
        barrier()
        gather_ms = g0.elapsed_time(g1)
        assert float(agg_all[abi.A_EPISODES]) == m.shape[0] * world
    # e2e: host buffers
    w0 = time.perf_counter()
    e2e_steps = 0.0
    for _ in range(args.steps):
        if serial:
            rs = [bgr.rollout(track, j[0], j[4], j[5], aggregate=True) for j in jobs]
        else:
            rs = bgr.rollout_groups(track, [(j[0], j[4], j[5]) for j in jobs], aggregate=True, order=[2, 1, 0])
        e2e_steps += sum(float(r.aggregate[abi.A_STEPS]) for r in rs)
    e2e_s = time.perf_counter() - w0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    se = torch.tensor([e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(se, op=dist.ReduceOp.SUM)
    if rank == 0:
        fp32_tflops, _ = bgr.fp32_peak_probe(local)
        flops = np.mean([rf.flops_per_step(*k) for k in kinds])
        ach = total_steps / world * flops / (ms_total * 1e-3) / 1e12
        hits = sum(float(j[3].aggregate[abi.A_CONE_HITS]) for j in jobs)
        line = {
            "metric": METRIC, "value": total_steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 laws + f64 integrators (x, y, yaw, v) and metric sums",
            "data": "synthetic",
            "config": {"workload": "config 5: skidpad figure-eight, mixed-controller Monte Carlo with noisy pose / cone "
                                   "perturbations (Philox4x32-10), NCCL metric gather", "episodes_per_gpu": per * len(kinds),
                       "kinds": ["+".join(k) for k in kinds], "sim_steps_cap": S, "track": "skidpad n=2048, 45 cones",
                       "cone_hits_last_step_rank0": hits, "parallelism": f"episode sharding x{world}",
                       "launch": "kinds back to back on one stream" if serial else
                                 "kinds on three concurrent streams, latency-bound Stanley kinds first",
                       "l2": "state lives in registers; 112 B in + 88 B out per episode"},
            "clocks": clocks,
            "e2e": {"value": float(se.item()) / float(te.item()), "unit": UNIT,
                    "h2d_bytes_per_step": per * len(kinds) * 112, "d2h_bytes_per_step": per * len(kinds) * 88 + 3 * 64},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp32", "achieved": ach, "peak": fp32_tflops, "unit": "TFLOP/s", "frac": ach / fp32_tflops,
                         "traffic": None, "note": "full kernel build (cones + noise): 128 registers, 2 CTAs per SM"},
        }
        if gather_ms is not None:
            line["gather_ms"] = gather_ms
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_mpc(args, bgr, abi, rf, torch, dist, rank, world, local, dev, barrier):
    """BASELINE config 4 (secondary bench line): 16k episodes x 4096 candidates x horizon 30."""
    E, K, H = 16384, 4096, 30
    ta = bgr.tracks.stadium_oval(1024)
    track = bgr.Track.from_arrays(ta, device=local)
    cfg = bgr.MpcConfig(horizon=H, dt=0.05, explicit_rollout=bool(int(os.environ.get("BGR_MPC_EXPLICIT", "0"))))
    states, ref_start = bgr.sweeps.mpc_states(ta, E, start=rank * E)
    bank = bgr.candidate_bank(K, H)
    d_state, d_ref, d_bank = (torch.from_numpy(a).to(dev) for a in (states, ref_start, bank))
    for _ in range(args.warmup):
        bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
    barrier()
    kernel_ms = bgr.last_kernel_ms()
    # One evaluation is three short kernels (~0.07 ms on the device) behind ~0.09 ms of Python + launch work: the call
    # is captured ONCE into a CUDA graph (stream-ordered scratch allocation, the kernels, the free) and replayed per
    # step.  BGR_MPC_GRAPH=0, or a capture failure, falls back to plain calls.
    graph, launch_mode = None, "one API call per step"
    if os.environ.get("BGR_MPC_GRAPH", "1") != "0":
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                res = bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
            g.replay()
            torch.cuda.synchronize()
            eager = bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
            torch.cuda.synchronize()
            if not (torch.equal(res.best_idx, eager.best_idx) and torch.equal(res.best_cost, eager.best_cost)):
                raise RuntimeError("graph replay disagrees with the eager call")
            graph, launch_mode = g, "CUDA graph replay of one bgr_mpc_eval call"
        except Exception as exc:             # noqa: BLE001 -- the eager path is always available
            launch_mode = f"one API call per step (graph capture unavailable: {str(exc)[:80]})"
            torch.cuda.synchronize()
    barrier()
    launches0 = bgr.launch_count()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        if graph is not None:
            graph.replay()
        else:
            res = bgr.mpc_evaluate(track, cfg, d_state, d_ref, d_bank)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    # kernels launched inside the timed region: counted by the library per API call; a graph replay re-launches the
    # kernels captured from one call (prepare, factored evaluator, finalize)
    per_call = 1 if cfg.explicit_rollout else 2
    timed_launches = (per_call + 1) * args.steps if graph is not None else bgr.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    units = float(E) * K * H * world * args.steps
    w0 = time.perf_counter()
    for _ in range(args.steps):
        bgr.mpc_evaluate(track, cfg, states, ref_start, bank)
    e2e_s = time.perf_counter() - w0
    if rank == 0:
        fp32_tflops, _ = bgr.fp32_peak_probe(local)
        if cfg.explicit_rollout:
            # explicit stage rollout: 29 algorithmic flops per candidate-step on the FP32 pipes
            ach = float(E) * K * H * rf.FLOPS["mpc_candidate_step"] / (kernel_ms * 1e-3) / 1e12
            roof = {"bound": "fp32", "achieved": ach, "peak": fp32_tflops, "unit": "TFLOP/s", "frac": ach / fp32_tflops,
                    "traffic": None, "kernel_ms_last": kernel_ms}
        else:
            # factored evaluator: 3 fp64 FMAs + 1 add per (episode, candidate) pair = 7 flops on the FP64 CUDA-core
            # pipe (64 FMA lanes per SM); nominal peak = 148 SMs x 64 x 2 x 1.965 GHz
            ach = float(E) * K * 7.0 / (kernel_ms * 1e-3) / 1e12
            peak64 = 148 * 64 * 2 * 1.965e9 / 1e12
            roof = {"bound": "fp64", "achieved": ach, "peak": peak64, "unit": "TFLOP/s", "frac": ach / peak64,
                    "traffic": None, "kernel_ms_last": kernel_ms, "peak_source": "nominal 148 SM x 64 lanes x 2 x 1.965 GHz",
                    "note": "the explicit 30-stage rollout needs 29 flops per candidate-step; the factored form needs 7 per "
                            "PAIR -- candidate-steps/s is the BASELINE metric, this roofline is of the kernel that runs"}
        print(json.dumps({
            "metric": "MPC candidate-steps/sec (episodes x candidates x horizon)", "value": units / (ms * 1e-3),
            "unit": "candidate-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config 4: MPC candidate-sequence rollout/cost evaluation, 16384 episodes x 4096 "
                                   "candidates x horizon 30", "solves_per_s": E * world * args.steps / (ms * 1e-3),
                       "evaluator": "explicit fp32 stage rollout" if cfg.explicit_rollout else
                       "factored (prefix-sum statistics, 3 fp64 FMAs + 1 add per episode-candidate pair)",
                       "launch": launch_mode},
            "e2e": {"value": float(E) * K * H * args.steps / e2e_s, "unit": "candidate-steps/s",
                    "h2d_bytes_per_step": states.nbytes + ref_start.nbytes + bank.nbytes, "d2h_bytes_per_step": E * 16},
            "gpu_launches": int(timed_launches),
            "roofline": roof,
        }), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
