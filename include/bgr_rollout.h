/*
 * bgr_rollout.h -- C ABI of the B200-native batched closed-loop rollout engine.
 *
 * This is the drop-in boundary for the ONE hot path of grimgrimberg/BGR_PathPlanning_Control that
 * this library replaces: tracking law -> bicycle-model step -> lap-metric accumulation, for many
 * independent episodes at once.  The reference has no FFI of its own (it is
 * 100 % Python, duck typed); each entry point below names the reference
 * interface it stands in for.  INTEGRATION.md shows the ctypes binding a reference maintainer adds.
 *
 * Conventions
 *   - every function returns an int status: BGR_OK (0) or a negative BGR_ERR_*; nothing throws
 *     across the ABI; bgr_last_error_string() describes the last failure on the calling thread.
 *   - pointers named d_* are DEVICE pointers owned by the caller (e.g. torch tensors); pointers named
 *     h_* are HOST pointers.  The library allocates nothing it does not free (stream ordered) before
 *     returning, except: the opaque handles (track: with the LQG gain tables it caches on first use; MPC bank: with
 *     the candidate statistics it caches per cost configuration) and, per calling host thread, the host-buffer
 *     pipeline of bgr_rollout_host[_async] -- four internal streams, a pool of events and a ring of three device
 *     buffer sets sized for the largest batch seen (kept until the process ends: nothing is torn down from a
 *     thread-exit destructor, where the CUDA runtime may already be unloading) -- plus the CUDA event pair of
 *     bgr_last_kernel_ms.
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  Calls are asynchronous on
 *     that stream unless the name ends in _host (those synchronise the stream before returning).
 *   - handles are logically immutable after creation (the internal LQG table cache is mutex protected);
 *     concurrent calls on different streams / host threads are allowed.  A handle must outlive every
 *     asynchronous call that uses it.
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef BGR_ROLLOUT_H_
#define BGR_ROLLOUT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden: only this ABI is exported */
#endif

#define BGR_ABI_VERSION 3

/* ---- status codes ---- */
#define BGR_OK 0
#define BGR_ERR_BAD_ARG (-1)
#define BGR_ERR_CUDA (-2)
#define BGR_ERR_TOO_LARGE (-3)   /* track does not fit the shared-memory staging budget */
#define BGR_ERR_ALLOC (-4)
#define BGR_ERR_UNSUPPORTED (-5)

/* ---- law / model selectors (reference classes they stand for) ---- */
#define BGR_STEER_PURE_PURSUIT 0 /* PurePursuitController.compute_steering  core/controllers.py:193-230 */
#define BGR_STEER_STANLEY 1      /* StanleyController.compute_steering (live def) core/controllers.py:336-385 */
#define BGR_ACCEL_PID 0          /* AccelerationPIDController.compute_acceleration core/controllers.py:32-47 */
#define BGR_ACCEL_LQG 1          /* LQGAccelerationController.compute_acceleration core/controllers.py:117-138 */
#define BGR_MODEL_KINEMATIC 0    /* oracle-defined rear-axle kinematic bicycle (no reference code) */
#define BGR_MODEL_DYNAMIC 1      /* DynamicBicycleModel.state_equations core/data/car_state.py:178-223 */
#define BGR_PRECISION_FP32 0     /* production: fp32 laws, fp64 integrators for x, y, yaw, v */
#define BGR_PRECISION_FP64 1     /* audit mode: every operation in fp64, reference operation order */

/* ---- per-episode parameter row: float32[BGR_N_PARAMS] (defaults = vehicle_config.py:4-39) ---- */
#define BGR_P_LF 0            /* LOOKAHEAD_DISTANCE      vehicle_config.py:37 */
#define BGR_P_KP 1            /* kp_accel                :19 */
#define BGR_P_KI 2            /* ki_accel                :20 */
#define BGR_P_KD 3            /* kd_accel                :21 */
#define BGR_P_TARGET_SPEED 4  /* TARGET_SPEED            :14 */
#define BGR_P_K_EXPO 5        /* k_expo                  :24 */
#define BGR_P_K_STANLEY 6     /* StanleyController.k      core/controllers.py:325,585 */
#define BGR_P_K_SOFT 7        /* StanleyController.k_soft core/controllers.py:325,585 */
#define BGR_P_MU 8            /* mu                      vehicle_config.py:13 */
#define BGR_P_C_F 9           /* c_f                     :11 */
#define BGR_P_C_R 10          /* c_r                     :12 */
#define BGR_P_MASS 11         /* m                       :7 */
#define BGR_P_I_Z 12          /* I_z                     :8 */
#define BGR_N_PARAMS 16       /* 13..15 reserved (must be 0); rows are 64 B = one 512-bit line */

/* ---- per-episode vehicle state row: double[BGR_N_STATE]  (State, core/data/car_state.py:103-110) ---- */
#define BGR_X 0
#define BGR_Y 1
#define BGR_YAW 2
#define BGR_V 3
#define BGR_R 4
#define BGR_BETA 5
#define BGR_N_STATE 6

/* ---- per-episode metric row: float32[BGR_N_METRICS]
 *      (preformance_analysis/metrics_calculator.py:1-8 over the rows of simulation_logger.py:8-10) ---- */
#define BGR_M_LAT_MEAN 0   /* calculate_lateral_error()[0] */
#define BGR_M_LAT_STD 1    /* calculate_lateral_error()[1]  (pandas std, ddof = 1; NaN for one row) */
#define BGR_M_HEAD_MEAN 2  /* calculate_heading_error()[0] */
#define BGR_M_HEAD_STD 3   /* calculate_heading_error()[1] */
#define BGR_M_LAP_TIME 4   /* calculate_lap_time(): (rows - 1) * dt */
#define BGR_M_MAX_ABS_LAT 5
#define BGR_M_X 6          /* final state, rounded to float32 */
#define BGR_M_Y 7
#define BGR_M_YAW 8
#define BGR_M_V 9
#define BGR_M_R 10
#define BGR_M_BETA 11
#define BGR_N_METRICS 12

/* ---- per-episode status row: int32[BGR_N_STATUS] ---- */
#define BGR_S_FLAGS 0        /* BGR_ST_* bits */
#define BGR_S_STEPS 1        /* logged rows == executed steps */
#define BGR_S_TARGET_IND 2   /* final target index (PP: on the unrolled path) */
#define BGR_S_NEAREST_IND 3  /* final nearest index on the base lap */
#define BGR_S_CONE_HITS 4    /* distinct cones hit */
#define BGR_S_LAPS 5
#define BGR_S_NEAR_TIES 6    /* steps with a threshold decision closer than tie_eps (audit channel) */
#define BGR_S_FIRST_TIE 7    /* first such step, -1 if none */
#define BGR_S_FALLBACKS 8    /* nearest searches that needed the exact global scan (fp32 kernels: an episode's first
                                search descends from waypoint 0 and only scans when that takes more than 16 moves) */
#define BGR_N_STATUS 10

#define BGR_ST_PATH_END 1
#define BGR_ST_LAPS_DONE 2
#define BGR_ST_TIMEOUT 4
#define BGR_ST_NONFINITE 8

/* ---- optional per-step trajectory dump row: double[BGR_N_TRAJ], layout [episode][step][field] ---- */
#define BGR_T_X 0
#define BGR_T_Y 1
#define BGR_T_YAW 2
#define BGR_T_V 3
#define BGR_T_R 4
#define BGR_T_BETA 5
#define BGR_T_STEER 6
#define BGR_T_ACCEL 7
#define BGR_T_TARGET_IND 8
#define BGR_T_NEAREST_IND 9
#define BGR_T_E_CT 10
#define BGR_T_THETA_E 11
#define BGR_N_TRAJ 12

/* ---- per-device aggregate row: double[BGR_N_AGG] (fixed-order reduction over the call's episodes) ---- */
#define BGR_A_EPISODES 0
#define BGR_A_STEPS 1
#define BGR_A_SUM_LAT_MEAN 2
#define BGR_A_SUM_LAT_STD 3
#define BGR_A_SUM_LAP_TIME 4
#define BGR_A_MIN_LAP_TIME 5
#define BGR_A_CONE_HITS 6
#define BGR_A_FINISHED 7     /* episodes with PATH_END or LAPS_DONE */
#define BGR_N_AGG 8

typedef struct bgr_track bgr_track_t; /* opaque, device resident */

/* Centreline + cones as the reference's controllers receive them: cx, cy (nodes/planner.py:88) and
 * curve (nodes/planner.py:89), all float64 host arrays of length n (the BASE LAP; must not repeat
 * the first point at the end when closed). */
typedef struct bgr_track_desc {
  const double* h_x;
  const double* h_y;
  const double* h_kappa;
  int32_t n;
  int32_t closed;          /* 1: nearest search and lap counter wrap around; 0: open path */
  const double* h_cones_xy; /* interleaved x, y; may be NULL */
  int32_t n_cones;
  int32_t reserved;
  double cone_reach;       /* cones within cone_reach of a waypoint are that waypoint's candidates;
                              must be >= hit_radius + 6 * sigma_cone + BGR_CONE_GUARD (checked per rollout) */
} bgr_track_desc_t;
#define BGR_CONE_GUARD 2.0 /* [m] car-to-nearest-waypoint distance up to which candidate lists are exact */
#define BGR_MAX_CONES 512

typedef struct bgr_track_info {
  int32_t n, closed, n_cones, device;
  int64_t smem_bytes_fp32, smem_bytes_fp64;
  double guard_min, guard_median; /* exactness-guard radii of the windowed nearest search [m] */
  double ds_min, ds_max;          /* waypoint spacing [m] */
  int64_t rival_total;            /* total length of the rival lists of the nearest search */
  int32_t rival_max;              /* longest list */
  int32_t rival_unlisted;         /* waypoints without a usable list (exact scan outside their guard) */
  int32_t twin_count;             /* waypoints with a twin triple (a second lap on the same line) */
  int32_t reserved;
} bgr_track_info_t;

#define BGR_CFG_AUDIT 1 /* run the full kernel build so that the near-tie audit channel
                           (BGR_S_NEAR_TIES / BGR_S_FIRST_TIE) is filled even without cones/noise/traj */
#define BGR_CFG_STANLEY_SHADOWED 2 /* with BGR_STEER_STANLEY: the SHADOWED (dead-code) class body,
                           core/controllers.py:234-324 -- adaptive gain k / (v + k_soft), steering-rate limit, low-pass
                           filter, carried previous_steering.  A labelled extra: the live law is :336-385. */

#define BGR_CFG_DYNAMIC_SCHEDULE 4 /* full-build rollouts of more than one wave of episodes: a one-wave grid whose lanes
                           pull their next row from a queue the moment their episode ends, instead of one episode per
                           thread.  For batches with heavy-tailed episode lengths; output rows are bit-identical either
                           way.  Off by default (measured neutral on BASELINE config 5). */

#define BGR_CFG_KEEP_ORDER 8 /* fp32 rollouts that can end before the step cap (open path, lap target) launch their
                           episodes in order of target speed, slowest first, so that a warp holds episodes of like length
                           (a bucket sort of the rows' BGR_P_TARGET_SPEED on the device, part of the call).  Thread t then
                           works on row perm[t]; every output row is the same as without it.  This bit keeps the launch
                           in row order. */

typedef struct bgr_rollout_cfg {
  int32_t steer_kind, accel_kind, model_kind, precision;
  int32_t max_steps;       /* step cap (legacy loop: T / dt, old/animation_loop.py:16-17,71) */
  int32_t laps;            /* PP: base lap tiled this many times (monotone index, no wrap) */
  int32_t laps_target;     /* stop after this many completed laps; 0 = off */
  int32_t flags;           /* BGR_CFG_* bits */
  double dt;               /* vehicle_config.py:39 */
  double wheelbase;        /* WB = l_f + l_r, vehicle_config.py:16 */
  double l_f, l_r;         /* :9-10 */
  double max_steer;        /* :35 */
  double max_accel;        /* :30 */
  double max_decel;        /* :31 (negative) */
  double hit_radius;       /* 0 = cone hits off */
  double tie_eps;          /* [m] near-tie audit threshold; 0 = default 2e-5 */
  /* observation noise (config 5); all sigmas 0 = off.  Philox4x32-10, key = (seed, episode id) */
  double sigma_xy, sigma_yaw, sigma_v, sigma_cone;
  uint32_t noise_seed;
  uint32_t split_steps;    /* 0 = one launch.  > 0: two-phase execution -- a first launch capped at split_steps, then the
                              episodes still running are re-run from scratch, densely packed, to max_steps.  Results
                              are bit-identical; pays off when most episodes end early (laps_target, open paths) */
  int64_t episode_id_base; /* global id of episode 0 of this call (multi-GPU shards) */
  /* BGR_CFG_STANLEY_SHADOWED only (core/controllers.py:234): 0 = the reference defaults deg2rad(30) rad/s and 0.5 */
  double stanley_max_steer_rate;
  double stanley_alpha;
  /* ---- receding-horizon re-planning replay (BGR_STEER_PURE_PURSUIT only; replan_horizon = 0: off) -- the regime the
   *      live stack actually drives in (nodes/planner.py:26,62,72-77 feeding nodes/controller.py:56-62): the controller
   *      sees a SHORT local path; the planner keeps its own look-ahead index on it (scan :72-77 with
   *      conf.LOOKAHEAD_DISTANCE) and plans a new path, index reset to 0 (:62), whenever that index has passed
   *      replan_after (:26).  fsd-path-planning itself is out of scope: the local path planned from a car pose is the
   *      WINDOW of the handle's centreline that starts at the waypoint nearest to the (observed) car,
   *      local point k = waypoint j0 + k * replan_stride (mod n when closed, clipped to n - 1 when open).
   *      The reported target index (BGR_S_TARGET_IND, BGR_T_TARGET_IND) is the controller's LOCAL index;
   *      open tracks end (BGR_ST_PATH_END) when the target is the last waypoint; cfg.laps is ignored. */
  int32_t replan_horizon;   /* M >= 2 points per local path (about 40 live) */
  int32_t replan_stride;    /* >= 1; 0 = 1 */
  int32_t replan_after;     /* re-plan when the planner's index > this; reference value 3 (nodes/planner.py:26) */
  int32_t reserved2;
  double planner_lookahead; /* the planner's own look-ahead (nodes/planner.py:75); 0 = 5.5 (vehicle_config.py:37) */
  /* BGR_CFG_STANLEY_SHADOWED only: the optional look-ahead advance of the dead-code variant (core/controllers.py:
   * 276-281) -- from the nearest waypoint the index moves forward while it is closer than this distance and below
   * n - 1; the error terms of the law are taken there.  0 = off (the constructor default, :234). */
  double stanley_lookahead;
} bgr_rollout_cfg_t;

#define BGR_MPC_EXPLICIT_ROLLOUT 1 /* bgr_mpc_cfg_t.flags: score candidates by the explicit stage-by-stage rollout
                                     even in fp32 (default: the algebraically factored evaluator) */
typedef struct bgr_mpc_cfg {
  int32_t horizon;         /* MPCController.N  core/controllers.py:472,587 */
  int32_t precision;       /* BGR_PRECISION_FP64 = explicit rollout in the reference's operation order */
  double dt;               /* MPCController.dt */
  double wheelbase;        /* conf.WB, :525 */
  double target_speed;     /* conf.TARGET_SPEED, :540 */
  double q[4];             /* diag Q :483 */
  double r[2];             /* diag R :484 */
  double max_accel, max_decel, max_steer; /* :565-566 */
  int32_t flags;           /* BGR_MPC_* */
  int32_t reserved;
} bgr_mpc_cfg_t;

/* ---- library ---- */
int bgr_abi_version(void);
const char* bgr_last_error_string(void);
int bgr_device_count(int* out_count);
/* Identity of this build: hash of the kernel / ABI sources and of the compile-time knobs it was built with.  bench.py
 * prints it and matches it against the build id stored with the ncu summaries under profiles/, so that profile
 * evidence of an older kernel is never quoted for a newer one. */
const char* bgr_build_id(void);
/* Streams for callers that hold no CUDA runtime of their own (the *_async entry points): a non-blocking stream on
 * `device`, cudaStreamSynchronize, cudaStreamDestroy.  Any cudaStream_t the caller already owns works as well. */
int bgr_stream_create(int device, void** out_stream);
int bgr_stream_synchronize(void* stream);
int bgr_stream_destroy(void* stream);

/* ---- track handle: replaces the (N,3) `path` ndarray + `curve` the reference hands its controllers
 *      every call (nodes/controller.py:56,62) ---- */
int bgr_track_create(const bgr_track_desc_t* desc, int device, bgr_track_t** out_track);
void bgr_track_destroy(bgr_track_t* track);
int bgr_track_info(const bgr_track_t* track, bgr_track_info_t* out_info);
/* host copies of derived per-waypoint tables, each double[n]: path heading (core/controllers.py:361-368)
 * and the exactness-guard radius of the nearest search. */
int bgr_track_tables(const bgr_track_t* track, double* h_path_yaw, double* h_guard_radius);

/* ---- batched closed-loop rollout: replaces the per-step Python loop
 *      Controller.update (nodes/controller.py:28-94) / old/animation_loop.py:71-136
 *      + State.update (core/data/car_state.py:118-125) + metrics_calculator (:1-8). ---- */
int bgr_rollout(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                const float* d_params,      /* [n_episodes][BGR_N_PARAMS] */
                const double* d_init_state, /* [n_episodes][BGR_N_STATE] */
                float* d_metrics,           /* [n_episodes][BGR_N_METRICS] */
                int32_t* d_status,          /* [n_episodes][BGR_N_STATUS] */
                double* d_traj,             /* NULL or [n_episodes][max_steps][BGR_N_TRAJ] */
                double* d_aggregate,        /* NULL or [BGR_N_AGG] */
                void* stream);

/* Same call with HOST buffers: H2D of params/init, the rollout, D2H of metrics/status/aggregate, and a
 * stream synchronise, all inside.  This is what a reference-side caller holding numpy arrays uses. */
int bgr_rollout_host(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                     const float* h_params, const double* h_init_state, float* h_metrics,
                     int32_t* h_status, double* h_traj, double* h_aggregate, void* stream);

/* bgr_rollout_host WITHOUT the final synchronise: everything (chunked H2D, kernels, D2H) is queued and the call
 * returns; the results are valid after bgr_stream_synchronize(stream) (or any later synchronising call on that
 * stream).  The host buffers must stay alive and untouched until then, and should be pinned (cudaHostAlloc /
 * torch pin_memory) -- pageable memory makes the copies synchronous.  Consecutive calls from one host thread
 * pipeline: the first upload of call i + 1 runs under the last kernel and download of call i.  h_traj is not
 * supported here (BGR_ERR_UNSUPPORTED): a trajectory dump is a debugging aid, not a streaming workload. */
int bgr_rollout_host_async(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                           const float* h_params, const double* h_init_state, float* h_metrics,
                           int32_t* h_status, double* h_aggregate, void* stream);

/* ---- one control step for n independent vehicles (the scalar drop-in classes call this with n = 1):
 *      compute_steering (core/controllers.py:193 / :336), curvature lookup (nodes/controller.py:62),
 *      compute_acceleration (:32 / :117), State.update (core/data/car_state.py:118).
 *      `phases` selects what runs: BGR_PHASE_*.  HOST pointers; synchronous. ---- */
#define BGR_PHASE_STEER 1
#define BGR_PHASE_ACCEL 2
#define BGR_PHASE_MODEL 4
#define BGR_PHASE_ERRORS 8
#define BGR_PHASE_KAPPA_GIVEN 16 /* curvature comes from h_cmd[.][BGR_CMD_KAPPA], not from the lookup:
                                    compute_acceleration(current_speed, curvature) takes it as an argument */
/* command row: double[BGR_N_CMD] */
#define BGR_CMD_STEER 0
#define BGR_CMD_ACCEL 1
#define BGR_CMD_KAPPA 2
#define BGR_N_CMD 4
/* controller memory row: double[BGR_N_CTRL] */
#define BGR_C_PID_INTEGRAL 0
#define BGR_C_PID_LAST_INPUT 1
#define BGR_C_PID_HAS_LAST 2
#define BGR_C_LQG_XHAT0 3
#define BGR_C_LQG_XHAT1 4
#define BGR_C_LQG_A_PREV 5
#define BGR_C_STEP 6         /* steps this episode has taken so far: index into the shared Kalman-gain sequence (LQG)
                                and counter of the observation-noise stream, so that N stepped calls that carry their
                                controller row draw the same noise as one N-step bgr_rollout */
#define BGR_C_PREV_STEER 7   /* shadowed Stanley: previous_steering (core/controllers.py:250,319) */
#define BGR_N_CTRL 8
/* step output row: double[BGR_N_OUT] */
#define BGR_O_STEER 0
#define BGR_O_ACCEL 1
#define BGR_O_V_DESIRED 2    /* .maxspeed (PID, nodes/controller.py:71) / .v_desired (LQG) */
#define BGR_O_KAPPA 3
#define BGR_O_E_CT 4
#define BGR_O_THETA_E 5
#define BGR_O_NEAREST_IND 6
#define BGR_N_OUT 8
int bgr_step_host(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int32_t phases, int64_t n,
                  const double* h_params,  /* [n][BGR_N_PARAMS] -- fp64 rows: the reference's scalars keep every bit */
                  double* h_state,         /* [n][BGR_N_STATE] in/out (out only if BGR_PHASE_MODEL) */
                  int32_t* h_target_ind,   /* [n] in/out */
                  double* h_ctrl,          /* [n][BGR_N_CTRL] in/out */
                  const double* h_cmd,     /* NULL, or [n][BGR_N_CMD]: (steer, accel) commands to use for the
                                              phases that are switched off, and the given curvature */
                  double* h_out,           /* [n][BGR_N_OUT] */
                  void* stream);

/* ---- LQG design shared by all episodes: LQR gain (control.dlqr call, core/controllers.py:97) and the
 *      data-independent Kalman gain sequence (:145,151,154,160).  Host math, no device work. ---- */
int bgr_lqg_design(double dt, int32_t n_steps, double* h_lqr_gain /*[2]*/, double* h_kf_gains /*[n_steps][2] or NULL*/);
/* The estimate covariance LQGAccelerationController.P (core/controllers.py:109) after n_steps calls of
 * kalman_filter_update (:145,:160), row major 2 x 2; n_steps = 0 gives the initial identity.  Same recursion as the
 * gain sequence above (it is data independent), for callers that inspect ctrl.P. */
int bgr_lqg_covariance(double dt, int32_t n_steps, double* h_P /*[4]*/);

/* ---- MPC candidate-sequence rollout + cost evaluation: replaces the cvxpy/OSQP solve inside
 *      MPCController.compute_control (core/controllers.py:486-568) by evaluating the reference's own
 *      cost (:545,:548) under its linearised dynamics (:522-527) for K candidate control sequences
 *      per episode and taking the cheapest feasible (:536) one. ---- */
int bgr_mpc_eval(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                 const double* d_state,      /* [n_episodes][4] = x, y, yaw, v */
                 const int32_t* d_ref_start, /* [n_episodes] first reference waypoint (path[0] of :499-500), in [0, n);
                                                an out-of-range DEVICE value is wrapped (closed track) or clamped (open),
                                                bgr_mpc_eval_host rejects it with BGR_ERR_BAD_ARG */
                 const int32_t* d_ref_len,   /* NULL (= path runs to the end of the base lap) or [n_episodes] */
                 const float* d_bank,        /* [K][horizon][2] = (accel, steer) candidate bank */
                 int32_t n_candidates,
                 float* d_best_u,            /* [n_episodes][2] = (acceleration, steering) as returned at :568 */
                 float* d_best_cost,         /* [n_episodes]; +inf when no candidate is feasible */
                 int32_t* d_best_idx,        /* [n_episodes]; -1 when none feasible */
                 float* d_costs,             /* NULL or [n_episodes][K] (+inf = infeasible) */
                 void* stream);
int bgr_mpc_eval_host(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                      const double* h_state, const int32_t* h_ref_start, const int32_t* h_ref_len,
                      const float* h_bank, int32_t n_candidates, float* h_best_u, float* h_best_cost,
                      int32_t* h_best_idx, float* h_costs, void* stream);

/* ---- device-resident candidate bank: upload the bank ONCE, score it for many batches of states.  The handle also
 *      keeps the per-candidate prefix-sum statistics of the factored evaluator for every (dt, R, q_v) it has been used
 *      with, so a call costs the episode coefficients + the pair scoring only; the host variant then moves 36 B per
 *      episode in and 16 B out instead of re-sending ~1 MB of bank per call.  The bank's horizon must equal
 *      cfg->horizon and it must live on the track's device. ---- */
typedef struct bgr_mpc_bank bgr_mpc_bank_t; /* opaque, device resident */
int bgr_mpc_bank_create(int device, const float* h_bank /* [K][horizon][2] */, int32_t n_candidates, int32_t horizon,
                        bgr_mpc_bank_t** out_bank);
void bgr_mpc_bank_destroy(bgr_mpc_bank_t* bank);
int bgr_mpc_eval_bank(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                      const double* d_state, const int32_t* d_ref_start, const int32_t* d_ref_len,
                      bgr_mpc_bank_t* bank, float* d_best_u, float* d_best_cost, int32_t* d_best_idx,
                      float* d_costs, void* stream);
int bgr_mpc_eval_bank_host(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                           const double* h_state, const int32_t* h_ref_start, const int32_t* h_ref_len,
                           bgr_mpc_bank_t* bank, float* h_best_u, float* h_best_cost, int32_t* h_best_idx,
                           float* h_costs, void* stream);

/* The same without the final synchronise: the uploads, kernels and downloads are queued on `stream` and the call
 * returns; the result arrays are valid after the caller has synchronised `stream` (bgr_stream_synchronize), and neither
 * they nor the inputs may be touched before.  Consecutive calls on one handle alternate between two device arenas, so
 * issuing call i + 1 on a SECOND stream before waiting for call i lets its upload run under call i's kernels -- a
 * streaming caller with two sets of pinned buffers then pays the kernels, not kernels + copies.  At most two calls in
 * flight per handle (a third waits for the arena it needs). */
int bgr_mpc_eval_bank_host_async(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                                 const double* h_state, const int32_t* h_ref_start, const int32_t* h_ref_len,
                                 bgr_mpc_bank_t* bank, float* h_best_u, float* h_best_cost, int32_t* h_best_idx,
                                 void* stream);

/* ---- batched system identification: identify_dynamics (old/test_runs_system_id.py:84-111) for every episode of a
 *      trajectory dump: least squares x_{k+1} = A x_k + B u_k + c over the logged rows, x = [x, y, yaw, v, yaw_rate,
 *      beta], u = [accel, steering] (LinearRegression with an intercept on hstack([X, U]) -> Y).
 *      AB row r of an episode = [A[r][0..5] | B[r][0..1] | c[r]].  ok = 0 (and NaN) for rank-deficient logs. ---- */
#define BGR_SYSID_COLS 9
int bgr_identify_dynamics(int64_t n_episodes, int32_t max_rows,
                          const double* d_traj,   /* [n_episodes][max_rows][BGR_N_TRAJ] as written by bgr_rollout */
                          const int32_t* d_rows,  /* [n_episodes] logged rows (BGR_S_STEPS) */
                          double* d_AB,           /* [n_episodes][6][BGR_SYSID_COLS] */
                          int32_t* d_ok,          /* [n_episodes] */
                          void* stream);
int bgr_identify_dynamics_host(int64_t n_episodes, int32_t max_rows, const double* h_traj, const int32_t* h_rows,
                               double* h_AB, int32_t* h_ok, void* stream);

/* ---- measurement helpers ---- */
/* FP32 FMA-chain microbenchmark: the roofline denominator for the step kernels (MEASURED_PEAKS.json has
 * only HBM and bf16).  Returns achieved TFLOP/s (2 flops per FMA) and lane-instructions/s. */
int bgr_fp32_peak_probe(int device, double* out_tflops, double* out_ginst_per_s);
/* The same with DFMA chains: the FP64 CUDA-core peak, denominator of the MPC evaluator's roofline. */
int bgr_fp64_peak_probe(int device, double* out_tflops);
/* Duration [ms] of the last rollout / mpc kernel launched by this thread on `stream`, measured with
 * CUDA events recorded on that stream around the kernel (synchronises the stream). */
int bgr_last_kernel_ms(double* out_ms);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t bgr_launch_count(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* BGR_ROLLOUT_H_ */
