/*
 * bgr_oracle.c -- plain-C fp64 twin of the CPU oracle (oracle/laws.py + oracle/loop.py).  TEST INFRASTRUCTURE.
 *
 * Same algorithm, same operation order, one episode at a time; OpenMP over episodes.  It exists so that parity
 * of the CUDA path can be checked on thousands of full-length episodes in seconds (the numpy oracle manages ~1e4
 * steps/s) and so that bench.py can quote a compiled CPU baseline beside the Python one.  It is itself pinned
 * against the numpy oracle (tests/test_oracle_c.py), which is pinned bit-exact against golden vectors frozen from
 * the reference's own code (tests/golden/reference_kat.json).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it (oracle/__init__.py).
 *
 * Reference lines restated (relative to /root/reference):
 *   pure pursuit           core/controllers.py:193-230
 *   live Stanley           core/controllers.py:336-385
 *   normalize_angle        core/controllers.py:592-603   (Python floor-mod)
 *   speed target           core/controllers.py:58-73, :165-180
 *   PID (fixed dt)         core/controllers.py:32-47 + restated simple-pid  (PARITY UNPINNED, see oracle/__init__.py)
 *   LQG                    core/controllers.py:117-162   (gain K passed in: control.dlqr is restated with scipy)
 *   dynamic bicycle        core/data/car_state.py:178-223
 *   metrics                preformance_analysis/metrics_calculator.py:1-8 (pandas std => ddof = 1)
 *   loop order             nodes/controller.py:61-63, old/animation_loop.py:71,87-114
 * ORACLE-DEFINED (no reference code): kinematic bicycle, error logging for non-Stanley laws, cone hits, lap rule,
 * Philox observation noise -- exactly as documented in oracle/loop.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_PARAMS 16
#define N_METRICS 12
#define N_STATUS 10
#define N_TRAJ 12
enum { P_LF, P_KP, P_KI, P_KD, P_TARGET_SPEED, P_K_EXPO, P_K_STANLEY, P_K_SOFT, P_MU, P_C_F, P_C_R, P_MASS, P_I_Z };
enum { ST_PATH_END = 1, ST_LAPS_DONE = 2, ST_TIMEOUT = 4, ST_NONFINITE = 8 };

typedef struct {
  int32_t steer, accel, model; /* 0/1 as BGR_STEER_*, BGR_ACCEL_*, BGR_MODEL_* */
  int32_t max_steps, laps, laps_target, closed, n, n_cones;
  double dt, WB, l_f, l_r, max_steer, max_accel, max_decel, hit_radius;
  double lqr_k0, lqr_k1;
  double sigma_xy, sigma_yaw, sigma_v, sigma_cone;
  uint32_t seed, noise_on;
  int64_t episode_id_base;
  /* re-planning replay (oracle/loop.py step 2'; nodes/planner.py:26,62,72-77); replan_horizon = 0: off */
  int32_t replan_horizon, replan_stride, replan_after, reserved;
  double planner_lookahead;
} oracle_cfg_t;

static const double PI = 3.141592653589793;

/* (angle + pi) % (2 pi) - pi with Python's floor-mod */
static double normalize_angle(double a) {
  const double two_pi = 2 * PI;
  double t = a + PI;
  double m = fmod(t, two_pi);
  if (m != 0.0 && (m < 0.0)) m += two_pi;
  return m - PI;
}

static double clip(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }

/* ---- Philox4x32-10 + Box-Muller, identical to oracle/philox.py ---- */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    c[1] = (uint32_t)p1;
    c[3] = (uint32_t)p0;
    c[0] = n0;
    c[2] = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
static double u01(uint32_t r) { return ((double)(r >> 8) + 0.5) * (1.0 / 16777216.0); }
static void gaussians4(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, double z[4]) {
  uint32_t c[4] = {c0, c1, 0u, 0u};
  philox(c, k0, k1);
  double ua = u01(c[0]), ub = u01(c[1]), uc = u01(c[2]), ud = u01(c[3]);
  double ra = sqrt(-2.0 * log(ua)), rc = sqrt(-2.0 * log(uc));
  z[0] = ra * cos(2.0 * PI * ub);
  z[1] = ra * sin(2.0 * PI * ub);
  z[2] = rc * cos(2.0 * PI * ud);
  z[3] = rc * sin(2.0 * PI * ud);
}

/* core/controllers.py:353-376: global argmin (first minimum), heading error, cross-track error */
static int nearest_errors(const double* cx, const double* cy, int n, double x, double y, double yaw, double* theta_e,
                          double* e_ct) {
  int ind = 0;
  double best = hypot(cx[0] - x, cy[0] - y);
  for (int i = 1; i < n; ++i) {
    double d = hypot(cx[i] - x, cy[i] - y);
    if (d < best) {
      best = d;
      ind = i;
    }
  }
  double pdx, pdy;
  if (ind < n - 1) {
    pdx = cx[ind + 1] - cx[ind];
    pdy = cy[ind + 1] - cy[ind];
  } else {
    pdx = cx[ind] - cx[ind - 1];
    pdy = cy[ind] - cy[ind - 1];
  }
  double path_yaw = atan2(pdy, pdx);
  *theta_e = normalize_angle(path_yaw - yaw);
  double ddx = cx[ind] - x, ddy = cy[ind] - y;
  *e_ct = -ddx * sin(path_yaw) + ddy * cos(path_yaw);
  return ind;
}

static void episode(const oracle_cfg_t* c, const double* cx, const double* cy, const double* kappa,
                    const double* cones, const double* p, const double* init, int64_t episode_id, double* metrics,
                    int32_t* status, double* traj, unsigned char* hit) {
  const int n = c->n;
  const long n_tot = c->steer == 0 ? (long)n * c->laps : n;
  double x = init[0], y = init[1], yaw = init[2], v = init[3], r = init[4], beta = init[5];
  double pid_integral = 0.0, pid_last = 0.0;
  int pid_has_last = 0;
  double xh0 = 0.0, xh1 = 0.0, P00 = 1.0, P01 = 0.0, P10 = 0.0, P11 = 1.0, a_prev = 0.0;
  long ti = 0;
  long plan_org = -1, plan_ti = 0; /* Planner.current_path (its first waypoint) and Planner.target_ind */
  const int replan = c->steer == 0 && c->replan_horizon > 0;
  int prev_near = -1, lap = 0, flags = 0, steps = 0, cone_hits = 0;
  double s_lat = 0, s_head = 0, max_lat = 0;
  /* two-pass std like numpy: keep the rows */
  double* lat = (double*)malloc(sizeof(double) * 2 * (size_t)c->max_steps);
  double* hed = lat + c->max_steps;
  double* pc = NULL;
  if (c->n_cones > 0) {
    pc = (double*)malloc(sizeof(double) * 2 * (size_t)c->n_cones);
    memcpy(pc, cones, sizeof(double) * 2 * (size_t)c->n_cones);
    memset(hit, 0, (size_t)c->n_cones);
    if (c->noise_on && c->sigma_cone > 0.0) {
      for (int i = 0; i < c->n_cones; ++i) {
        double z[4];
        gaussians4((uint32_t)i, 1u, c->seed, (uint32_t)episode_id, z);
        pc[2 * i] += c->sigma_cone * z[0];
        pc[2 * i + 1] += c->sigma_cone * z[1];
      }
    }
  }
  const int obs_noise = c->noise_on && (c->sigma_xy > 0 || c->sigma_yaw > 0 || c->sigma_v > 0);

  for (int k = 0; k < c->max_steps; ++k) {
    /* 1. observation */
    double ox = x, oy = y, oyaw = yaw, ov = v;
    if (obs_noise) {
      double z[4];
      gaussians4((uint32_t)k, 0u, c->seed, (uint32_t)episode_id, z);
      ox = x + c->sigma_xy * z[0];
      oy = y + c->sigma_xy * z[1];
      oyaw = yaw + c->sigma_yaw * z[2];
      ov = v + c->sigma_v * z[3];
    }
    /* 2. steering */
    double delta;
    long tim;
    if (replan) {
      const long M = c->replan_horizon, q = c->replan_stride;
#define LOCAL_WP(kk) (c->closed ? (plan_org + q * (kk)) % n : ((plan_org + q * (kk)) < n - 1 ? (plan_org + q * (kk)) : n - 1))
      if (plan_org < 0 || plan_ti > c->replan_after) { /* nodes/planner.py:26 */
        long j0 = 0;
        double best = hypot(cx[0] - ox, cy[0] - oy);
        for (long i = 1; i < n; ++i) { /* the planned path starts at the waypoint nearest to the observed car */
          double d = hypot(cx[i] - ox, cy[i] - oy);
          if (d < best) { best = d; j0 = i; }
        }
        plan_org = j0;
        plan_ti = 0; /* :62 */
      }
      while (plan_ti < M - 1) { /* the planner's own scan, :72-77 */
        long m = LOCAL_WP(plan_ti);
        if (hypot(cx[m] - ox, cy[m] - oy) >= c->planner_lookahead) break;
        ++plan_ti;
      }
      ti = plan_ti; /* data["target_ind"], :101 */
      while (ti < M) { /* core/controllers.py:212-218 on the local path */
        long m = LOCAL_WP(ti);
        if (hypot(cx[m] - ox, cy[m] - oy) >= p[P_LF]) break;
        ++ti;
      }
      if (ti >= M) ti = M - 1; /* :219-220 */
      tim = LOCAL_WP(ti);
#undef LOCAL_WP
      double alpha = atan2(cy[tim] - oy, cx[tim] - ox) - oyaw;
      alpha = normalize_angle(alpha);
      delta = atan2(2.0 * c->WB * sin(alpha) / p[P_LF], 1.0);
      delta = clip(delta, -c->max_steer, c->max_steer);
    } else if (c->steer == 0) {
      while (ti < n_tot) { /* :212-218 on the unrolled (tiled) path */
        long m = ti % n;
        double d = hypot(cx[m] - ox, cy[m] - oy);
        if (d >= p[P_LF]) break;
        ++ti;
      }
      if (ti >= n_tot) ti = n_tot - 1; /* :219-220 */
      tim = ti % n;
      double alpha = atan2(cy[tim] - oy, cx[tim] - ox) - oyaw; /* :225 */
      alpha = normalize_angle(alpha);                          /* :226 */
      delta = atan2(2.0 * c->WB * sin(alpha) / p[P_LF], 1.0);  /* :228 */
      delta = clip(delta, -c->max_steer, c->max_steer);        /* :229 */
    } else {
      double th, ec;
      int ind = nearest_errors(cx, cy, n, ox, oy, oyaw, &th, &ec);
      delta = th + atan2(p[P_K_STANLEY] * ec, ov + p[P_K_SOFT]); /* :379 */
      delta = normalize_angle(delta);                            /* :380 */
      delta = clip(delta, -c->max_steer, c->max_steer);          /* :381 */
      ti = ind;
      tim = ind;
    }
    /* 3. curvature lookup, 4. acceleration */
    double kap = kappa[tim];
    double v_des = p[P_TARGET_SPEED] * exp(-p[P_K_EXPO] * fabs(kap));
    double a;
    if (c->accel == 0) {
      double error = v_des - ov;
      double d_input = ov - (pid_has_last ? pid_last : ov);
      double proportional = p[P_KP] * error;
      pid_integral += p[P_KI] * error * c->dt;
      double derivative = -p[P_KD] * d_input / c->dt;
      pid_last = ov;
      pid_has_last = 1;
      a = clip(proportional + pid_integral + derivative, c->max_decel, c->max_accel);
    } else {
      const double dt = c->dt;
      double z = ov - v_des;                                  /* :123 */
      double xp0 = (1.0 * xh0 + 0.0 * xh1) + dt * a_prev;    /* :144  A @ x_hat + B * a */
      double xp1 = (dt * xh0 + 1.0 * xh1) + 0.0 * a_prev;
      /* P_pred = A @ P @ A.T + Q_kf  (:145) -- numpy evaluates (A @ P) @ A.T */
      double ap00 = 1.0 * P00 + 0.0 * P10, ap01 = 1.0 * P01 + 0.0 * P11;
      double ap10 = dt * P00 + 1.0 * P10, ap11 = dt * P01 + 1.0 * P11;
      double pp00 = (ap00 * 1.0 + ap01 * 0.0) + 0.1, pp01 = (ap00 * dt + ap01 * 1.0) + 0.0;
      double pp10 = (ap10 * 1.0 + ap11 * 0.0) + 0.0, pp11 = (ap10 * dt + ap11 * 1.0) + 0.1;
      double yk = z - (1.0 * xp0 + 0.0 * xp1);                /* :148 */
      double S = (1.0 * pp00 + 0.0 * pp10) * 1.0 + (1.0 * pp01 + 0.0 * pp11) * 0.0 + 0.5; /* :151 */
      double invS = 1.0 / S;
      double k0 = (pp00 * 1.0 + pp01 * 0.0) * invS, k1 = (pp10 * 1.0 + pp11 * 0.0) * invS; /* :154 */
      xh0 = xp0 + k0 * yk;                                    /* :157 */
      xh1 = xp1 + k1 * yk;
      /* P = (I - K C) @ P_pred  (:160) */
      double m00 = 1.0 - k0 * 1.0, m01 = 0.0 - k0 * 0.0, m10 = 0.0 - k1 * 1.0, m11 = 1.0 - k1 * 0.0;
      P00 = m00 * pp00 + m01 * pp10;
      P01 = m00 * pp01 + m01 * pp11;
      P10 = m10 * pp00 + m11 * pp10;
      P11 = m10 * pp01 + m11 * pp11;
      a = -(c->lqr_k0 * xh0 + c->lqr_k1 * xh1);              /* :130 */
      a = clip(a, c->max_decel, c->max_accel);                /* :133 */
      a_prev = a;                                             /* :136 */
    }
    /* 5. logging on the TRUE state */
    double theta_e, e_ct;
    int near = nearest_errors(cx, cy, n, x, y, yaw, &theta_e, &e_ct);
    lat[k] = e_ct;
    hed[k] = theta_e;
    s_lat += e_ct;
    s_head += theta_e;
    if (fabs(e_ct) > max_lat) max_lat = fabs(e_ct);
    /* 6. cone hits */
    if (c->hit_radius > 0.0) {
      for (int i = 0; i < c->n_cones; ++i)
        if (!hit[i] && hypot(pc[2 * i] - x, pc[2 * i + 1] - y) <= c->hit_radius) {
          hit[i] = 1;
          ++cone_hits;
        }
    }
    /* 7. lap counter */
    if (prev_near >= 0 && c->closed) {
      if (prev_near - near > n / 2) ++lap;
      else if (near - prev_near > n / 2) --lap;
    }
    prev_near = near;
    if (traj != NULL) {
      double* row = traj + (size_t)k * N_TRAJ;
      row[0] = x; row[1] = y; row[2] = yaw; row[3] = v; row[4] = r; row[5] = beta; row[6] = delta; row[7] = a;
      row[8] = (double)ti; row[9] = (double)near; row[10] = e_ct; row[11] = theta_e;
    }
    /* 8. vehicle step */
    if (c->model == 1) {
      double vv = fmax(v, 1e-5);                                        /* :203-204 */
      double F_yf = -p[P_C_F] * (beta + (c->l_f * r) / vv - delta);     /* :207 */
      double F_yr = -p[P_C_R] * (beta - (c->l_r * r) / vv);             /* :208 */
      double lim = p[P_MU] * p[P_MASS] * 9.81;
      F_yf = clip(F_yf, -lim, lim);                                     /* :211 */
      F_yr = clip(F_yr, -lim, lim);                                     /* :212 */
      double xn = x + vv * cos(yaw + beta) * c->dt;                     /* :215 */
      double yn = y + vv * sin(yaw + beta) * c->dt;                     /* :216 */
      double yawn = yaw + r * c->dt;                                    /* :217 */
      double vn = vv + a * c->dt;                                       /* :218 */
      double rn = r + (c->l_f * F_yf - c->l_r * F_yr) / p[P_I_Z] * c->dt; /* :219 */
      double bn = beta + (F_yr / (p[P_MASS] * vv) - r) * c->dt;         /* :220 */
      x = xn; y = yn; yaw = yawn; v = vn; r = rn; beta = bn;
    } else {
      double xn = x + v * cos(yaw) * c->dt;
      double yn = y + v * sin(yaw) * c->dt;
      double yawn = yaw + v / c->WB * tan(delta) * c->dt;
      double vn = v + a * c->dt;
      x = xn; y = yn; yaw = yawn; v = vn;
    }
    /* 9. termination */
    steps = k + 1;
    if (!(isfinite(x) && isfinite(y) && isfinite(yaw) && isfinite(v) && isfinite(r) && isfinite(beta)))
      flags |= ST_NONFINITE;
    if (c->steer == 0 && !replan && ti >= n_tot - 1) flags |= ST_PATH_END;
    if (replan && !c->closed && tim >= n - 1) flags |= ST_PATH_END;
    if (c->steer == 1 && !c->closed && ti >= n_tot - 1) flags |= ST_PATH_END;
    if (c->laps_target > 0 && lap >= c->laps_target) flags |= ST_LAPS_DONE;
    if (flags) break;
  }
  if (!flags) flags = ST_TIMEOUT;

  /* metrics_calculator.py:1-8: mean, std (ddof = 1, two-pass like numpy), lap time */
  double nn = (double)steps, lat_mean = s_lat / nn, head_mean = s_head / nn, vl = 0, vh = 0;
  for (int k = 0; k < steps; ++k) {
    vl += (lat[k] - lat_mean) * (lat[k] - lat_mean);
    vh += (hed[k] - head_mean) * (hed[k] - head_mean);
  }
  metrics[0] = lat_mean;
  metrics[1] = steps > 1 ? sqrt(vl / (nn - 1.0)) : NAN;
  metrics[2] = head_mean;
  metrics[3] = steps > 1 ? sqrt(vh / (nn - 1.0)) : NAN;
  metrics[4] = (nn - 1.0) * c->dt;
  metrics[5] = max_lat;
  metrics[6] = x; metrics[7] = y; metrics[8] = yaw; metrics[9] = v; metrics[10] = r; metrics[11] = beta;
  status[0] = flags; status[1] = steps; status[2] = (int32_t)ti; status[3] = prev_near; status[4] = cone_hits;
  status[5] = lap; status[6] = 0; status[7] = -1; status[8] = 0; status[9] = 0;
  free(lat);
  free(pc);
}

/* params [E][16] double, init [E][6] double -> metrics [E][12] double, status [E][10] int32,
 * traj NULL or [E][max_steps][12] double.  Returns 0. */
int oracle_rollout(const oracle_cfg_t* cfg, const double* cx, const double* cy, const double* kappa,
                   const double* cones, int64_t n_episodes, const double* params, const double* init, double* metrics,
                   int32_t* status, double* traj, int n_threads) {
#pragma omp parallel num_threads(n_threads > 0 ? n_threads : 1)
  {
    unsigned char* hit = (unsigned char*)malloc(cfg->n_cones > 0 ? (size_t)cfg->n_cones : 1);
#pragma omp for schedule(dynamic, 4)
    for (int64_t e = 0; e < n_episodes; ++e) {
      episode(cfg, cx, cy, kappa, cones, params + e * N_PARAMS, init + e * 6, cfg->episode_id_base + e,
              metrics + e * N_METRICS, status + e * N_STATUS,
              traj != NULL ? traj + (size_t)e * cfg->max_steps * N_TRAJ : NULL, hit);
    }
    free(hit);
  }
  return 0;
}
