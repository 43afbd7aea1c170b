// The fp64 AUDIT closed-loop rollout kernel (reference operation order; the production fp32 kernel is
// bgr_rollout_f32.cuh): ONE EPISODE PER THREAD, the whole time loop in-kernel, vehicle /
// controller / metric state in registers, the base-lap track tables staged once per CTA into shared
// memory by TMA bulk copies.  No tensor cores: nothing here is a dense contraction.
//
// Per step (order = oracle/loop.py, which composes nodes/controller.py:61-63 and
// old/animation_loop.py:71,87-114 around State.update, core/data/car_state.py:118-125):
//   observe -> steering law -> curvature lookup -> acceleration law -> error logging ->
//   cone hits -> lap counter -> vehicle step -> termination test.
#pragma once

#include <type_traits>

#include "bgr_launch.cuh"

namespace bgr {

template <typename Real, int STEER, int ACCEL, int MODEL, bool EXTRAS>
__global__ void __launch_bounds__(kRolloutThreads) rollout_kernel(const __grid_constant__ RolloutArgs A) {
  using R2 = typename RT<Real>::R2;
  using R4 = typename RT<Real>::R4;
  constexpr bool kF64 = std::is_same<Real, double>::value;

  extern __shared__ __align__(128) unsigned char smem[];
  const int n = A.trk.n;
  const int n_pad = A.trk.n_pad;
  R2* s_xy = reinterpret_cast<R2*>(smem);
  R4* s_attr = reinterpret_cast<R4*>(smem + static_cast<size_t>(n_pad) * sizeof(R2));
  float* s_guard2 = reinterpret_cast<float*>(smem + static_cast<size_t>(n_pad) * (sizeof(R2) + sizeof(R4)));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(n_pad) * (sizeof(R2) + sizeof(R4) + 4));

  // ---- stage the track: TMA bulk copies, one mbarrier phase ----
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t b_xy = static_cast<uint32_t>(n_pad * sizeof(R2));
    const uint32_t b_at = static_cast<uint32_t>(n_pad * sizeof(R4));
    const uint32_t b_g = static_cast<uint32_t>(n_pad * sizeof(float));
    mbar_expect_tx(bar, b_xy + b_at + b_g);
    if constexpr (kF64) {
      tma_bulk_g2s(s_xy, A.trk.xy_d, b_xy, bar);
      tma_bulk_g2s(s_attr, A.trk.attr_d, b_at, bar);
    } else {
      tma_bulk_g2s(s_xy, A.trk.xy_f, b_xy, bar);
      tma_bulk_g2s(s_attr, A.trk.attr_f, b_at, bar);
    }
    tma_bulk_g2s(s_guard2, A.trk.guard2, b_g, bar);
  }

  const long long t_glob = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = t_glob < (A.count != nullptr ? static_cast<long long>(__ldg(A.count)) : A.n_episodes);
  // row of this thread's episode: identity, or through the survivor list of a two-phase run
  const long long e = active ? (A.idx != nullptr ? static_cast<long long>(__ldg(A.idx + t_glob)) : t_glob) : 0;
  const long long er = e;

  // ---- per-episode parameters: four coalesced 128-bit loads of the 64 B row ----
  Real Lf, kp, ki, kd, tspeed, k_expo, k_st, k_soft, mu, c_f, c_r, mass, I_z;
  bool params_f64 = false;
  if constexpr (EXTRAS) params_f64 = A.params64 != nullptr;
  if (params_f64) {
    // single-step drop-in API: the reference's scalars are fp64, keep every bit of them
    const double* p = A.params64 + er * BGR_N_PARAMS;
    Lf = static_cast<Real>(p[BGR_P_LF]); kp = static_cast<Real>(p[BGR_P_KP]); ki = static_cast<Real>(p[BGR_P_KI]);
    kd = static_cast<Real>(p[BGR_P_KD]); tspeed = static_cast<Real>(p[BGR_P_TARGET_SPEED]);
    k_expo = static_cast<Real>(p[BGR_P_K_EXPO]); k_st = static_cast<Real>(p[BGR_P_K_STANLEY]);
    k_soft = static_cast<Real>(p[BGR_P_K_SOFT]); mu = static_cast<Real>(p[BGR_P_MU]);
    c_f = static_cast<Real>(p[BGR_P_C_F]); c_r = static_cast<Real>(p[BGR_P_C_R]);
    mass = static_cast<Real>(p[BGR_P_MASS]); I_z = static_cast<Real>(p[BGR_P_I_Z]);
  } else {
    const float4* prow = reinterpret_cast<const float4*>(A.params) + er * (BGR_N_PARAMS / 4);
    const float4 q0 = __ldg(prow + 0), q1 = __ldg(prow + 1), q2 = __ldg(prow + 2), q3 = __ldg(prow + 3);
    Lf = q0.x; kp = q0.y; ki = q0.z; kd = q0.w;
    tspeed = q1.x; k_expo = q1.y; k_st = q1.z; k_soft = q1.w;
    mu = q2.x; c_f = q2.y; c_r = q2.z; mass = q2.w;
    I_z = q3.x;
  }

  const double* irow = A.init + er * BGR_N_STATE;
  double X = __ldg(irow + 0), Y = __ldg(irow + 1), YAW = __ldg(irow + 2), V = __ldg(irow + 3);
  Real r = static_cast<Real>(__ldg(irow + 4)), beta = static_cast<Real>(__ldg(irow + 5));

  const Real dt = static_cast<Real>(A.dt);
  const Real inv_dt = static_cast<Real>(1.0 / A.dt);
  const Real WB = static_cast<Real>(A.wheelbase);
  const Real max_steer = static_cast<Real>(A.max_steer);
  const Real max_accel = static_cast<Real>(A.max_accel), max_decel = static_cast<Real>(A.max_decel);
  const bool closed = A.trk.closed != 0;
  const int n_total = A.n_total;

  // look-ahead threshold in the metric the searches compare in (fp32: squared distance; fp64: hypot)
  const Real Lf_metric = kF64 ? Lf : Lf * Lf;
  const Real tie_eps = static_cast<Real>(A.tie_eps);
  const Real tie_pp = kF64 ? tie_eps : tie_eps * (Real(2) * Lf + tie_eps);

  // controller memory (zero for a fresh episode; the single-step API carries it in/out)
  Real pid_integ = 0, pid_last = 0;
  bool pid_has_last = false;
  Real xh0 = 0, xh1 = 0, a_prev = 0;
  Real prev_steer = 0;   // shadowed Stanley: previous_steering (core/controllers.py:250)
  int k_gain0 = 0;
  int phases = BGR_PHASE_STEER | BGR_PHASE_ACCEL | BGR_PHASE_MODEL | BGR_PHASE_ERRORS;
  const Real K0 = static_cast<Real>(A.lqr_k0), K1 = static_cast<Real>(A.lqr_k1);

  // metric accumulators (fp64: four adds per step)
  double s_lat = 0, s_lat2 = 0, s_head = 0, s_head2 = 0;
  Real max_lat = 0;
  int steps = 0, flags = 0, ti = 0, tim = 0, near = -1, lap = 0;
  // re-planning replay (nodes/planner.py): first waypoint of the current local path, the planner's own index on it
  int plan_org = -1, plan_ti = 0;
  const bool replan = EXTRAS && STEER == BGR_STEER_PURE_PURSUIT && A.replan_M > 0;
  int cone_hits = 0, ties = 0, first_tie = -1, fallbacks = 0;
  uint32_t hitmask[EXTRAS ? BGR_MAX_CONES / 32 : 1];
  if constexpr (EXTRAS) {
#pragma unroll
    for (int i = 0; i < BGR_MAX_CONES / 32; ++i) hitmask[i] = 0u;
  }

  if constexpr (EXTRAS) {
    phases = A.phases;
    if (A.ctrl_in != nullptr) {
      const double* c = A.ctrl_in + er * BGR_N_CTRL;
      pid_integ = static_cast<Real>(c[BGR_C_PID_INTEGRAL]);
      pid_last = static_cast<Real>(c[BGR_C_PID_LAST_INPUT]);
      pid_has_last = c[BGR_C_PID_HAS_LAST] != 0.0;
      xh0 = static_cast<Real>(c[BGR_C_LQG_XHAT0]);
      xh1 = static_cast<Real>(c[BGR_C_LQG_XHAT1]);
      a_prev = static_cast<Real>(c[BGR_C_LQG_A_PREV]);
      k_gain0 = static_cast<int>(c[BGR_C_STEP]);
      prev_steer = static_cast<Real>(c[BGR_C_PREV_STEER]);
    }
    if (A.ti_in != nullptr) {
      ti = A.ti_in[er];
      ti = ti < 0 ? 0 : ti;
      tim = ti % n;
    }
  }

  mbar_wait(bar, 0);  // track tables have landed

  if (active) {
    for (int k = 0; k < A.max_steps; ++k) {
      // ---- position as (rounded, residual) so that waypoint offsets are exact to fp32 ----
      Real xr, xl, yr, yl;
      if constexpr (kF64) {
        xr = X; xl = 0; yr = Y; yl = 0;
      } else {
        xr = static_cast<float>(X); xl = static_cast<float>(X - static_cast<double>(xr));
        yr = static_cast<float>(Y); yl = static_cast<float>(Y - static_cast<double>(yr));
      }
      const Real v = static_cast<Real>(V);
      const Real yaw = heading_for_laws<Real>(YAW);

      // ---- 1. observation (config 5: Philox noise on the observed pose) ----
      Real oxl = xl, oyl = yl, oyaw = yaw, ov = v;
      bool noisy = false;
      if constexpr (EXTRAS) {
        if (A.noise_on) {
          Real z[4];
          // counter = the episode's own step count (a stepped closed loop draws the same stream as one rollout)
          gaussians4<Real>(philox4x32_10(static_cast<uint32_t>(k_gain0 + k), kStreamObs, 0u, 0u, A.seed,
                                         static_cast<uint32_t>(A.episode_id_base + e)), z);
          // observed x = x + sigma z  =>  offsets (p - x_obs) = (p - xr) - (xl + sigma z)
          oxl = xl + static_cast<Real>(A.sigma_xy) * z[0];
          oyl = yl + static_cast<Real>(A.sigma_xy) * z[1];
          oyaw = yaw + static_cast<Real>(A.sigma_yaw) * z[2];
          ov = v + static_cast<Real>(A.sigma_v) * z[3];
          noisy = true;
        }
      }

      auto dist = [&](int j, Real lx, Real ly) -> Real {
        const R2 p = s_xy[j];
        const Real dx = (p.x - xr) - lx;
        const Real dy = (p.y - yr) - ly;
        if constexpr (kF64) return hypot(dx, dy);      // np.hypot, core/controllers.py:215,355
        else return dx * dx + dy * dy;
      };

      bool tie_now = false;

      // ---- nearest waypoint = np.argmin(np.hypot(cx - x, cy - y)) (core/controllers.py:353-356) ----
      // Local descent from the previous nearest index, accepted only inside the per-waypoint exactness
      // guard; otherwise the exact global scan.
      auto nearest = [&](Real lx, Real ly, int start) -> int {
        int j = start;
        Real dj = 0, d_nb = Real(INFINITY);
        bool global = start < 0;
        if (!global) {
          dj = dist(j, lx, ly);
          bool moved = false;
          while (true) {
            int jn = j + 1;
            if (jn == n) {
              if (!closed) break;
              jn = 0;
            }
            const Real dn = dist(jn, lx, ly);
            if (!(dn < dj || (dn == dj && jn < j))) {
              d_nb = dn;
              break;
            }
            d_nb = dj;
            j = jn;
            dj = dn;
            moved = true;
          }
          if (!moved) {
            while (true) {
              int jp = j - 1;
              if (jp < 0) {
                if (!closed) break;
                jp = n - 1;
              }
              const Real dp = dist(jp, lx, ly);
              if (!(dp < dj || (dp == dj && jp < j))) {
                d_nb = fmin(d_nb, dp);
                break;
              }
              d_nb = dj;
              j = jp;
              dj = dp;
            }
          }
          const Real dj2 = kF64 ? dj * dj : dj;
          if (!(dj2 < static_cast<Real>(s_guard2[j]))) {
            // outside the plain guard: best of {j} U rivals(j) within the lists' reach, else the exact scan
            const int2 rr = __ldg(A.trk.rival_range + j);
            if (rr.y >= 0 && dj2 < static_cast<Real>(kRivalAccept * kRivalAccept)) {
              const Real d_own2 = dj2;
              for (int q = 0; q < rr.y; ++q) {
                const int2 ent = __ldg(A.trk.rival_list + rr.x + q);
                if (static_cast<Real>(__int_as_float(ent.y)) > d_own2) break;   // sorted: nobody further can win
                const int i = ent.x;
                const Real di = dist(i, lx, ly);
                if (di < dj || (di == dj && i < j)) {
                  d_nb = fmin(d_nb, dj);
                  dj = di;
                  j = i;
                } else {
                  d_nb = fmin(d_nb, di);
                }
              }
            } else {
              global = true;
            }
          }
        }
        if (global) {
          ++fallbacks;
          j = 0;
          dj = dist(0, lx, ly);
          d_nb = Real(INFINITY);
          for (int i = 1; i < n; ++i) {
            const Real d = dist(i, lx, ly);
            if (d < dj) {
              d_nb = dj;
              dj = d;
              j = i;
            } else {
              d_nb = fmin(d_nb, d);
            }
          }
        }
        if constexpr (EXTRAS) {
          // audit: runner-up within tie_eps [m] of the winner
          if constexpr (kF64) {
            tie_now |= (d_nb - dj) <= tie_eps;
          } else {
            tie_now |= d_nb <= dj + tie_eps * (Real(2) * sqrtf(dj) + tie_eps);
          }
        }
        return j;
      };

      auto errors_at = [&](int j, Real lx, Real ly, Real yaw_used, Real& e_ct, Real& theta_e) {
        const R4 at = s_attr[j];
        const R2 p = s_xy[j];
        const Real ddx = (p.x - xr) - lx;                           // :374
        const Real ddy = (p.y - yr) - ly;                           // :375
        e_ct = -ddx * at.z + ddy * at.w;                            // :376
        theta_e = normalize_angle(at.y - yaw_used);                 // :371
      };

      // ---- 2. steering law ----
      Real delta;
      int j_obs = -1;
      bool steer_on = true, accel_on = true;
      if constexpr (EXTRAS) {
        steer_on = (phases & BGR_PHASE_STEER) != 0;
        accel_on = (phases & BGR_PHASE_ACCEL) != 0;
      }
      if (!steer_on) {
        delta = static_cast<Real>(A.cmd[er * BGR_N_CMD + BGR_CMD_STEER]);
      } else if constexpr (STEER == BGR_STEER_PURE_PURSUIT) {
        if (replan) {
          // receding-horizon replay: Planner.update (nodes/planner.py:26-77), then compute_steering on the LOCAL path
          // (nodes/controller.py:56-61); local point k is waypoint plan_org + k * replan_q of the centreline
          if (plan_org < 0 || plan_ti > A.replan_after) {            // :26
            plan_org = nearest(oxl, oyl, near);                      // the new path starts at the car
            plan_ti = 0;                                             // :62
          }
          const int M = A.replan_M, q = A.replan_q;
          auto waypoint_of = [&](int kk) -> int {
            const long long g = static_cast<long long>(plan_org) + static_cast<long long>(q) * kk;
            return closed ? static_cast<int>(g % n) : static_cast<int>(g < n - 1 ? g : n - 1);
          };
          const Real pl = static_cast<Real>(A.plan_Lf);
          const Real pl_metric = kF64 ? pl : pl * pl;
          const Real tie_pl = kF64 ? tie_eps : tie_eps * (Real(2) * pl + tie_eps);
          while (plan_ti < M - 1) {                                  // the planner's own scan, :72-77
            const Real d = dist(waypoint_of(plan_ti), oxl, oyl);
            tie_now |= fabs(d - pl_metric) <= tie_pl;
            if (d >= pl_metric) break;
            ++plan_ti;
          }
          int lt = plan_ti;                                          // data["target_ind"], :101
          while (lt < M) {                                           // core/controllers.py:212-218
            const Real d = dist(waypoint_of(lt), oxl, oyl);
            tie_now |= fabs(d - Lf_metric) <= tie_pp;
            if (d >= Lf_metric) break;
            ++lt;
          }
          if (lt >= M) lt = M - 1;                                   // :219-220
          ti = lt;
          tim = waypoint_of(lt);
        }
        // forward scan for the first index with distance >= Lf (core/controllers.py:212-220)
        while (!replan && ti < n_total) {
          const Real d = dist(tim, oxl, oyl);
          if constexpr (EXTRAS) tie_now |= fabs(d - Lf_metric) <= tie_pp;
          if (d >= Lf_metric) break;
          ++ti;
          tim = (tim + 1 == n) ? 0 : tim + 1;
        }
        if (!replan && ti >= n_total) {
          ti = n_total - 1;
          tim = n - 1;
        }
        const R2 tp = s_xy[tim];
        const Real tdx = (tp.x - xr) - oxl, tdy = (tp.y - yr) - oyl;
        Real alpha = atan2(tdy, tdx) - oyaw;                         // :225
        alpha = normalize_angle(alpha);                              // :226
        if constexpr (kF64) {
          delta = atan2(2.0 * WB * sin(alpha) / Lf, 1.0);            // :228
        } else {
          delta = atanf(Real(2) * WB * sinf(alpha) / Lf);
        }
        delta = clampr(delta, -max_steer, max_steer);                // :229
      } else {
        // live Stanley law (core/controllers.py:336-385)
        j_obs = nearest(oxl, oyl, near);
        bool shadowed = false;
        if constexpr (EXTRAS) shadowed = A.stanley_shadowed != 0;
        int j_law = j_obs;
        if constexpr (EXTRAS) {
          if (shadowed && A.st_lookahead > 0.0) {
            // optional look-ahead advance of the shadowed variant (core/controllers.py:276-281): from the nearest
            // waypoint, forward while it is closer than lookahead_distance and below the last index
            const Real la = static_cast<Real>(A.st_lookahead);
            const Real la_metric = kF64 ? la : la * la;
            const Real tie_la = kF64 ? tie_eps : tie_eps * (Real(2) * la + tie_eps);
            while (j_law < n - 1) {                                  // :277
              const Real d = dist(j_law, oxl, oyl);                  // :278
              tie_now |= fabs(d - la_metric) <= tie_la;
              if (d >= la_metric) break;                             // :279
              ++j_law;                                               // :281
            }
          }
        }
        Real e_ct, theta_e;
        errors_at(j_law, oxl, oyl, oyaw, e_ct, theta_e);
        if (shadowed) {
          // the SHADOWED class body (core/controllers.py:305-322): adaptive gain, rate limit, low-pass filter
          const Real adaptive_k = k_st / (ov + k_soft);              // :305
          delta = theta_e + atan2(adaptive_k * e_ct, Real(1));       // :308
          delta = normalize_angle(delta);                            // :309
          const Real max_delta = static_cast<Real>(A.st_rate) * dt;  // :312
          delta = clampr(delta, prev_steer - max_delta, prev_steer + max_delta);   // :313
          const Real al = static_cast<Real>(A.st_alpha);
          delta = al * delta + (Real(1) - al) * prev_steer;          // :316
          prev_steer = delta;                                        // :319
          delta = clampr(delta, -max_steer, max_steer);              // :322
        } else {
          delta = theta_e + atan2(k_st * e_ct, ov + k_soft);         // :379
          delta = normalize_angle(delta);                            // :380
          delta = clampr(delta, -max_steer, max_steer);              // :381
        }
        ti = j_law;                                                  // the law returns the index it used (:324 / :385)
        tim = j_law;
      }

      // ---- 3. curvature lookup (nodes/controller.py:62) ----
      Real kappa = s_attr[tim].x;
      if constexpr (EXTRAS) {
        if (phases & BGR_PHASE_KAPPA_GIVEN) kappa = static_cast<Real>(A.cmd[er * BGR_N_CMD + BGR_CMD_KAPPA]);
      }

      // ---- 4. acceleration law ----
      const Real v_des = tspeed * exp(-k_expo * fabs(kappa));        // core/controllers.py:72 / :179
      Real acc;
      if (!accel_on) {
        acc = static_cast<Real>(A.cmd[er * BGR_N_CMD + BGR_CMD_ACCEL]);
      } else if constexpr (ACCEL == BGR_ACCEL_PID) {
        // AccelerationPIDController (:32-47) around restated simple_pid with fixed dt
        const Real err = v_des - ov;
        const Real d_in = pid_has_last ? ov - pid_last : Real(0);
        pid_integ += ki * err * dt;
        Real deriv;
        if constexpr (kF64) deriv = -kd * d_in / dt;
        else deriv = -kd * d_in * inv_dt;
        pid_last = ov;
        pid_has_last = true;
        acc = clampr(kp * err + pid_integ + deriv, max_decel, max_accel);   // :46
      } else {
        // LQGAccelerationController (:117-138); Kalman gain from the shared data-independent sequence
        const Real zk = ov - v_des;                                  // :123
        const Real xp0 = xh0 + dt * a_prev;                          // :144
        const Real xp1 = dt * xh0 + xh1;
        const Real yk = zk - xp0;                                    // :148
        const int kg = min(k_gain0 + k, A.kf_len - 1);
        const double2 g = __ldg(reinterpret_cast<const double2*>(A.kf_gain) + kg);
        xh0 = xp0 + static_cast<Real>(g.x) * yk;                     // :157
        xh1 = xp1 + static_cast<Real>(g.y) * yk;
        acc = clampr(-(K0 * xh0 + K1 * xh1), max_decel, max_accel);  // :130-133
        a_prev = acc;                                                // :136
      }

      // ---- 5. error logging on the TRUE state (Stanley's formulas, :353-376) ----
      int j_true;
      Real e_ct = 0, theta_e = 0;
      bool errors_on = true;
      if constexpr (EXTRAS) errors_on = (phases & BGR_PHASE_ERRORS) != 0;
      if (STEER == BGR_STEER_STANLEY && !noisy && steer_on) {
        j_true = j_obs;
      } else if (errors_on) {
        j_true = nearest(xl, yl, near);
      } else {
        j_true = near < 0 ? 0 : near;
      }
      if (errors_on) {
        errors_at(j_true, xl, yl, yaw, e_ct, theta_e);
        s_lat += static_cast<double>(e_ct);
        s_lat2 += static_cast<double>(e_ct) * static_cast<double>(e_ct);
        s_head += static_cast<double>(theta_e);
        s_head2 += static_cast<double>(theta_e) * static_cast<double>(theta_e);
        max_lat = fmax(max_lat, fabs(e_ct));
      }

      // ---- 6. cone hits (oracle-defined): distinct cones within hit_radius of (x, y) ----
      if constexpr (EXTRAS) {
        if (A.cones_on && errors_on) {
          const Real hr = static_cast<Real>(A.hit_radius);
          const Real hr_metric = kF64 ? hr : hr * hr;
          const Real dnear = dist(j_true, xl, yl);
          const Real guard_metric = kF64 ? Real(BGR_CONE_GUARD) : Real(BGR_CONE_GUARD * BGR_CONE_GUARD);
          int first = 0, cnt = A.trk.n_cones;
          const bool listed = dnear <= guard_metric;
          if (listed) {
            const int32_t cr = __ldg(A.trk.cone_range + j_true);
            first = cr >> 10;
            cnt = cr & 1023;
          }
          for (int i = 0; i < cnt; ++i) {
            const int c = listed ? __ldg(A.trk.cone_list + first + i) : i;
            const double2 cp = __ldg(A.trk.cones + c);
            Real cdx = static_cast<Real>(cp.x - X), cdy = static_cast<Real>(cp.y - Y);
            // per-(cone, episode) perturbation: only drawn when the unperturbed cone is within reach of the hit
            // circle (8 sigma per component: P < 1e-15 of mattering)
            const Real far_ = hr + Real(11.4) * static_cast<Real>(A.sigma_cone);
            if (A.sigma_cone > 0.0 && cdx * cdx + cdy * cdy <= far_ * far_) {
              Real z[4];
              gaussians4<Real>(philox4x32_10(static_cast<uint32_t>(c), kStreamCone, 0u, 0u, A.seed,
                                             static_cast<uint32_t>(A.episode_id_base + e)), z);
              cdx += static_cast<Real>(A.sigma_cone) * z[0];
              cdy += static_cast<Real>(A.sigma_cone) * z[1];
            }
            Real dc;
            if constexpr (kF64) dc = hypot(cdx, cdy);
            else dc = cdx * cdx + cdy * cdy;
            const Real band = kF64 ? tie_eps : tie_eps * (Real(2) * hr + tie_eps);
            tie_now |= fabs(dc - hr_metric) <= band;
            if (dc <= hr_metric) {
              const uint32_t bit = 1u << (c & 31);
              if (!(hitmask[c >> 5] & bit)) {
                hitmask[c >> 5] |= bit;
                ++cone_hits;
              }
            }
          }
        }
      }

      // ---- 7. lap counter (oracle-defined, closed tracks) ----
      if (closed && near >= 0 && errors_on) {
        const int dj = near - j_true;
        if (dj > n / 2) ++lap;
        else if (-dj > n / 2) --lap;
      }
      near = j_true;

      if constexpr (EXTRAS) {
        if (tie_now) {
          ++ties;
          if (first_tie < 0) first_tie = k;
        }
        if (A.traj != nullptr) {
          double* row = A.traj + (static_cast<size_t>(e) * A.traj_stride + k) * BGR_N_TRAJ;
          row[BGR_T_X] = X; row[BGR_T_Y] = Y; row[BGR_T_YAW] = YAW; row[BGR_T_V] = V;
          row[BGR_T_R] = r; row[BGR_T_BETA] = beta; row[BGR_T_STEER] = delta; row[BGR_T_ACCEL] = acc;
          row[BGR_T_TARGET_IND] = ti; row[BGR_T_NEAREST_IND] = j_true;
          row[BGR_T_E_CT] = e_ct; row[BGR_T_THETA_E] = theta_e;
        }
      }

      if constexpr (EXTRAS) {
        if (A.step_out != nullptr) {
          double* o = A.step_out + er * BGR_N_OUT;
          o[BGR_O_STEER] = delta; o[BGR_O_ACCEL] = acc; o[BGR_O_V_DESIRED] = v_des; o[BGR_O_KAPPA] = kappa;
          o[BGR_O_E_CT] = e_ct; o[BGR_O_THETA_E] = theta_e; o[BGR_O_NEAREST_IND] = j_true; o[7] = 0.0;
        }
      }

      // ---- 8. vehicle step: State.update(a, delta) (core/data/car_state.py:118-125) ----
      bool model_on = true;
      if constexpr (EXTRAS) model_on = (phases & BGR_PHASE_MODEL) != 0;
      if (!model_on) {
        // single-step API with the model phase off: state is left untouched
      } else if constexpr (MODEL == BGR_MODEL_DYNAMIC) {
        // DynamicBicycleModel.state_equations (core/data/car_state.py:178-223)
        const double Vc = fmax(V, 1e-5);                               // :203-204
        const Real vv = static_cast<Real>(Vc);
        Real F_yf, F_yr;
        if constexpr (kF64) {
          F_yf = -c_f * (beta + (A.l_f * r) / vv - delta);             // :207
          F_yr = -c_r * (beta - (A.l_r * r) / vv);                     // :208
        } else {
          const Real inv_v = Real(1) / vv;
          F_yf = -c_f * (beta + static_cast<Real>(A.l_f) * r * inv_v - delta);
          F_yr = -c_r * (beta - static_cast<Real>(A.l_r) * r * inv_v);
        }
        const Real fmax_ = mu * mass * Real(9.81);
        F_yf = clampr(F_yf, -fmax_, fmax_);                            // :211
        F_yr = clampr(F_yr, -fmax_, fmax_);                            // :212
        Real sn, cs;
        if constexpr (kF64) {
          cs = cos(YAW + beta);
          sn = sin(YAW + beta);
          X = X + Vc * cs * A.dt;                                      // :215
          Y = Y + Vc * sn * A.dt;                                      // :216
          YAW = YAW + r * A.dt;                                        // :217
          V = Vc + acc * A.dt;                                         // :218
          const Real r_next = r + (A.l_f * F_yf - A.l_r * F_yr) / I_z * A.dt;   // :219
          beta = beta + (F_yr / (mass * Vc) - r) * A.dt;               // :220
          r = r_next;
        } else {
          sincosf(yaw + beta, &sn, &cs);
          X = fma(static_cast<double>(vv * cs), A.dt, X);
          Y = fma(static_cast<double>(vv * sn), A.dt, Y);
          YAW = fma(static_cast<double>(r), A.dt, YAW);
          V = fma(static_cast<double>(acc), A.dt, Vc);
          const Real r_next = r + (static_cast<Real>(A.l_f) * F_yf - static_cast<Real>(A.l_r) * F_yr) / I_z * dt;
          beta = beta + (F_yr / (mass * vv) - r) * dt;
          r = r_next;
        }
      } else {
        // oracle-defined kinematic bicycle: x += v cos(yaw) dt; y += v sin(yaw) dt;
        // yaw += v / WB * tan(delta) dt; v += a dt   (all right-hand sides from the old state)
        if constexpr (kF64) {
          const double cs = cos(YAW), sn = sin(YAW);
          const double Xn = X + V * cs * A.dt;
          const double Yn = Y + V * sn * A.dt;
          const double YAWn = YAW + V / WB * tan(delta) * A.dt;
          V = V + acc * A.dt;
          X = Xn; Y = Yn; YAW = YAWn;
        } else {
          float sn, cs;
          sincosf(yaw, &sn, &cs);
          X = fma(static_cast<double>(v * cs), A.dt, X);
          Y = fma(static_cast<double>(v * sn), A.dt, Y);
          YAW = fma(static_cast<double>(v / WB * tanf(delta)), A.dt, YAW);
          V = fma(static_cast<double>(acc), A.dt, V);
        }
      }

      // ---- 9. termination (old/animation_loop.py:71 + oracle-defined lap rule) ----
      steps = k + 1;
      if (!(isfinite(X) && isfinite(Y) && isfinite(YAW) && isfinite(V) && isfinite(r) && isfinite(beta)))
        flags |= BGR_ST_NONFINITE;
      if (STEER == BGR_STEER_PURE_PURSUIT && !replan && ti >= n_total - 1) flags |= BGR_ST_PATH_END;
      if (replan && !closed && tim >= n - 1) flags |= BGR_ST_PATH_END;
      if (STEER == BGR_STEER_STANLEY && !closed && ti >= n_total - 1) flags |= BGR_ST_PATH_END;
      if (A.laps_target > 0 && lap >= A.laps_target) flags |= BGR_ST_LAPS_DONE;
      if (flags) break;
    }
    if (!flags) flags = BGR_ST_TIMEOUT;
  }

  // ---- per-episode results: metrics_calculator.py:1-8 over the logged rows ----
  double lat_mean = 0, lat_std = 0, lap_time = 0;
  if (active) {
    const double nn = static_cast<double>(steps);
    lat_mean = s_lat / nn;
    const double head_mean = s_head / nn;
    const double nan_ = __longlong_as_double(0x7ff8000000000000LL);
    lat_std = steps > 1 ? sqrt(fmax(0.0, (s_lat2 - s_lat * s_lat / nn) / (nn - 1.0))) : nan_;   // ddof = 1
    const double head_std = steps > 1 ? sqrt(fmax(0.0, (s_head2 - s_head * s_head / nn) / (nn - 1.0))) : nan_;
    lap_time = (nn - 1.0) * A.dt;                                      // timestamp[-1] - timestamp[0]
    float4* m = reinterpret_cast<float4*>(A.metrics + e * BGR_N_METRICS);
    m[0] = make_float4(static_cast<float>(lat_mean), static_cast<float>(lat_std),
                       static_cast<float>(head_mean), static_cast<float>(head_std));
    m[1] = make_float4(static_cast<float>(lap_time), static_cast<float>(max_lat), static_cast<float>(X),
                       static_cast<float>(Y));
    m[2] = make_float4(static_cast<float>(YAW), static_cast<float>(V), static_cast<float>(r),
                       static_cast<float>(beta));
    int32_t* st = A.status + e * BGR_N_STATUS;
    st[BGR_S_FLAGS] = flags; st[BGR_S_STEPS] = steps; st[BGR_S_TARGET_IND] = ti; st[BGR_S_NEAREST_IND] = near;
    st[BGR_S_CONE_HITS] = cone_hits; st[BGR_S_LAPS] = lap; st[BGR_S_NEAR_TIES] = ties;
    st[BGR_S_FIRST_TIE] = first_tie; st[BGR_S_FALLBACKS] = fallbacks; st[9] = 0;
    if constexpr (EXTRAS) {
      if (A.state_out != nullptr) {
        double* so = A.state_out + e * BGR_N_STATE;
        so[BGR_X] = X; so[BGR_Y] = Y; so[BGR_YAW] = YAW; so[BGR_V] = V; so[BGR_R] = r; so[BGR_BETA] = beta;
      }
      if (A.ctrl_out != nullptr) {
        double* c = A.ctrl_out + e * BGR_N_CTRL;
        c[BGR_C_PID_INTEGRAL] = pid_integ; c[BGR_C_PID_LAST_INPUT] = pid_last;
        c[BGR_C_PID_HAS_LAST] = pid_has_last ? 1.0 : 0.0;
        c[BGR_C_LQG_XHAT0] = xh0; c[BGR_C_LQG_XHAT1] = xh1; c[BGR_C_LQG_A_PREV] = a_prev;
        c[BGR_C_STEP] = static_cast<double>(k_gain0 + steps); c[BGR_C_PREV_STEER] = prev_steer;
      }
      if (A.ti_out != nullptr) A.ti_out[e] = ti;
    }
  }

}

// Host-side launcher, one explicit specialisation per (Real, STEER, ACCEL, MODEL, EXTRAS).
#define BGR_DEFINE_LAUNCH(Real, STEER, ACCEL, MODEL, EXTRAS)                                              \
  template <>                                                                                            \
  cudaError_t launch_rollout<Real, STEER, ACCEL, MODEL, EXTRAS>(const RolloutArgs& args, int grid,        \
                                                                 size_t smem, int smem_optin,            \
                                                                 cudaStream_t stream, int* query_resident) { \
    auto kern = rollout_kernel<Real, STEER, ACCEL, MODEL, EXTRAS>;                                        \
    /* per-function attribute: always the device's opt-in maximum, so concurrent callers cannot lower it */ \
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin); \
    if (err != cudaSuccess) return err;                                                                  \
    if (query_resident != nullptr) {                                                                     \
      int ctas = 0;                                                                                      \
      err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, kRolloutThreads, smem);           \
      *query_resident = ctas * kRolloutThreads;                                                          \
      return err;                                                                                        \
    }                                                                                                    \
    kern<<<grid, kRolloutThreads, smem, stream>>>(args);                                                 \
    return cudaGetLastError();                                                                           \
  }

}  // namespace bgr
