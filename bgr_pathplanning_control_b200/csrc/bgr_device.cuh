// Device-side helpers shared by the rollout / step / MPC kernels (sm_100a).
//
// Everything is templated on `Real`:
//   float  -- production path: the tracking laws, tyre model and nearest/look-ahead searches run in
//             fp32; the four integrators x, y, yaw, v are carried in fp64 and positions are handed to
//             the fp32 code as (rounded, residual) pairs so that waypoint offsets lose nothing;
//   double -- audit path: every operation in fp64, in the reference's operation order
//             (compiled with --fmad=false so nothing is contracted).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/bgr_rollout.h"

namespace bgr {

constexpr double kPi = 3.14159265358979323846;
constexpr double kTwoPi = 2.0 * 3.14159265358979323846;
constexpr int kRolloutThreads = 256;
constexpr double kRivalReach = 2.5;                                   // [m] rival lists are complete up to here (host)
constexpr double kTightGuard = 0.5;                                   // [m] a guard below this counts as tight (host: search_direct)
constexpr double kRivalAccept = kRivalReach * (1.0 - 1e-4) - 1e-4;   // [m] ... and trusted up to here (device)

struct alignas(32) D4 {
  double x, y, z, w;
};

template <typename Real>
struct RT;
template <>
struct RT<float> {
  using R2 = float2;
  using R4 = float4;
};
template <>
struct RT<double> {
  using R2 = double2;
  using R4 = D4;
};

// Device view of a track handle.  Per-waypoint tables: the base lap followed by a cyclic pad (entry j >= n
// replicates entry j - n) so that unrolled forward scans need no index wrap; n_pad = n + 12 rounded up to 4 (the
// fp32 xy table also carries 4 cyclic entries IN FRONT of waypoint 0: entry t holds waypoint t - 4):
//   xy    : (x, y)                                 nodes/planner.py:88
//   attr  : (kappa, path_yaw, sin path_yaw, cos path_yaw)   nodes/planner.py:89, core/controllers.py:361-368
//   guard2: squared exactness-guard radius of the local nearest search
struct TrackView {
  const float2* xy_f;
  const float4* attr_f;
  const double2* xy_d;
  const D4* attr_d;
  const float* guard2;
  const int2* rival_range;    // [n] (first, count) into rival_list; count < 0: none
  const int2* rival_list;     // (waypoint, bits of its squared reach-in distance), ascending per list
  const int2* twin;           // [n] (twin waypoint t or -1, bits of the squared twin guard): inside it the exact argmin is
                              // the best of the descent's answer and t - 1, t, t + 1 (bgr_api.cu: compute_guards)
  const float4* scan_blk;     // [ceil(n / 32)] (centre x, centre y, radius, -) of 32 consecutive waypoints: the exact
                              // scan of the fp32 kernels skips every block that provably holds nothing within reach
  const int32_t* cone_range;  // [n] (first << 10) | count  into cone_list
  const float* cone_near;     // [n] distance from waypoint j to its closest listed cone, rounded down
  const int32_t* cone_list;
  const double2* cones;       // [n_cones]
  const float2* cones_f;      // [n_cones] rounded copy: the fp32 kernels' cheap distance pre-test in front of the exact one
  int32_t n;
  int32_t n_pad;
  int32_t closed;
  int32_t n_cones;
  float ds_max;
};

// float views of the configuration, converted once on the host (the fp32 kernels read them straight from the
// constant bank instead of converting doubles inside the time loop)
struct RolloutF32 {
  float dt, inv_dt, WB, inv_WB, l_f, l_r, max_steer, tan_max_steer, max_accel, max_decel;
  float tie_eps, hit_radius, K0, K1, sigma_xy, sigma_yaw, sigma_v, sigma_cone;
  float st_max_delta, st_alpha;   // shadowed Stanley: max_steer_rate * dt, alpha
  float st_lookahead;             // shadowed Stanley: optional look-ahead advance (core/controllers.py:276-281); 0 = off
  float plan_Lf;                  // re-planning replay: the planner's own look-ahead distance
  float dt_lo;                    // dt - float(dt): the double-float integrators advance by dt_hi + dt_lo
};

struct RolloutArgs {
  TrackView trk;
  RolloutF32 f;
  const float* params;
  const double* init;
  float* metrics;
  int32_t* status;
  double* traj;
  const double* kf_gain;  // [kf_len][2] shared Kalman gain sequence (LQG) or nullptr
  const float2* kf_gain_f; // the same sequence rounded to fp32 (production kernels)
  int32_t kf_len;
  // two-phase execution (episodes that outlive the first, short launch are re-run from scratch, densely packed):
  int32_t* queue;         // nullptr (static schedule), or a zeroed device counter: DYNAMIC episode scheduling -- every
                          // lane of a one-wave grid pulls its next row from it (full fp32 build only)
  const int32_t* idx;     // nullptr, or [count] episode row handled by thread t (rows of params / init / outputs)
  const int32_t* count;   // nullptr (= n_episodes threads are live), or device scalar: number of live threads
  long long n_episodes;
  long long episode_id_base;
  int32_t max_steps;
  int32_t traj_stride;    // rows per episode in the trajectory dump (the configured step count, also in a capped phase)
  int32_t laps;
  int32_t laps_target;
  int32_t lap_stop;       // laps_target, or INT_MAX when the lap rule is off (one compare in the step kernels)
  int32_t n_total;  // PP: n * laps; Stanley: n
  int32_t ti_chunk_max;  // fp32 kernels: last target index from which a 3-waypoint chunk may start (n_total - 4;
                         // INT_MIN on tracks of fewer than 8 waypoints, whose scans run one waypoint at a time)
  double dt, wheelbase, l_f, l_r, max_steer, max_accel, max_decel;
  double hit_radius, tie_eps;
  double lqr_k0, lqr_k1;
  double inv_dt, inv_wheelbase, tan_max_steer;  // host-computed loop invariants of the fp32 kernels
  double sigma_xy, sigma_yaw, sigma_v, sigma_cone;
  double st_rate, st_alpha;   // shadowed Stanley variant (core/controllers.py:234-324)
  double st_lookahead;        // ... its optional look-ahead advance (:276-281); 0 = off
  int32_t stanley_shadowed;
  // receding-horizon re-planning replay (nodes/planner.py:26,62,72-77); replan_M = 0: off
  int32_t replan_M, replan_q, replan_after;
  double plan_Lf;
  uint32_t seed;
  int32_t cones_on;
  int32_t attr_in_smem;   // fp32 kernels: attr / guard tables staged in shared memory (short tracks) or read via L1
  int32_t noise_on;
  int32_t search_direct;  // fp32 full build: skip the five-point window of the nearest search and run the complete search
                          // straight away (host: pose noise on, or most of the track inside a tight exactness guard)
  // single-step plumbing (EXTRAS kernels only; all nullptr for a rollout)
  int32_t phases;
  const double* ctrl_in;   // [E][BGR_N_CTRL]
  double* ctrl_out;
  const int32_t* ti_in;    // [E]
  int32_t* ti_out;
  double* state_out;       // [E][BGR_N_STATE] final state in fp64
  const double* params64;  // [E][BGR_N_PARAMS] fp64 parameter rows (replace `params` when set)
  const double* cmd;       // [E][BGR_N_CMD] (steer, accel, kappa) used for switched-off phases
  double* step_out;        // [E][BGR_N_OUT] outputs of the LAST executed step
};

// ------------------------------------------------------------------------------------------------
// TMA (bulk async copy) staging of the track tables into shared memory: one elected thread arms an
// mbarrier with the byte count and issues cp.async.bulk per table (SASS: UBLKCP); everybody waits on
// the barrier phase.  Sizes and addresses are multiples of 16 B by construction (n_pad % 4 == 0).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}

// ------------------------------------------------------------------------------------------------
// normalize_angle (core/controllers.py:592-603): (angle + pi) % (2 pi) - pi with floor-mod => [-pi, pi)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double normalize_angle(double a) {
  double t = a + kPi;
  double m = fmod(t, kTwoPi);
  if (m != 0.0) {
    if (m < 0.0) m += kTwoPi;
  } else {
    m = 0.0;
  }
  return m - kPi;
}

// fp32 version for arguments known to lie in (-3 pi, 3 pi) -- every call site subtracts two angles
// that are each within [-pi, pi] (+ one atan2).  Two-term 2 pi keeps the result accurate to an ulp.
__device__ __forceinline__ float normalize_angle(float a) {
  const float kPiF = 3.14159274101257324f;       // float(pi)
  const float kTwoPiHi = 6.28318548202514648f;   // float(2 pi)
  const float kTwoPiLo = -1.74845553146951715e-7f;  // 2 pi - float(2 pi)
  // k = +1 above pi, -1 below -pi, else 0; the two fused steps round exactly like (a -+ hi) -+ lo and leave a
  // untouched for k = 0 (five instructions, no branch)
  const float k = (a >= kPiF ? 1.0f : 0.0f) - (a < -kPiF ? 1.0f : 0.0f);
  a = fmaf(k, -kTwoPiHi, a);
  return fmaf(k, -kTwoPiLo, a);
}

// Heading handed to the laws.  double: the unwrapped yaw itself (reference semantics).  float: yaw
// reduced to [-pi, pi] in fp64 first (yaw is never wrapped by the vehicle model, car_state.py:217, and
// grows to tens of radians; fp32 would lose 1e-6 rad there).
template <typename Real>
__device__ __forceinline__ Real heading_for_laws(double yaw);
template <>
__device__ __forceinline__ double heading_for_laws<double>(double yaw) {
  return yaw;
}
template <>
__device__ __forceinline__ float heading_for_laws<float>(double yaw) {
  double k = rint(yaw * (1.0 / kTwoPi));
  return static_cast<float>(fma(-k, kTwoPi, yaw));
}

template <typename Real>
__device__ __forceinline__ Real clampr(Real v, Real lo, Real hi) {
  return fmin(fmax(v, lo), hi);  // np.clip = minimum(maximum(v, lo), hi)
}

__device__ __forceinline__ void sincos_r(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincos_r(double a, double* s, double* c) {
  *s = sin(a);
  *c = cos(a);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11); identical integer stream to oracle/philox.py
// ------------------------------------------------------------------------------------------------
struct U4 {
  uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                            uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

template <typename Real>
__device__ __forceinline__ Real u01_24(uint32_t r) {
  return (static_cast<Real>(r >> 8) + Real(0.5)) * Real(1.0 / 16777216.0);
}

// sin / cos of 2 pi u for u in (0, 1).  float: sincospif has an exact argument reduction (no 2 pi rounding, no
// Payne-Hanek path); double: the oracle's own expression cos(2 pi u), sin(2 pi u)
__device__ __forceinline__ void sincos_2pi(float u, float* s, float* c) { sincospif(2.0f * u, s, c); }
__device__ __forceinline__ void sincos_2pi(double u, double* s, double* c) { sincos_r(kTwoPi * u, s, c); }

// two Box-Muller pairs -> four standard normals (oracle/philox.py:gaussians4)
template <typename Real>
__device__ __forceinline__ void gaussians4(U4 r, Real z[4]) {
  Real ua = u01_24<Real>(r.x), ub = u01_24<Real>(r.y), uc = u01_24<Real>(r.z), ud = u01_24<Real>(r.w);
  Real ra = sqrt(Real(-2.0) * log(ua));
  Real rc = sqrt(Real(-2.0) * log(uc));
  Real s, c;
  sincos_2pi(ub, &s, &c);
  z[0] = ra * c;
  z[1] = ra * s;
  sincos_2pi(ud, &s, &c);
  z[2] = rc * c;
  z[3] = rc * s;
}

// fp32 production path: the same transform on the SFU (MUFU.LG2 / MUFU.SIN / MUFU.COS).  The integer stream and the
// 24-bit uniforms are identical to the oracle's; the normals differ from its fp64 Box-Muller by ~1e-6 (absolute, in
// units of sigma) -- an observation perturbation of 5e-8 m at sigma_xy = 0.05 m, far inside the parity budget, for
// ~90 instructions per noisy step less than logf + sincospif.
template <>
__device__ __forceinline__ void gaussians4<float>(U4 r, float z[4]) {
  const float ua = u01_24<float>(r.x), ub = u01_24<float>(r.y), uc = u01_24<float>(r.z), ud = u01_24<float>(r.w);
  const float ra = sqrtf(fmaxf(-2.0f * __logf(ua), 0.0f));
  const float rc = sqrtf(fmaxf(-2.0f * __logf(uc), 0.0f));
  // angle 2 pi u - pi in [-pi, pi): cos(2 pi u) = -cos(angle), sin(2 pi u) = -sin(angle)
  const float ab = fmaf(6.28318530717958647692f, ub, -3.14159265358979323846f);
  const float ad = fmaf(6.28318530717958647692f, ud, -3.14159265358979323846f);
  z[0] = -ra * __cosf(ab);
  z[1] = -ra * __sinf(ab);
  z[2] = -rc * __cosf(ad);
  z[3] = -rc * __sinf(ad);
}

constexpr uint32_t kStreamObs = 0;
constexpr uint32_t kStreamCone = 1;

}  // namespace bgr
