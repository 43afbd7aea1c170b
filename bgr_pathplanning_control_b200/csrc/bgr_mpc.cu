// MPC candidate-sequence rollout + cost evaluation (config 4).
//
// Replaces the cvxpy/OSQP solve inside MPCController.compute_control (core/controllers.py:486-568) by
// scoring K candidate control sequences per episode with the reference's own cost (:545, :548) under
// its own linearised dynamics (:522-527) and feasibility rule (:536), and taking the cheapest one.
//
// Two evaluators:
//   * PRODUCTION (precision fp32 in the ABI; evaluated in fp64): the FACTORED form.  Under the reference's
//     linearisation x and y do not depend on the controls (:523-524) and yaw / v are affine in PREFIX SUMS of the
//     control sequence (:525-526), so the whole (H + 1)-stage cost of candidate c for episode e is
//         J(e, c) = C0_e + G_c + C1_e SS_c + C2_e SS2_c + C3_e SA_c ,   feasible  <=>  v0_e >= vreq_c
//     with five per-candidate statistics (one pass over the bank) and four per-episode coefficients (one pass over
//     the states): 4 fused multiply-adds per (episode, candidate) pair instead of H stages.  Same cost, same argmin.
//   * AUDIT (precision fp64): the explicit stage-by-stage rollout below, in the reference's operation order.
//
// Mapping of the explicit rollout: ONE CANDIDATE PER THREAD.  A thread loads its candidate's H x (accel, steer) controls ONCE
// into registers and then walks over a strided list of episodes, so the candidate bank is read from
// HBM/L2 exactly once per CTA row and the inner loop touches only registers plus one broadcast
// shared-memory word per stage (the per-episode, candidate-independent x/y cost: :523-524 make x and y
// independent of the controls).  Candidate blocks of one episode meet through a 64-bit atomicMin on
// (cost bits << 32 | candidate index): order independent, hence deterministic, and the lowest index
// wins exact ties like np.argmin.  No tensor cores: there is no contraction here.
#include "bgr_launch.cuh"

namespace bgr {


template <typename Real>
struct EpisodeStage {
  Real xy[kMpcMaxH + 1];  // Q0 ex^2 + Q1 ey^2 at stage k (k = H: terminal, :548)
  Real yaw0, v0, c_yaw;   // c_yaw = v0 / WB (fp64 path) or v0 / WB * dt (fp32 path)
  unsigned long long best;
};

// First reference waypoint of an episode, made safe: ref_start is caller-supplied device data, and one bad row must not
// read out of bounds for the whole batch.  Closed tracks wrap it into [0, n) (floor-mod, negatives included); open
// tracks clamp it to [0, n - 1].  (bgr_mpc_eval_host rejects such rows up front; device callers get this.)
__device__ __forceinline__ int safe_ref_start(const MpcArgs& A, long long e) {
  const int n = A.trk.n;
  const int s = A.ref_start[e];
  if (A.trk.closed) {
    const int m = s % n;
    return m < 0 ? m + n : m;
  }
  return s < 0 ? 0 : (s > n - 1 ? n - 1 : s);
}

// Stage the candidate-independent part of one episode (threads 0..H of the CTA).
template <typename Real>
__device__ __forceinline__ void stage_episode(const MpcArgs& A, long long e, EpisodeStage<Real>* st) {
  const int t = threadIdx.x;
  if (t > A.H) return;
  const double* s = A.state + e * 4;
  const double x0 = s[0], y0 = s[1], yaw0 = s[2], v0 = s[3];
  const double cos_t = cos(yaw0), sin_t = sin(yaw0);               // :514-515
  // x_t, y_t by the same repeated additions as :523-524
  double sx = x0, sy = y0;
  for (int k = 0; k < t; ++k) {
    sx = sx + v0 * cos_t * A.dt;
    sy = sy + v0 * sin_t * A.dt;
  }
  // reference point of stage t (:539-542); the terminal cost reuses the last stage's (:548)
  const int n = A.trk.n;
  const int start = safe_ref_start(A, e);
  int len = A.ref_len != nullptr ? A.ref_len[e] : (A.trk.closed ? 0x3fffffff : n - start);
  if (len < 1) len = 1;
  int k_ref = t < A.H ? t : A.H - 1;
  if (k_ref >= len) k_ref = len - 1;
  long long idx = static_cast<long long>(start) + k_ref;
  if (A.trk.closed) idx %= n;
  else if (idx > n - 1) idx = n - 1;
  const double2 p = A.trk.xy_d[idx];
  const double ex = sx - p.x, ey = sy - p.y;
  st->xy[t] = static_cast<Real>(A.q0 * ex * ex + A.q1 * ey * ey);
  if (t == 0) {
    st->yaw0 = static_cast<Real>(yaw0);
    st->v0 = static_cast<Real>(v0);
    if constexpr (sizeof(Real) == 8) st->c_yaw = static_cast<Real>(v0 / A.wheelbase);
    else st->c_yaw = static_cast<Real>(v0 / A.wheelbase * A.dt);
    st->best = ~0ull;
  }
}

template <typename Real, int HT>  // HT > 0: horizon known at compile time, controls in registers
__global__ void __launch_bounds__(kMpcThreads) mpc_kernel(const __grid_constant__ MpcArgs A) {
  __shared__ EpisodeStage<Real> stage[2];
  const int H = HT > 0 ? HT : A.H;
  const int c = blockIdx.x * kMpcThreads + threadIdx.x;
  const bool cvalid = c < A.K;
  const int lane = threadIdx.x & 31;

  constexpr int HR = HT > 0 ? HT : 1;
  Real ua[HR], us[HR];
  const float2* my = reinterpret_cast<const float2*>(A.bank) + static_cast<size_t>(cvalid ? c : 0) * H;
  if constexpr (HT > 0) {
#pragma unroll
    for (int k = 0; k < HT; ++k) {
      const float2 u = __ldg(my + k);
      ua[k] = u.x;
      us[k] = u.y;
    }
  }

  const Real dt = static_cast<Real>(A.dt);
  const Real vt = static_cast<Real>(A.target_speed);
  const Real q2 = static_cast<Real>(A.q2), q3 = static_cast<Real>(A.q3);
  const Real r0 = static_cast<Real>(A.r0), r1 = static_cast<Real>(A.r1);

  long long e = blockIdx.y;
  int b = 0;
  if (e < A.n_episodes) stage_episode<Real>(A, e, &stage[0]);
  __syncthreads();
  for (; e < A.n_episodes; e += gridDim.y, b ^= 1) {
    const long long e_next = e + gridDim.y;
    if (e_next < A.n_episodes) stage_episode<Real>(A, e_next, &stage[b ^ 1]);

    const EpisodeStage<Real>& S = stage[b];
    Real yaw = S.yaw0, v = S.v0, cost = 0, vmin = Real(0);
    const Real cy = S.c_yaw;
    auto stage_step = [&](Real a, Real s, int k) {
      const Real ev = v - vt;
      // :545  quad_form(x - ref, Q) then quad_form(u, R)
      if constexpr (sizeof(Real) == 8) {
        cost += S.xy[k] + q2 * yaw * yaw + q3 * ev * ev;
        cost += r0 * a * a + r1 * s * s;
        yaw = yaw + cy * s * dt;                                  // :525
        v = v + a * dt;                                           // :526
      } else {
        cost += fmaf(q3 * ev, ev, fmaf(q2 * yaw, yaw, S.xy[k]));
        cost += fmaf(r0 * a, a, r1 * s * s);
        yaw = fmaf(cy, s, yaw);
        v = fmaf(a, dt, v);
      }
      vmin = fmin(vmin, v);                                       // :536  v_{k+1} >= 0
    };
    if constexpr (HT > 0) {
#pragma unroll
      for (int k = 0; k < HT; ++k) stage_step(ua[k], us[k], k);
    } else {
      for (int k = 0; k < H; ++k) {
        const float2 u = __ldg(my + k);
        stage_step(static_cast<Real>(u.x), static_cast<Real>(u.y), k);
      }
    }
    {
      const Real ev = v - vt;                                      // :548 terminal, last stage's reference
      if constexpr (sizeof(Real) == 8) cost += S.xy[H] + q2 * yaw * yaw + q3 * ev * ev;
      else cost += fmaf(q3 * ev, ev, fmaf(q2 * yaw, yaw, S.xy[H]));
    }
    float costf = static_cast<float>(cost);
    if (!(vmin >= Real(0)) || !cvalid || !(costf == costf)) costf = INFINITY;   // infeasible / NaN
    if (A.costs != nullptr && cvalid) A.costs[static_cast<size_t>(e) * A.K + c] = costf;

    // warp argmin with np.argmin's first-minimum rule: min over cost bits (costs are >= 0, so the
    // IEEE bit pattern is order preserving), then the lowest lane holding it
    const uint32_t bits = __float_as_uint(costf);
    const uint32_t wmin = __reduce_min_sync(0xffffffffu, bits);
    const uint32_t who = __ballot_sync(0xffffffffu, bits == wmin);
    if (lane == __ffs(who) - 1 && wmin < 0x7f800000u) {
      atomicMin(&stage[b].best, (static_cast<unsigned long long>(wmin) << 32) | static_cast<uint32_t>(c));
    }
    __syncthreads();
    if (threadIdx.x == 0 && stage[b].best != ~0ull) atomicMin(A.packed + e, stage[b].best);
    // stage[b] is rewritten only after the NEXT iteration's barrier, so thread 0's read above is safe
  }
}


// ================================================================================================
// factored evaluator (production)
// ================================================================================================
struct CandStats {   // per candidate: statistics of the prefix sums of its control sequence
  double G;          // sum_k (r0 a_k^2 + r1 s_k^2)  +  q3 * sum_{k=0..H} A_k^2          (:545 control cost + speed part)
  double SS, SS2;    // sum_{k=0..H} S_k,  sum S_k^2       S_k = dt * sum_{i<k} steer_i   (yaw_k = yaw0 + v0/WB * S_k, :525)
  double SA;         // sum_{k=0..H} A_k                   A_k = dt * sum_{i<k} accel_i   (v_k = v0 + A_k, :526)
  double vreq;       // -min_{k=1..H} A_k : feasible (v_k >= 0 for all k, :536) iff v0 >= vreq
};
constexpr double kMpcSane = 1e100;   // statistics / coefficients beyond this are treated as non-finite
struct EpiCoef {     // per episode
  double C0, C1, C2, C3, v0;
};

__device__ __forceinline__ void candidate_stats(const float* __restrict__ bank, int K, int H, double dt, double r0,
                                                double r1, double q3, CandStats* __restrict__ out, int c) {
  if (c >= K) return;
  const float2* u = reinterpret_cast<const float2*>(bank) + static_cast<size_t>(c) * H;
  double A = 0.0, S = 0.0, U = 0.0, SA = 0.0, SA2 = 0.0, SS = 0.0, SS2 = 0.0, Amin = 0.0;
  bool first = true;
  for (int k = 0; k < H; ++k) {          // stage k sees the prefix sums of the controls before it
    const float2 uk = __ldg(u + k);
    const double a = uk.x, s = uk.y;
    U += r0 * a * a + r1 * s * s;
    A = A + a * dt;                       // v_{k+1} - v0, accumulated like :526
    S = S + s * dt;                       // (yaw_{k+1} - yaw0) / (v0 / WB), :525
    SA += A; SA2 += A * A;                // stages 1..H (stage 0 contributes A_0 = 0)
    SS += S; SS2 += S * S;
    Amin = first ? A : fmin(Amin, A);
    first = false;
  }
  CandStats r;
  r.G = U + q3 * SA2;
  r.SS = SS; r.SS2 = SS2; r.SA = SA;
  r.vreq = -Amin;
  // A candidate with a non-finite (or absurdly large) statistic can never be selected: it is made infeasible here,
  // once, so that the pair loop of the evaluator needs no NaN test (and cannot overflow)
  if (!(fabs(r.G) < kMpcSane && fabs(SS) < kMpcSane && fabs(SS2) < kMpcSane && fabs(SA) < kMpcSane &&
        fabs(r.vreq) < kMpcSane)) {
    r.G = r.SS = r.SS2 = r.SA = 0.0;
    r.vreq = __longlong_as_double(0x7ff0000000000000LL);          // +inf: v0 >= vreq is false for every episode
  }
  out[c] = r;
}

__device__ __forceinline__ void episode_coef(const MpcArgs& A, EpiCoef* __restrict__ out, long long e) {
  if (e >= A.n_episodes) return;
  A.packed[e] = ~0ull;                                             // argmin accumulator of this episode
  const double* s = A.state + e * 4;
  const double x0 = s[0], y0 = s[1], yaw0 = s[2], v0 = s[3];
  const double cos_t = cos(yaw0), sin_t = sin(yaw0);               // :514-515
  const int n = A.trk.n;
  const int start = safe_ref_start(A, e);
  int len = A.ref_len != nullptr ? A.ref_len[e] : (A.trk.closed ? 0x3fffffff : n - start);
  if (len < 1) len = 1;
  double sx = x0, sy = y0, XY = 0.0;
  // reference point of stage t: waypoint start + min(t, H - 1, len - 1), modulo n (closed) or clipped (open); the
  // index is carried incrementally (one 64-bit modulo per episode instead of one per stage)
  const int last_ref = min(A.H - 1, len - 1);
  long long idx = static_cast<long long>(start);
  if (A.trk.closed) idx %= n;
  else if (idx > n - 1) idx = n - 1;
  int wp = static_cast<int>(idx);
  const double step_x = v0 * cos_t * A.dt, step_y = v0 * sin_t * A.dt;
  for (int t = 0; t <= A.H; ++t) {                                  // stages 0..H-1 and the terminal term (:548)
    const double2 p = A.trk.xy_d[wp];
    const double ex = sx - p.x, ey = sy - p.y;
    XY += A.q0 * ex * ex + A.q1 * ey * ey;
    sx = sx + step_x;                                               // :523
    sy = sy + step_y;                                               // :524
    if (t < last_ref) {
      ++wp;
      if (wp >= n) wp = A.trk.closed ? 0 : n - 1;
    }
  }
  const double cyw = v0 / A.wheelbase;                              // yaw_k = yaw0 + cyw * S_k
  const double ev0 = v0 - A.target_speed;
  const double n1 = static_cast<double>(A.H + 1);
  EpiCoef r;
  r.C0 = XY + A.q2 * n1 * yaw0 * yaw0 + A.q3 * n1 * ev0 * ev0;
  r.C1 = 2.0 * A.q2 * yaw0 * cyw;
  r.C2 = A.q2 * cyw * cyw;
  r.C3 = 2.0 * A.q3 * ev0;
  r.v0 = v0;
  if (!(fabs(r.C0) < kMpcSane && fabs(r.C1) < kMpcSane && fabs(r.C2) < kMpcSane && fabs(r.C3) < kMpcSane &&
        fabs(v0) < kMpcSane)) {      // non-finite state: no candidate is feasible (the solver-failure fallback)
    r.C0 = r.C1 = r.C2 = r.C3 = 0.0;
    r.v0 = __longlong_as_double(0xfff0000000000000LL);            // -inf
  }
  out[e] = r;
}

// one launch: the first `cand_blocks` CTAs compute candidate statistics, the rest episode coefficients
__global__ void __launch_bounds__(128) mpc_prepare_kernel(const __grid_constant__ MpcArgs A, int cand_blocks,
                                                          CandStats* __restrict__ cand, EpiCoef* __restrict__ epi) {
  if (static_cast<int>(blockIdx.x) < cand_blocks) {
    candidate_stats(A.bank, A.K, A.H, A.dt, A.r0, A.r1, A.q3, cand, blockIdx.x * 128 + threadIdx.x);
  } else {
    episode_coef(A, epi, static_cast<long long>(blockIdx.x - cand_blocks) * 128 + threadIdx.x);
  }
}

// One EPISODE per lane (its coefficients in registers), warp w of the CTA takes candidates w, w + 8, ... ; the
// candidate statistics are staged in shared memory in chunks and read as warp-wide broadcasts, so the inner loop is
// 4 fp64 FMAs + one compare per (episode, candidate) pair with no cross-thread traffic at all.  Each thread keeps
// the first minimum of its slice (ascending candidate index, strict <); the 8 slices of an episode meet in one
// 64-bit atomicMin on (cost bits << 32 | index): order independent, lowest index wins exact ties like np.argmin.
constexpr int kFactThreads = 256;
constexpr int kFactSlices = kFactThreads / 32;
constexpr int kFactChunk = 256;                  // candidates per CTA (10 KB of statistics)
#ifndef BGR_MPC_EPISODES_PER_LANE
#define BGR_MPC_EPISODES_PER_LANE 4   // measured on config 4: 1 -> 88 us (first version), 2 -> 50 us, 4 -> 46 us
#endif
constexpr int kFactPerLane = BGR_MPC_EPISODES_PER_LANE;   // episodes per lane
constexpr int kFactEpisodes = 32 * kFactPerLane;          // episodes per CTA

// One CTA scores 128 episodes x 256 candidates.  Lane l of every warp owns episodes l, l + 32, l + 64, l + 96 of the
// CTA's block: a candidate's five statistics are read from shared memory ONCE (broadcast) and used for four cost
// evaluations -- the first version (one episode per lane, all K candidates per CTA) was bound by exactly those loads
// (ncu: LSU pipe 66 % busy, 24 instructions per pair).  The eight warps take every eighth candidate; their winners
// are merged through shared memory, then one 64-bit atomicMin per episode merges the candidate chunks (grid.y).
// grid = (ceil(E / 128), ceil(K / 256)): 2 048 small CTAs for config 4 -- 7 waves of 148 x 2, no tail to speak of.
__global__ void __launch_bounds__(kFactThreads) mpc_factored_kernel(const CandStats* __restrict__ cand,
                                                                    const EpiCoef* __restrict__ epi, int K,
                                                                    long long n_episodes,
                                                                    unsigned long long* __restrict__ packed,
                                                                    float* __restrict__ costs) {
  __shared__ __align__(16) CandStats s_c[kFactChunk];
  __shared__ unsigned long long s_best[kFactSlices][kFactEpisodes];
  const int lane = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const long long e_first = static_cast<long long>(blockIdx.x) * kFactEpisodes + lane;   // + 32 r for r = 0..
  EpiCoef ec[kFactPerLane];
  float best[kFactPerLane];
  int best_i[kFactPerLane];
#pragma unroll
  for (int r = 0; r < kFactPerLane; ++r) {
    const long long e = e_first + 32 * r;
    ec[r] = epi[e < n_episodes ? e : 0];
    best[r] = INFINITY;
    best_i[r] = -1;
  }
  const int c0 = blockIdx.y * kFactChunk;
  const int nc = min(kFactChunk, K - c0);
  {  // coalesced copy of the chunk (5 doubles per candidate)
    const double* src = reinterpret_cast<const double*>(cand + c0);
    double* dst = reinterpret_cast<double*>(s_c);
    for (int i = threadIdx.x; i < nc * 5; i += kFactThreads) dst[i] = src[i];
  }
  __syncthreads();
#pragma unroll 4
  for (int c = slice; c < nc; c += kFactSlices) {
    const CandStats cs = s_c[c];
#pragma unroll
    for (int r = 0; r < kFactPerLane; ++r) {
      // three FMAs and one add (the four episodes of a lane are independent chains: depth does not matter, count does)
      const double J = fma(ec[r].C3, cs.SA, fma(ec[r].C2, cs.SS2, fma(ec[r].C1, cs.SS, cs.G))) + ec[r].C0;
      // (J is finite: non-finite inputs were made infeasible by the prepare kernel)   infeasible: :536
      const float costf = ec[r].v0 >= cs.vreq ? fmaxf(static_cast<float>(J), 0.0f) : INFINITY;   // 0: round-off below an exact zero
      if (costs != nullptr) {
        const long long e = e_first + 32 * r;
        if (e < n_episodes) costs[static_cast<size_t>(e) * K + c0 + c] = costf;
      }
      if (costf < best[r]) {    // ascending c: strict < keeps the first minimum
        best[r] = costf;
        best_i[r] = c0 + c;
      }
    }
  }
  // (cost bits << 32 | index): costs are >= 0, so the packed words order like (cost, index); ~0 = nothing feasible
#pragma unroll
  for (int r = 0; r < kFactPerLane; ++r)
    s_best[slice][lane + 32 * r] =
        best[r] < INFINITY ? (static_cast<unsigned long long>(__float_as_uint(best[r])) << 32) | static_cast<uint32_t>(best_i[r])
                           : ~0ull;
  __syncthreads();
  if (threadIdx.x < kFactEpisodes) {
    unsigned long long m = s_best[0][threadIdx.x];
#pragma unroll
    for (int sl = 1; sl < kFactSlices; ++sl) m = min(m, s_best[sl][threadIdx.x]);
    const long long e = static_cast<long long>(blockIdx.x) * kFactEpisodes + threadIdx.x;
    if (e < n_episodes && m != ~0ull) atomicMin(packed + e, m);
  }
}

// candidate statistics only (a bank handle computes them once per cost configuration and keeps them)
cudaError_t launch_mpc_cand_stats(const MpcArgs& A, void* cand_out, cudaStream_t stream) {
  const int cand_blocks = (A.K + 127) / 128;
  mpc_prepare_kernel<<<cand_blocks, 128, 0, stream>>>(A, cand_blocks, static_cast<CandStats*>(cand_out), nullptr);
  return cudaGetLastError();
}

// `cand_ready`: scratch_cand already holds the statistics of this bank for this cost configuration (bank handle)
cudaError_t launch_mpc_factored(const MpcArgs& A, void* scratch_cand, void* scratch_epi, cudaStream_t stream,
                                int sm_count, bool cand_ready) {
  (void)sm_count;
  CandStats* cand = static_cast<CandStats*>(scratch_cand);
  EpiCoef* epi = static_cast<EpiCoef*>(scratch_epi);
  const int cand_blocks = cand_ready ? 0 : (A.K + 127) / 128;
  const unsigned epi_blocks = static_cast<unsigned>((A.n_episodes + 127) / 128);
  mpc_prepare_kernel<<<cand_blocks + epi_blocks, 128, 0, stream>>>(A, cand_blocks, cand, epi);
  const long long chunks = (A.K + kFactChunk - 1) / kFactChunk;
  if (chunks > 65535) return cudaErrorInvalidValue;   // K < 16.7 M candidates
  const dim3 grid(static_cast<unsigned>((A.n_episodes + kFactEpisodes - 1) / kFactEpisodes), static_cast<unsigned>(chunks));
  mpc_factored_kernel<<<grid, kFactThreads, 0, stream>>>(cand, epi, A.K, A.n_episodes, A.packed, A.costs);
  return cudaGetLastError();
}

size_t mpc_cand_bytes(int K) { return static_cast<size_t>(K) * sizeof(CandStats); }
size_t mpc_epi_bytes(long long E) { return static_cast<size_t>(E) * sizeof(EpiCoef); }

// packed (cost bits, index) -> (acceleration, steering) with the return convention of :565-568
__global__ void mpc_finalize_kernel(const unsigned long long* __restrict__ packed, const float* __restrict__ bank,
                                    int H, long long n_episodes, float max_accel, float max_decel, float max_steer,
                                    float* __restrict__ best_u, float* __restrict__ best_cost,
                                    int32_t* __restrict__ best_idx) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= n_episodes) return;
  const unsigned long long p = packed[e];
  if (p == ~0ull) {  // no feasible candidate: the solver-failure fallback (:558-562)
    if (best_u != nullptr) {
      best_u[2 * e] = 0.0f;
      best_u[2 * e + 1] = -0.0f;
    }
    if (best_cost != nullptr) best_cost[e] = INFINITY;
    if (best_idx != nullptr) best_idx[e] = -1;
    return;
  }
  const int idx = static_cast<int>(p & 0xffffffffull);
  const float2 u0 = reinterpret_cast<const float2*>(bank)[static_cast<size_t>(idx) * H];
  if (best_u != nullptr) {
    best_u[2 * e] = fminf(fmaxf(u0.x, max_decel), max_accel);             // :565
    best_u[2 * e + 1] = -fminf(fmaxf(u0.y, -max_steer), max_steer);       // :566, :568
  }
  if (best_cost != nullptr) best_cost[e] = __uint_as_float(static_cast<uint32_t>(p >> 32));
  if (best_idx != nullptr) best_idx[e] = idx;
}

template <typename Real, int HT>
static cudaError_t launch_mpc_t(const MpcArgs& A, cudaStream_t stream, int sm_count) {
  auto kern = mpc_kernel<Real, HT>;
  int per_sm = 1;
  cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMpcThreads, 0);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  const int gx = (A.K + kMpcThreads - 1) / kMpcThreads;
  long long gy = (static_cast<long long>(sm_count) * per_sm) / gx;
  if (gy < 1) gy = 1;
  if (gy > A.n_episodes) gy = A.n_episodes;
  if (gy > 65535) gy = 65535;
  kern<<<dim3(gx, static_cast<unsigned>(gy)), kMpcThreads, 0, stream>>>(A);
  return cudaGetLastError();
}

cudaError_t launch_mpc(const MpcArgs& A, int precision, cudaStream_t stream, int sm_count) {
  if (precision == BGR_PRECISION_FP64) {
    switch (A.H) {
      case 10: return launch_mpc_t<double, 10>(A, stream, sm_count);
      default: return launch_mpc_t<double, 0>(A, stream, sm_count);
    }
  }
  switch (A.H) {
    case 10: return launch_mpc_t<float, 10>(A, stream, sm_count);
    case 20: return launch_mpc_t<float, 20>(A, stream, sm_count);
    case 30: return launch_mpc_t<float, 30>(A, stream, sm_count);
    default: return launch_mpc_t<float, 0>(A, stream, sm_count);
  }
}

cudaError_t launch_mpc_finalize(const unsigned long long* packed, const float* bank, int H, long long n_episodes,
                                float max_accel, float max_decel, float max_steer, float* best_u,
                                float* best_cost, int32_t* best_idx, cudaStream_t stream) {
  const int threads = 256;
  const long long blocks = (n_episodes + threads - 1) / threads;
  mpc_finalize_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(
      packed, bank, H, n_episodes, max_accel, max_decel, max_steer, best_u, best_cost, best_idx);
  return cudaGetLastError();
}

}  // namespace bgr
