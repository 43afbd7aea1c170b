// The C ABI (include/bgr_rollout.h): argument checking, the track handle (derived tables, exactness
// guards, cone candidate lists), kernel dispatch, host-buffer convenience entry points, LQG design and
// the measurement helpers.  Nothing here computes a step on the CPU: without a CUDA device every compute
// entry point returns BGR_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "bgr_launch.cuh"
#include "bgr_track.cuh"

namespace bgr {

static thread_local char g_error[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t err, const char* what) {
  if (err == cudaSuccess) return BGR_OK;
  set_error("%s: %s (%s)", what, cudaGetErrorString(err), cudaGetErrorName(err));
  return BGR_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define BGR_CUDA(call)                                   \
  do {                                                   \
    int rc_ = ::bgr::check_cuda((call), #call);          \
    if (rc_ != BGR_OK) return rc_;                       \
  } while (0)

// CUDA-event pair around the last hot kernel launched by this thread (bgr_last_kernel_ms)
struct LastKernel {
  cudaEvent_t start = nullptr, stop = nullptr;
  int device = -1;
  bool valid = false;
};
static thread_local LastKernel g_last;

static int ensure_events(int device) {
  if (g_last.start != nullptr && g_last.device == device) return BGR_OK;
  if (g_last.start != nullptr) {
    cudaEventDestroy(g_last.start);
    cudaEventDestroy(g_last.stop);
    g_last.start = g_last.stop = nullptr;
  }
  BGR_CUDA(cudaEventCreate(&g_last.start));
  BGR_CUDA(cudaEventCreate(&g_last.stop));
  g_last.device = device;
  g_last.valid = false;
  return BGR_OK;
}

struct DeviceInfo {
  int sm_count = 0;
  int max_smem_optin = 0;
};
static int device_info(int device, DeviceInfo* out) {
  static std::mutex mu;
  static std::vector<DeviceInfo> cache;
  std::lock_guard<std::mutex> lock(mu);
  if (static_cast<int>(cache.size()) <= device) cache.resize(device + 1);
  if (cache[device].sm_count == 0) {
    BGR_CUDA(cudaDeviceGetAttribute(&cache[device].sm_count, cudaDevAttrMultiProcessorCount, device));
    BGR_CUDA(cudaDeviceGetAttribute(&cache[device].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  }
  *out = cache[device];
  return BGR_OK;
}

// ------------------------------------------------------------------------------------------------
// LQG design (host math): LQR gain of control.dlqr (core/controllers.py:97) by Riccati iteration and the
// data-independent Kalman gain sequence (:145, :151, :154, :160).
// ------------------------------------------------------------------------------------------------
static void lqr_gain_host(double dt, double K[2]) {
  // A = [[1,0],[dt,1]], B = [dt,0]^T, Q = diag(10,100), R = 0.001   (:83-94)
  const long double a10 = dt, b0 = dt, q0 = 10.0L, q1 = 100.0L, R = 0.001L;
  long double p00 = q0, p01 = 0, p11 = q1;  // symmetric P
  long double k0 = 0, k1 = 0;
  for (int it = 0; it < 200000; ++it) {
    // PA = P A ; A = [[1,0],[a10,1]]
    const long double pa00 = p00 + p01 * a10, pa01 = p01, pa10 = p01 + p11 * a10, pa11 = p11;
    // BtPA = B^T P A = b0 * row0(PA)
    const long double s = R + b0 * p00 * b0;
    const long double g0 = b0 * pa00, g1 = b0 * pa01;
    k0 = g0 / s;
    k1 = g1 / s;
    // AtPA = A^T (P A)
    const long double m00 = pa00 + a10 * pa10, m01 = pa01 + a10 * pa11, m11 = pa11;
    const long double n00 = m00 - g0 * k0 + q0, n01 = m01 - g0 * k1, n11 = m11 - g1 * k1 + q1;
    const long double diff = fabsl(n00 - p00) + fabsl(n01 - p01) + fabsl(n11 - p11);
    p00 = n00; p01 = n01; p11 = n11;
    if (diff <= 1e-19L * (fabsl(p00) + fabsl(p11))) break;
  }
  K[0] = static_cast<double>(k0);
  K[1] = static_cast<double>(k1);
}

static void kalman_gains_host(double dt, int n_steps, double* out /*[n][2] or nullptr*/, double* P_out = nullptr) {
  // A = [[1,0],[dt,1]], C = [1,0], Q_kf = 0.1 I, R_kf = 0.5, P0 = I   (:83-87, :101-109)
  double p00 = 1.0, p01 = 0.0, p10 = 0.0, p11 = 1.0;
  for (int k = 0; k < n_steps; ++k) {
    // AP = A P
    const double ap00 = p00, ap01 = p01, ap10 = dt * p00 + p10, ap11 = dt * p01 + p11;
    // P_pred = AP A^T + Q_kf ; A^T = [[1,dt],[0,1]]
    const double pp00 = ap00 + 0.1, pp01 = ap00 * dt + ap01, pp10 = ap10, pp11 = ap10 * dt + ap11 + 0.1;
    const double S = pp00 + 0.5;                 // :151
    const double k0 = pp00 / S, k1 = pp10 / S;   // :154  P_pred C^T S^-1
    // P = (I - K C) P_pred                       // :160
    p00 = (1.0 - k0) * pp00;
    p01 = (1.0 - k0) * pp01;
    p10 = -k1 * pp00 + pp10;
    p11 = -k1 * pp01 + pp11;
    if (out != nullptr) {
      out[2 * k] = k0;
      out[2 * k + 1] = k1;
    }
  }
  if (P_out != nullptr) {
    P_out[0] = p00; P_out[1] = p01; P_out[2] = p10; P_out[3] = p11;
  }
}

// ------------------------------------------------------------------------------------------------
// Exactness guard of the local nearest-waypoint search.
//
// The kernels find the nearest waypoint by descending from the previous step's answer until both index
// neighbours are no closer, i.e. until the car position p lies in the slab
//   slab(j) = { p : |p - w_j| <= |p - w_{j-1}|  and  |p - w_j| <= |p - w_{j+1}| }.
// j is the GLOBAL argmin unless some other waypoint i is closer, i.e. unless p is in the half plane
//   H_i = { p : |p - w_i| <= |p - w_j| }.
// guard(j) = min over i not in {j-1, j, j+1} of dist(w_j, slab(j) /\ H_i): every p of slab(j) with
// |p - w_j| < guard(j) has j as its exact global argmin.  w_j is in slab(j) but not in H_i, so the point
// of slab(j) /\ H_i nearest to w_j lies on the bisector line of (w_i, w_j) restricted to the slab:
// a 1-D interval problem.  O(n^2) on the host, once per track.
// ------------------------------------------------------------------------------------------------
// Rival lists extend the guard to tracks that run close to themselves (skidpad: the second lap of a circle lies on
// top of the first, so guard(j) ~ 0 all along it).  rivals(j) = every non-neighbour i with
// dist(w_j, slab(j) /\ H_i) < reach.  For p in slab(j) with |p - w_j| < reach the global argmin i* either is j or
// satisfies p in H_{i*}, hence dist(w_j, slab(j) /\ H_{i*}) <= |p - w_j| < reach, i.e. i* is in rivals(j): the exact
// argmin is the best of {j} U rivals(j) -- a handful of evaluations instead of a scan over all n waypoints.
constexpr int kMaxRivals = 1024;   // longer lists fall back to the exact scan (a list is still cheaper than 32 scans)

// Twin triples: on a track that drives the same line twice (skidpad circles) the closest rival of waypoint j is always
// its twin t(j) on the other lap or one of t's index neighbours, while every OTHER waypoint needs the car to be metres
// away before it can win.  twin[j] = (t, squared guard_twin) with guard_twin(j) = the smallest reach-in distance over all
// waypoints outside {j - 1, j, j + 1} U {t - 1, t, t + 1} (cyclic indices, as the kernels' padded table reads them):
// for p in slab(j) with |p - w_j| < guard_twin(j) the exact argmin is the best of j and the twin triple -- three
// evaluations from shared memory instead of a walk through the sorted rival list.  Same exactness argument as the
// guard itself (only the excluded set is larger, and it is evaluated exhaustively).
//
// Layout of rival_list: every list starts at an EVEN entry (the fp32 kernel reads entries in pairs, one 16-byte load)
// and is followed by two or three SENTINEL entries (waypoint 0, reach-in distance +inf) such that the pair after the
// last pair holding a real entry starts with a sentinel: the device loop needs no count, it stops at the first entry
// the car is too close to w_j for -- at the latest the sentinel.  rival_range[j] = (first entry, number of REAL
// entries), count < 0: no usable list.  One more sentinel pair closes the array (the loop loads one pair ahead).
static void compute_guards(const std::vector<double>& x, const std::vector<double>& y, bool closed,
                           std::vector<double>* guard, std::vector<int2>* rival_range,
                           std::vector<int2>* rival_list, std::vector<int2>* twin) {
  const int n = static_cast<int>(x.size());
  const int2 sentinel = make_int2(0, 0x7f800000);
  twin->assign(n, make_int2(-1, 0));
  guard->assign(n, 0.0);
  rival_range->assign(n, make_int2(0, -1));
  rival_list->clear();
  const double reach2 = kRivalReach * kRivalReach;
  std::vector<std::pair<double, int32_t>> mine;   // (squared reach-in distance, waypoint)
  for (int j = 0; j < n; ++j) {
    int nb[2];
    int n_nb = 0;
    if (closed) {
      nb[n_nb++] = (j + n - 1) % n;
      nb[n_nb++] = (j + 1) % n;
    } else {
      if (j > 0) nb[n_nb++] = j - 1;
      if (j < n - 1) nb[n_nb++] = j + 1;
    }
    double ex[2], ey[2], eh[2];
    bool degenerate = false;
    for (int s = 0; s < n_nb; ++s) {
      ex[s] = x[nb[s]] - x[j];
      ey[s] = y[nb[s]] - y[j];
      eh[s] = 0.5 * (ex[s] * ex[s] + ey[s] * ey[s]);
      if (eh[s] == 0.0) degenerate = true;  // repeated waypoint: the descent cannot step over it
    }
    double best2 = INFINITY;
    if (degenerate) best2 = 0.0;
    mine.clear();
    bool lists_ok = !degenerate;
    for (int i = 0; i < n; ++i) {
      if (i == j || (n_nb > 0 && i == nb[0]) || (n_nb > 1 && i == nb[1])) continue;
      const double wx = x[i] - x[j], wy = y[i] - y[j];
      const double D2 = wx * wx + wy * wy;
      double cand2;
      if (D2 == 0.0) {  // coincident with a non-neighbour: the index tie rule decides, the rival check applies it
        cand2 = 0.0;
      } else {
        if (0.25 * D2 >= best2 && 0.25 * D2 >= reach2) continue;
        const double D = std::sqrt(D2);
        // bisector: q(t) = m + t u, m - w_j = (wx, wy) / 2, u = (-wy, wx) / D
        const double ux = -wy / D, uy = wx / D;
        double lo = -INFINITY, hi = INFINITY;
        bool empty = false;
        for (int s = 0; s < n_nb; ++s) {
          // (q - w_j) . e_s <= |e_s|^2 / 2   <=>   t (u . e_s) <= |e_s|^2 / 2 - (m - w_j) . e_s
          const double a = ux * ex[s] + uy * ey[s];
          const double b = eh[s] - 0.5 * (wx * ex[s] + wy * ey[s]);
          if (a > 0.0) hi = std::min(hi, b / a);
          else if (a < 0.0) lo = std::max(lo, b / a);
          else if (b < 0.0) empty = true;
        }
        if (empty || lo > hi) continue;
        double t = 0.0;
        if (lo > 0.0) t = lo;
        else if (hi < 0.0) t = hi;
        cand2 = 0.25 * D2 + t * t;
      }
      best2 = std::min(best2, cand2);
      if (cand2 < reach2) {
        if (static_cast<int>(mine.size()) < kMaxRivals) mine.emplace_back(cand2, i);
        else lists_ok = false;
      }
    }
    (*guard)[j] = std::sqrt(best2);  // may be +inf (e.g. a straight line)
    if (lists_ok) {
      (*rival_range)[j] = make_int2(static_cast<int>(rival_list->size()), static_cast<int>(mine.size()));
      // sorted by the distance from w_j at which the rival can first win: the device walks a list only up to the
      // car's own distance from w_j (stored 1e-4 m short, which covers the fp32 rounding of waypoints and offsets)
      std::sort(mine.begin(), mine.end());
      if (!mine.empty() && n >= 8) {
        const int t = mine.front().second;
        double gt2 = reach2;                       // everything not listed needs at least the lists' reach
        for (const auto& m : mine) {
          const int di = ((m.second - t) % n + n) % n;
          if (di == 0 || di == 1 || di == n - 1) continue;     // the twin triple itself (cyclic)
          gt2 = std::min(gt2, m.first);
          break;                                               // sorted: the first one outside the triple is the minimum
        }
        // the triple must not overlap j's own neighbourhood in a way the index tie rule could not see: it may (the
        // kernel compares true indices); shrink like the guard: 1e-4 relative + 1e-4 m
        const double g = std::max(0.0, std::sqrt(gt2) * (1.0 - 1e-4) - 1e-4);
        const float g2f = std::nextafterf(static_cast<float>(g * g), 0.0f);
        if (g2f > 0.0f) (*twin)[j] = make_int2(t, __builtin_bit_cast(int, g2f));
      }
      for (const auto& m : mine) {
        const double c = std::max(0.0, std::sqrt(m.first) - 1e-4);
        const float c2 = std::nextafterf(static_cast<float>(c * c), 0.0f);
        rival_list->push_back(make_int2(m.second, __builtin_bit_cast(int, c2)));
      }
      if (mine.size() % 2 != 0) rival_list->push_back(sentinel);
      rival_list->push_back(sentinel);
      rival_list->push_back(sentinel);
    }
  }
  rival_list->push_back(sentinel);
  rival_list->push_back(sentinel);
}

}  // namespace bgr

using namespace bgr;

// ================================================================================================
// library
// ================================================================================================
extern "C" int bgr_abi_version(void) { return BGR_ABI_VERSION; }

extern "C" const char* bgr_last_error_string(void) { return g_error; }

extern "C" int bgr_device_count(int* out_count) {
  if (out_count == nullptr) {
    set_error("bgr_device_count: out_count is NULL");
    return BGR_ERR_BAD_ARG;
  }
  *out_count = 0;
  BGR_CUDA(cudaGetDeviceCount(out_count));
  return BGR_OK;
}

extern "C" int64_t bgr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int bgr_last_kernel_ms(double* out_ms) {
  if (out_ms == nullptr) {
    set_error("bgr_last_kernel_ms: out_ms is NULL");
    return BGR_ERR_BAD_ARG;
  }
  if (!g_last.valid) {
    set_error("bgr_last_kernel_ms: no kernel has been launched by this thread");
    return BGR_ERR_BAD_ARG;
  }
  BGR_CUDA(cudaEventSynchronize(g_last.stop));
  float ms = 0.0f;
  BGR_CUDA(cudaEventElapsedTime(&ms, g_last.start, g_last.stop));
  *out_ms = static_cast<double>(ms);
  return BGR_OK;
}

extern "C" int bgr_lqg_design(double dt, int32_t n_steps, double* h_lqr_gain, double* h_kf_gains) {
  if (!(dt > 0.0) || n_steps < 0 || h_lqr_gain == nullptr) {
    set_error("bgr_lqg_design: need dt > 0, n_steps >= 0, h_lqr_gain != NULL");
    return BGR_ERR_BAD_ARG;
  }
  lqr_gain_host(dt, h_lqr_gain);
  if (h_kf_gains != nullptr && n_steps > 0) kalman_gains_host(dt, n_steps, h_kf_gains);
  return BGR_OK;
}

extern "C" int bgr_lqg_covariance(double dt, int32_t n_steps, double* h_P) {
  if (!(dt > 0.0) || n_steps < 0 || h_P == nullptr) {
    set_error("bgr_lqg_covariance: need dt > 0, n_steps >= 0, h_P != NULL");
    return BGR_ERR_BAD_ARG;
  }
  kalman_gains_host(dt, n_steps, nullptr, h_P);
  return BGR_OK;
}

#ifndef BGR_BUILD_ID
#define BGR_BUILD_ID "unknown"
#endif
extern "C" const char* bgr_build_id(void) { return BGR_BUILD_ID; }

extern "C" int bgr_stream_synchronize(void* stream) {
  BGR_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return BGR_OK;
}

extern "C" int bgr_stream_create(int device, void** out_stream) {
  if (out_stream == nullptr) {
    set_error("bgr_stream_create: out_stream is NULL");
    return BGR_ERR_BAD_ARG;
  }
  *out_stream = nullptr;
  int prev = 0;
  BGR_CUDA(cudaGetDevice(&prev));
  BGR_CUDA(cudaSetDevice(device));
  cudaStream_t st = nullptr;
  const cudaError_t err = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaSetDevice(prev);
  BGR_CUDA(err);
  *out_stream = st;
  return BGR_OK;
}

extern "C" int bgr_stream_destroy(void* stream) {
  if (stream != nullptr) BGR_CUDA(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
  return BGR_OK;
}

// ================================================================================================
// track handle
// ================================================================================================
template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
  *dst = nullptr;
  if (src.empty()) return BGR_OK;
  BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(dst), src.size() * sizeof(T)));
  BGR_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return BGR_OK;
}

extern "C" void bgr_track_destroy(bgr_track_t* t) {
  if (t == nullptr) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(t->device);
  cudaFree(t->xy_f);
  cudaFree(t->attr_f);
  cudaFree(t->xy_d);
  cudaFree(t->attr_d);
  cudaFree(t->guard2);
  cudaFree(t->rival_range);
  cudaFree(t->rival_list);
  cudaFree(t->twin);
  cudaFree(t->scan_blk);
  cudaFree(t->cone_range);
  cudaFree(t->cone_near);
  cudaFree(t->cone_list);
  cudaFree(t->cones);
  cudaFree(t->cones_f);
  for (auto& k : t->kf_tables) {
    cudaFree(k.d_gain);
    cudaFree(k.d_gain_f);
  }
  cudaSetDevice(prev);
  delete t;
}

extern "C" int bgr_track_create(const bgr_track_desc_t* desc, int device, bgr_track_t** out_track) {
  if (out_track == nullptr) {
    set_error("bgr_track_create: out_track is NULL");
    return BGR_ERR_BAD_ARG;
  }
  *out_track = nullptr;
  if (desc == nullptr || desc->h_x == nullptr || desc->h_y == nullptr || desc->h_kappa == nullptr) {
    set_error("bgr_track_create: desc / h_x / h_y / h_kappa is NULL");
    return BGR_ERR_BAD_ARG;
  }
  const int n = desc->n;
  if (n < 2 || (desc->closed && n < 3)) {
    set_error("bgr_track_create: need n >= 2 waypoints (>= 3 when closed), got %d", n);
    return BGR_ERR_BAD_ARG;
  }
  if (desc->n_cones < 0 || desc->n_cones > BGR_MAX_CONES || (desc->n_cones > 0 && desc->h_cones_xy == nullptr)) {
    set_error("bgr_track_create: n_cones must be in [0, %d] with h_cones_xy set, got %d", BGR_MAX_CONES,
              desc->n_cones);
    return BGR_ERR_BAD_ARG;
  }
  for (int i = 0; i < n; ++i) {
    if (!std::isfinite(desc->h_x[i]) || !std::isfinite(desc->h_y[i]) || !std::isfinite(desc->h_kappa[i])) {
      set_error("bgr_track_create: non-finite waypoint %d", i);
      return BGR_ERR_BAD_ARG;
    }
  }
  int n_dev = 0;
  BGR_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    set_error("bgr_track_create: device %d out of range (%d CUDA devices)", device, n_dev);
    return BGR_ERR_BAD_ARG;
  }
  DeviceInfo di;
  int rc = device_info(device, &di);
  if (rc != BGR_OK) return rc;
  const int n_pad = (n + 12 + 3) & ~3;   // fp32 xy table: 4 cyclic entries in front, >= 8 behind (bgr_rollout_f32.cuh)
  if (rollout_smem_bytes_xy_only(n_pad) > static_cast<size_t>(di.max_smem_optin)) {
    set_error("bgr_track_create: %d waypoints need %zu B of shared memory per CTA, device offers %d", n,
              rollout_smem_bytes_xy_only(n_pad), di.max_smem_optin);
    return BGR_ERR_TOO_LARGE;
  }

  bgr_track* t = new (std::nothrow) bgr_track();
  if (t == nullptr) {
    set_error("bgr_track_create: out of host memory");
    return BGR_ERR_ALLOC;
  }
  t->device = device;
  t->n = n;
  t->n_pad = n_pad;
  t->closed = desc->closed ? 1 : 0;
  t->n_cones = desc->n_cones;
  t->cone_reach = desc->cone_reach;
  t->x.assign(desc->h_x, desc->h_x + n);
  t->y.assign(desc->h_y, desc->h_y + n);
  t->kappa.assign(desc->h_kappa, desc->h_kappa + n);

  // path heading: forward difference, backward at the last index (core/controllers.py:361-368) --
  // no wrap-around even on a closed lap, exactly like the reference
  t->path_yaw.resize(n);
  for (int j = 0; j < n; ++j) {
    double dx, dy;
    if (j < n - 1) {
      dx = t->x[j + 1] - t->x[j];
      dy = t->y[j + 1] - t->y[j];
    } else {
      dx = t->x[j] - t->x[j - 1];
      dy = t->y[j] - t->y[j - 1];
    }
    t->path_yaw[j] = std::atan2(dy, dx);
  }
  std::vector<int2> rival_range;
  std::vector<int2> rival_list;
  std::vector<int2> twin;
  compute_guards(t->x, t->y, t->closed != 0, &t->guard_radius, &rival_range, &rival_list, &twin);
  for (const auto& tw : twin) t->twin_count += tw.x >= 0 ? 1 : 0;

  t->ds_min = INFINITY;
  t->ds_max = 0.0;
  const int n_seg = t->closed ? n : n - 1;
  for (int j = 0; j < n_seg; ++j) {
    const int k = (j + 1) % n;
    const double ds = std::hypot(t->x[k] - t->x[j], t->y[k] - t->y[j]);
    t->ds_min = std::min(t->ds_min, ds);
    t->ds_max = std::max(t->ds_max, ds);
  }
  {
    std::vector<double> g = t->guard_radius;
    std::sort(g.begin(), g.end());
    t->guard_min = g.front();
    t->guard_median = g[g.size() / 2];
    t->guard_tight_frac = static_cast<double>(std::lower_bound(g.begin(), g.end(), kTightGuard) - g.begin()) /
                          static_cast<double>(g.size());
  }

  std::vector<float2> xy_f(n_pad);
  std::vector<float4> attr_f(n_pad);
  std::vector<double2> xy_d(n_pad);
  std::vector<D4> attr_d(n_pad);
  std::vector<float> guard2(n_pad);
  for (int j = 0; j < n_pad; ++j) {
    const int s = j % n;   // cyclic pad
    const double sn = std::sin(t->path_yaw[s]), cs = std::cos(t->path_yaw[s]);
    xy_d[j] = make_double2(t->x[s], t->y[s]);
    attr_d[j] = D4{t->kappa[s], t->path_yaw[s], sn, cs};
    const int sf = ((j - 4) % n + n) % n;   // fp32 xy table entry j holds waypoint j - 4 (front pad)
    xy_f[j] = make_float2(static_cast<float>(t->x[sf]), static_cast<float>(t->y[sf]));
    attr_f[j] = make_float4(static_cast<float>(t->kappa[s]), static_cast<float>(t->path_yaw[s]),
                            static_cast<float>(sn), static_cast<float>(cs));
    // shrink by 1e-4 relative + 1e-4 m: the in-slab test itself runs in fp32 on rounded waypoints
    double g = t->guard_radius[s];
    g = std::isfinite(g) ? std::max(0.0, g * (1.0 - 1e-4) - 1e-4) : 1e18;
    guard2[j] = static_cast<float>(std::min(g * g, 1e30));
  }
  // Bounding circles of blocks of 32 consecutive waypoints, of the ROUNDED waypoints the fp32 kernels read: centre =
  // the block's mean (as floats), radius = the largest distance of a member from that float centre, rounded up.  The
  // exact scan evaluates a block only if |p - centre| - radius does not exceed the distance already in hand.
  std::vector<float4> scan_blk((n + 31) / 32);
  for (int b = 0; b < static_cast<int>(scan_blk.size()); ++b) {
    const int lo = 32 * b, hi = std::min(n, lo + 32);
    double mx = 0.0, my = 0.0;
    for (int i = lo; i < hi; ++i) {
      mx += static_cast<double>(static_cast<float>(t->x[i]));
      my += static_cast<double>(static_cast<float>(t->y[i]));
    }
    const float cx = static_cast<float>(mx / (hi - lo)), cy = static_cast<float>(my / (hi - lo));
    double r = 0.0;
    for (int i = lo; i < hi; ++i)
      r = std::max(r, std::hypot(static_cast<double>(static_cast<float>(t->x[i])) - static_cast<double>(cx),
                                 static_cast<double>(static_cast<float>(t->y[i])) - static_cast<double>(cy)));
    scan_blk[b] = make_float4(cx, cy, std::nextafterf(static_cast<float>(r * (1.0 + 1e-6) + 1e-6), INFINITY), 0.0f);
  }

  // rival lists: (first, count) per waypoint; count < 0 = no usable list (exact scan instead)
  for (int j = 0; j < n; ++j) {
    if (rival_range[j].y >= 0) {
      t->rival_total += rival_range[j].y;
      t->rival_max = std::max(t->rival_max, rival_range[j].y);
    } else {
      ++t->rival_unlisted;
    }
  }

  // cone candidate lists: cones within cone_reach of waypoint j
  std::vector<int32_t> cone_range(n, 0), cone_list;
  std::vector<float> cone_near(n, 0.0f);   // distance from waypoint j to its closest listed cone, rounded DOWN
  std::vector<double2> cones(t->n_cones);
  for (int c = 0; c < t->n_cones; ++c) cones[c] = make_double2(desc->h_cones_xy[2 * c], desc->h_cones_xy[2 * c + 1]);
  if (t->n_cones > 0) {
    for (int j = 0; j < n; ++j) {
      const int first = static_cast<int>(cone_list.size());
      int cnt = 0;
      double nearest_cone = desc->cone_reach;    // (an empty list: nothing closer than the reach)
      for (int c = 0; c < t->n_cones; ++c) {
        const double dcw = std::hypot(cones[c].x - t->x[j], cones[c].y - t->y[j]);
        if (dcw <= desc->cone_reach) {
          cone_list.push_back(c);
          ++cnt;
          nearest_cone = std::min(nearest_cone, dcw);
        }
      }
      cone_near[j] = std::nextafter(static_cast<float>(std::max(0.0, nearest_cone - 1e-4)), 0.0f);
      if (first >= (1 << 21)) {
        delete t;
        set_error("bgr_track_create: cone candidate lists overflow (cone_reach %.3f too large)", desc->cone_reach);
        return BGR_ERR_TOO_LARGE;
      }
      cone_range[j] = (first << 10) | cnt;  // cnt <= BGR_MAX_CONES = 512 < 1024
    }
    if (cone_list.empty()) cone_list.push_back(0);
  }

  int prev = 0;
  cudaGetDevice(&prev);
  rc = check_cuda(cudaSetDevice(device), "cudaSetDevice");
  if (rc == BGR_OK) rc = upload(&t->xy_f, xy_f);
  if (rc == BGR_OK) rc = upload(&t->scan_blk, scan_blk);
  if (rc == BGR_OK) rc = upload(&t->attr_f, attr_f);
  if (rc == BGR_OK) rc = upload(&t->xy_d, xy_d);
  if (rc == BGR_OK) rc = upload(&t->attr_d, attr_d);
  if (rc == BGR_OK) rc = upload(&t->guard2, guard2);
  if (rc == BGR_OK) rc = upload(&t->rival_range, rival_range);
  if (rc == BGR_OK) rc = upload(&t->rival_list, rival_list);
  if (rc == BGR_OK) rc = upload(&t->twin, twin);
  if (rc == BGR_OK && t->n_cones > 0) {
    rc = upload(&t->cone_range, cone_range);
    if (rc == BGR_OK) rc = upload(&t->cone_near, cone_near);
    if (rc == BGR_OK) rc = upload(&t->cone_list, cone_list);
    if (rc == BGR_OK) rc = upload(&t->cones, cones);
    if (rc == BGR_OK) {
      std::vector<float2> cones_f(t->n_cones);
      for (int c = 0; c < t->n_cones; ++c)
        cones_f[c] = make_float2(static_cast<float>(cones[c].x), static_cast<float>(cones[c].y));
      rc = upload(&t->cones_f, cones_f);
    }
  }
  cudaSetDevice(prev);
  if (rc != BGR_OK) {
    bgr_track_destroy(t);
    return rc;
  }
  *out_track = t;
  return BGR_OK;
}

extern "C" int bgr_track_info(const bgr_track_t* t, bgr_track_info_t* out) {
  if (t == nullptr || out == nullptr) {
    set_error("bgr_track_info: NULL argument");
    return BGR_ERR_BAD_ARG;
  }
  out->n = t->n;
  out->closed = t->closed;
  out->n_cones = t->n_cones;
  out->device = t->device;
  out->smem_bytes_fp32 = static_cast<int64_t>(rollout_smem_bytes<float>(t->n_pad));
  out->smem_bytes_fp64 = static_cast<int64_t>(rollout_smem_bytes<double>(t->n_pad));
  out->guard_min = t->guard_min;
  out->guard_median = t->guard_median;
  out->ds_min = t->ds_min;
  out->ds_max = t->ds_max;
  out->rival_total = t->rival_total;
  out->rival_max = t->rival_max;
  out->rival_unlisted = t->rival_unlisted;
  out->twin_count = t->twin_count;
  return BGR_OK;
}

extern "C" int bgr_track_tables(const bgr_track_t* t, double* h_path_yaw, double* h_guard_radius) {
  if (t == nullptr) {
    set_error("bgr_track_tables: track is NULL");
    return BGR_ERR_BAD_ARG;
  }
  if (h_path_yaw != nullptr) std::copy(t->path_yaw.begin(), t->path_yaw.end(), h_path_yaw);
  if (h_guard_radius != nullptr) std::copy(t->guard_radius.begin(), t->guard_radius.end(), h_guard_radius);
  return BGR_OK;
}

// ================================================================================================
// rollout
// ================================================================================================
namespace {

struct StepPlumbing {
  int32_t phases = BGR_PHASE_STEER | BGR_PHASE_ACCEL | BGR_PHASE_MODEL | BGR_PHASE_ERRORS;
  const double* ctrl_in = nullptr;
  double* ctrl_out = nullptr;
  const int32_t* ti_in = nullptr;
  int32_t* ti_out = nullptr;
  double* state_out = nullptr;
  const double* cmd = nullptr;
  const double* params64 = nullptr;
  double* step_out = nullptr;
  bool force_full = false;
};

using LaunchFn = cudaError_t (*)(const RolloutArgs&, int, size_t, int, cudaStream_t, int*);

template <typename Real, bool EXTRAS>
LaunchFn pick(int s, int a, int m) {
  const int key = s * 4 + a * 2 + m;
  switch (key) {
    case 0: return launch_rollout<Real, 0, 0, 0, EXTRAS>;
    case 1: return launch_rollout<Real, 0, 0, 1, EXTRAS>;
    case 2: return launch_rollout<Real, 0, 1, 0, EXTRAS>;
    case 3: return launch_rollout<Real, 0, 1, 1, EXTRAS>;
    case 4: return launch_rollout<Real, 1, 0, 0, EXTRAS>;
    case 5: return launch_rollout<Real, 1, 0, 1, EXTRAS>;
    case 6: return launch_rollout<Real, 1, 1, 0, EXTRAS>;
    default: return launch_rollout<Real, 1, 1, 1, EXTRAS>;
  }
}

int validate_cfg(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, const char* who) {
  if (track == nullptr || cfg == nullptr) {
    set_error("%s: track / cfg is NULL", who);
    return BGR_ERR_BAD_ARG;
  }
  if (cfg->steer_kind < 0 || cfg->steer_kind > 1 || cfg->accel_kind < 0 || cfg->accel_kind > 1 ||
      cfg->model_kind < 0 || cfg->model_kind > 1 || cfg->precision < 0 || cfg->precision > 1) {
    set_error("%s: unknown steer/accel/model/precision selector (%d, %d, %d, %d)", who, cfg->steer_kind,
              cfg->accel_kind, cfg->model_kind, cfg->precision);
    return BGR_ERR_BAD_ARG;
  }
  if (cfg->max_steps < 1 || cfg->laps < 1 || cfg->laps_target < 0 || !(cfg->dt > 0.0) || !(cfg->wheelbase > 0.0) ||
      !(cfg->max_steer > 0.0) || !(cfg->max_accel >= cfg->max_decel) || cfg->hit_radius < 0.0 ||
      cfg->tie_eps < 0.0 || cfg->sigma_xy < 0.0 || cfg->sigma_yaw < 0.0 || cfg->sigma_v < 0.0 ||
      cfg->sigma_cone < 0.0 || cfg->stanley_max_steer_rate < 0.0 || cfg->stanley_alpha < 0.0 || cfg->stanley_alpha > 1.0 ||
      !(cfg->stanley_lookahead >= 0.0)) {
    set_error("%s: invalid cfg (max_steps %d, laps %d, laps_target %d, dt %g, wheelbase %g, max_steer %g, ...)",
              who, cfg->max_steps, cfg->laps, cfg->laps_target, cfg->dt, cfg->wheelbase, cfg->max_steer);
    return BGR_ERR_BAD_ARG;
  }
  if (cfg->replan_horizon != 0) {
    if (cfg->steer_kind != BGR_STEER_PURE_PURSUIT || cfg->replan_horizon < 2 || cfg->replan_stride < 0 ||
        cfg->replan_after < 0 || cfg->planner_lookahead < 0.0 ||
        static_cast<long long>(cfg->replan_horizon) * (cfg->replan_stride > 0 ? cfg->replan_stride : 1) > 0x3fffffffLL) {
      set_error("%s: re-planning replay needs pure pursuit, replan_horizon >= 2, replan_stride >= 0, replan_after >= 0 "
                "(got steer %d, horizon %d, stride %d, after %d)", who, cfg->steer_kind, cfg->replan_horizon,
                cfg->replan_stride, cfg->replan_after);
      return BGR_ERR_BAD_ARG;
    }
  }
  if (static_cast<long long>(track->n) * cfg->laps > 0x3fffffffLL) {
    set_error("%s: n * laps overflows the target index", who);
    return BGR_ERR_BAD_ARG;
  }
  if (cfg->hit_radius > 0.0 && track->n_cones > 0) {
    const double need = cfg->hit_radius + 6.0 * cfg->sigma_cone + BGR_CONE_GUARD;
    if (track->cone_reach < need) {
      set_error("%s: track cone_reach %.3f < hit_radius + 6 sigma_cone + BGR_CONE_GUARD = %.3f", who,
                track->cone_reach, need);
      return BGR_ERR_BAD_ARG;
    }
  }
  return BGR_OK;
}

struct Plan {
  RolloutArgs A;
  LaunchFn fn = nullptr;
  size_t smem = 0;
  int smem_optin = 0;   // the device's opt-in shared-memory maximum (per-function attribute of every step kernel)
  long long wave = 0;   // episodes the device runs at once with this kernel (SMs x resident CTAs x threads)
  // dynamic episode scheduling (full fp32 build, plain rollouts): one wave of CTAs pulls rows from a device counter
  bool dynamic = false;
  long long wave_dyn = 0;
  // launch order by target speed (bgr_misc.cu): ranges of at least kOrderMin episodes that can end early
  bool order = false;
};
constexpr long long kOrderMin = 4096;

// Validate the configuration, fill the kernel arguments that do not depend on the episode range, pick the kernel
// specialisation and make sure the LQG tables are resident (cached on the track handle).
int make_plan(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, const StepPlumbing& sp, bool want_traj,
              cudaStream_t stream, Plan* plan) {
  int rc = validate_cfg(track, cfg, "bgr_rollout");
  if (rc != BGR_OK) return rc;
  DeviceInfo di;
  rc = device_info(track->device, &di);
  if (rc != BGR_OK) return rc;

  plan->smem_optin = di.max_smem_optin;
  const bool f64 = cfg->precision == BGR_PRECISION_FP64;
  const bool cones_on = cfg->hit_radius > 0.0 && track->n_cones > 0;
  const bool noise_on = cfg->sigma_xy > 0.0 || cfg->sigma_yaw > 0.0 || cfg->sigma_v > 0.0;
  const bool shadowed = (cfg->flags & BGR_CFG_STANLEY_SHADOWED) != 0 && cfg->steer_kind == BGR_STEER_STANLEY;
  const bool replan = cfg->replan_horizon > 0;
  const bool full = f64 || cones_on || noise_on || want_traj || (cfg->flags & BGR_CFG_AUDIT) || sp.force_full ||
                    shadowed || replan;

  plan->smem = f64 ? rollout_smem_bytes<double>(track->n_pad) : rollout_smem_bytes<float>(track->n_pad);
  bool attr_in_smem = true;
  if (!f64) {
    static const int force = [] { const char* e = std::getenv("BGR_ATTR_IN_SMEM"); return e ? std::atoi(e) : -1; }();
    // (the full build is register-limited to 3 CTAs per SM: it can afford 4/3 of the shared memory per CTA)
    attr_in_smem = force >= 0 ? force != 0 : plan->smem <= (full ? kSmemAllTablesLimit * 4 / 3 : kSmemAllTablesLimit);
    if (!attr_in_smem) plan->smem = rollout_smem_bytes_xy_only(track->n_pad);
  }
  if (plan->smem > static_cast<size_t>(di.max_smem_optin)) {
    set_error("bgr_rollout: track needs %zu B of shared memory per CTA in this precision, device offers %d",
              plan->smem, di.max_smem_optin);
    return BGR_ERR_TOO_LARGE;
  }

  if (replan && sp.ctrl_in != nullptr) {
    set_error("bgr_step_host: re-planning replay is a rollout mode (the planner's path and index are episode state)");
    return BGR_ERR_BAD_ARG;
  }
  RolloutArgs& A = plan->A;
  std::memset(&A, 0, sizeof(A));
  A.trk = track->view();
  A.episode_id_base = cfg->episode_id_base;
  A.max_steps = cfg->max_steps;
  A.laps = cfg->laps;
  A.laps_target = cfg->laps_target;
  A.lap_stop = cfg->laps_target > 0 ? cfg->laps_target : INT32_MAX;
  A.n_total = cfg->steer_kind == BGR_STEER_PURE_PURSUIT ? track->n * cfg->laps : track->n;
  A.ti_chunk_max = track->n >= 8 ? A.n_total - 4 : INT32_MIN;
  A.dt = cfg->dt;
  A.wheelbase = cfg->wheelbase;
  A.l_f = cfg->l_f;
  A.l_r = cfg->l_r;
  A.max_steer = cfg->max_steer;
  A.inv_dt = 1.0 / cfg->dt;
  A.inv_wheelbase = 1.0 / cfg->wheelbase;
  A.tan_max_steer = std::tan(cfg->max_steer);
  A.max_accel = cfg->max_accel;
  A.max_decel = cfg->max_decel;
  A.hit_radius = cfg->hit_radius;
  A.tie_eps = cfg->tie_eps > 0.0 ? cfg->tie_eps : 2e-5;
  A.sigma_xy = cfg->sigma_xy;
  A.sigma_yaw = cfg->sigma_yaw;
  A.sigma_v = cfg->sigma_v;
  A.sigma_cone = cfg->sigma_cone;
  A.f.dt = static_cast<float>(cfg->dt);
  A.f.dt_lo = static_cast<float>(cfg->dt - static_cast<double>(static_cast<float>(cfg->dt)));
  A.f.inv_dt = static_cast<float>(1.0 / cfg->dt);
  A.f.WB = static_cast<float>(cfg->wheelbase);
  A.f.inv_WB = static_cast<float>(1.0 / cfg->wheelbase);
  A.f.l_f = static_cast<float>(cfg->l_f);
  A.f.l_r = static_cast<float>(cfg->l_r);
  A.f.max_steer = static_cast<float>(cfg->max_steer);
  A.f.tan_max_steer = static_cast<float>(std::tan(cfg->max_steer));
  A.f.max_accel = static_cast<float>(cfg->max_accel);
  A.f.max_decel = static_cast<float>(cfg->max_decel);
  A.f.hit_radius = static_cast<float>(cfg->hit_radius);
  A.f.tie_eps = static_cast<float>(A.tie_eps);
  A.f.sigma_xy = static_cast<float>(cfg->sigma_xy);
  A.f.sigma_yaw = static_cast<float>(cfg->sigma_yaw);
  A.f.sigma_v = static_cast<float>(cfg->sigma_v);
  A.f.sigma_cone = static_cast<float>(cfg->sigma_cone);
  A.stanley_shadowed = shadowed ? 1 : 0;
  A.st_rate = cfg->stanley_max_steer_rate > 0.0 ? cfg->stanley_max_steer_rate : 30.0 * 3.14159265358979323846 / 180.0;
  A.st_alpha = cfg->stanley_alpha > 0.0 ? cfg->stanley_alpha : 0.5;
  A.f.st_max_delta = static_cast<float>(A.st_rate * cfg->dt);
  A.f.st_alpha = static_cast<float>(A.st_alpha);
  A.st_lookahead = shadowed ? cfg->stanley_lookahead : 0.0;
  A.f.st_lookahead = static_cast<float>(A.st_lookahead);
  A.replan_M = replan ? cfg->replan_horizon : 0;
  A.replan_q = cfg->replan_stride > 0 ? cfg->replan_stride : 1;
  A.replan_after = cfg->replan_after;
  A.plan_Lf = cfg->planner_lookahead > 0.0 ? cfg->planner_lookahead : 5.5;
  A.f.plan_Lf = static_cast<float>(A.plan_Lf);
  A.seed = cfg->noise_seed;
  A.cones_on = cones_on ? 1 : 0;
  A.attr_in_smem = attr_in_smem ? 1 : 0;
  A.noise_on = noise_on ? 1 : 0;
  // With pose noise, or on a track that mostly runs inside a tight guard (the skidpad's circles are driven twice: 90 % of
  // its waypoints), some lane of nearly every warp leaves the five-point window of the nearest search, so the warp pays
  // for the window AND the complete search: go to the complete search directly.  Same result either way.
  A.search_direct = (noise_on || track->guard_tight_frac > 0.25) ? 1 : 0;
  {   // (diagnostic override: BGR_SEARCH_DIRECT=0 / 1)
    static const int force = [] { const char* e = std::getenv("BGR_SEARCH_DIRECT"); return e ? std::atoi(e) : -1; }();
    if (force >= 0) A.search_direct = force != 0;
  }
  A.phases = sp.phases;
  A.ctrl_in = sp.ctrl_in;
  A.ctrl_out = sp.ctrl_out;
  A.ti_in = sp.ti_in;
  A.ti_out = sp.ti_out;
  A.state_out = sp.state_out;
  A.cmd = sp.cmd;
  A.params64 = sp.params64;
  A.step_out = sp.step_out;

  // LQG: LQR gain + the shared, data-independent Kalman gain sequence (uploaded once per (dt, length) and kept
  // on the track handle until it is destroyed)
  if (cfg->accel_kind == BGR_ACCEL_LQG) {
    int kf_len = cfg->max_steps;
    if (sp.ctrl_in != nullptr) kf_len = 4096;  // single-step API: index carried by the caller, clamped in-kernel
    std::lock_guard<std::mutex> lock(track->kf_mu);
    const bgr_track::KfTable* hit = nullptr;
    for (const auto& t : track->kf_tables)
      if (t.dt == cfg->dt && t.len >= kf_len) hit = &t;
    if (hit == nullptr) {
      const int len = std::max(kf_len, 4096);
      std::vector<double> gains(static_cast<size_t>(len) * 2);
      kalman_gains_host(cfg->dt, len, gains.data());
      std::vector<float2> gains_f(static_cast<size_t>(len));
      for (int i = 0; i < len; ++i)
        gains_f[i] = make_float2(static_cast<float>(gains[2 * i]), static_cast<float>(gains[2 * i + 1]));
      bgr_track::KfTable t;
      t.dt = cfg->dt;
      t.len = len;
      lqr_gain_host(cfg->dt, t.K);             // Riccati iteration: once per (track, dt), not per call
      BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(&t.d_gain), gains.size() * sizeof(double)));
      BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(&t.d_gain_f), gains_f.size() * sizeof(float2)));
      BGR_CUDA(cudaMemcpy(t.d_gain, gains.data(), gains.size() * sizeof(double), cudaMemcpyHostToDevice));
      BGR_CUDA(cudaMemcpy(t.d_gain_f, gains_f.data(), gains_f.size() * sizeof(float2), cudaMemcpyHostToDevice));
      track->kf_tables.push_back(t);
      hit = &track->kf_tables.back();
    }
    A.lqr_k0 = hit->K[0];
    A.lqr_k1 = hit->K[1];
    A.f.K0 = static_cast<float>(hit->K[0]);
    A.f.K1 = static_cast<float>(hit->K[1]);
    A.kf_gain = hit->d_gain;
    A.kf_gain_f = hit->d_gain_f;
    A.kf_len = hit->len;
  }

  if (f64) plan->fn = pick<double, true>(cfg->steer_kind, cfg->accel_kind, cfg->model_kind);
  else if (full) plan->fn = pick<float, true>(cfg->steer_kind, cfg->accel_kind, cfg->model_kind);
  else plan->fn = pick<float, false>(cfg->steer_kind, cfg->accel_kind, cfg->model_kind);
  (void)stream;
  // how many episodes the device holds at once with this kernel (occupancy query through the launcher)
  int resident = 0;
  BGR_CUDA(plan->fn(A, 0, plan->smem, plan->smem_optin, nullptr, &resident));
  plan->wave = static_cast<long long>(di.sm_count) * std::max(resident, kRolloutThreads);
  // Dynamic scheduling (BGR_CFG_DYNAMIC_SCHEDULE) is for batches whose episodes end at very different times AND whose
  // longest episodes are rare enough for the refilled lanes to matter; plain rollouts of the full fp32 build only (the
  // single-step plumbing and an explicit two-phase request keep the static schedule), used by the launch sites for
  // batches larger than one wave.  Measured on config 5: neutral -- that workload is bound by its
  // per-step cost, not by stragglers -- so it is opt-in.
  plan->dynamic = (cfg->flags & BGR_CFG_DYNAMIC_SCHEDULE) != 0 && full && !f64 && sp.ctrl_in == nullptr &&
                  !sp.force_full && cfg->split_steps == 0;
  // Launch order by target speed: for rollouts that can end before the step cap for a reason that scales with the speed
  // (end of an open path, lap target) -- a warp then holds episodes of like length.  Plain fp32 rollouts only; the
  // output rows are the same either way (BGR_CFG_KEEP_ORDER switches it off).
  // (Not combined with the dynamic schedule: measured on config 5, a queue that hands out the rows slowest-first is
  // slower than the plain queue for the heavy-tailed Stanley kinds -- 1.98e10 vs 2.21e10 sim-steps/s overall -- and
  // slower than the ordered static launch for pure pursuit.)
  plan->order = (cfg->flags & BGR_CFG_KEEP_ORDER) == 0 && !f64 && !plan->dynamic && sp.ctrl_in == nullptr &&
                sp.params64 == nullptr && !sp.force_full && cfg->split_steps == 0 &&
                (track->closed == 0 || cfg->laps_target > 0);
  if (plan->dynamic) {
    RolloutArgs Q = A;
    Q.queue = reinterpret_cast<int32_t*>(uintptr_t(16));   // (only selects the instantiation: nothing is launched)
    BGR_CUDA(plan->fn(Q, 0, plan->smem, plan->smem_optin, nullptr, &resident));
    plan->wave_dyn = static_cast<long long>(di.sm_count) * std::max(resident, kRolloutThreads);
  }
  return BGR_OK;
}

// Launch the step kernel for `count` episodes whose rows start at the given pointers; `id_offset` is their global
// position within the call (noise streams are keyed by the global episode id).  `max_steps` may cap the launch below
// the configured step count (first phase of a two-phase run); `d_idx` / `d_count` (both or neither) make thread t
// work on row d_idx[t] for t < *d_count (second phase).
int launch_range(const Plan& plan, long long id_offset, long long count, const float* d_params, const double* d_init,
                 float* d_metrics, int32_t* d_status, double* d_traj, int max_steps, const int32_t* d_idx,
                 const int32_t* d_count, cudaStream_t stream, int32_t* d_queue = nullptr) {
  RolloutArgs A = plan.A;
  A.params = d_params;
  A.init = d_init;
  A.metrics = d_metrics;
  A.status = d_status;
  A.traj = d_traj;
  A.n_episodes = count;
  A.episode_id_base += id_offset;
  A.traj_stride = plan.A.max_steps;          // rows of one episode in the trajectory dump
  A.max_steps = max_steps;
  A.idx = d_idx;
  A.count = d_count;
  int grid = static_cast<int>((count + kRolloutThreads - 1) / kRolloutThreads);
  if (d_queue != nullptr && plan.dynamic && d_idx == nullptr && count > plan.wave_dyn) {
    // dynamic schedule: one wave of CTAs, rows pulled from a zeroed counter
    BGR_CUDA(cudaMemsetAsync(d_queue, 0, sizeof(int32_t), stream));
    A.queue = d_queue;
    grid = static_cast<int>(std::min<long long>(grid, plan.wave_dyn / kRolloutThreads));
    {   // (tuning: BGR_DYN_GRID_PCT=<percent of one wave>)
      static const int pct = [] { const char* e = std::getenv("BGR_DYN_GRID_PCT"); return e ? std::atoi(e) : 100; }();
      if (pct > 0 && pct != 100) grid = std::max(1, static_cast<int>(static_cast<long long>(grid) * pct / 100));
    }
  }
  cudaError_t err = plan.fn(A, grid, plan.smem, plan.smem_optin, stream, nullptr);
  if (err == cudaSuccess) count_launch();
  return check_cuda(err, "rollout kernel launch");
}

// One rollout of `count` episodes, in one launch or -- when cfg.split_steps asks for it -- in two phases: a first
// launch capped at split_steps, then every episode that is still running (status TIMEOUT at the cap) is re-run from
// scratch to the full step count in a second launch in which the survivors are densely packed.  Episodes are
// deterministic functions of their own row, so the results are bit-identical to the single launch; what changes is
// that a warp no longer idles 31 lanes for the one episode in it that runs ten times longer than the rest.
// `d_work` is scratch for the survivor list: int32[count + 1] (last element = number of survivors).
int launch_rollout_phases(const Plan& plan, int split_steps, long long id_offset, long long count, const float* d_params,
                          const double* d_init, float* d_metrics, int32_t* d_status, double* d_traj, int32_t* d_work,
                          cudaStream_t stream, int32_t* d_queue = nullptr) {
  const int S = plan.A.max_steps;
  if (plan.order && d_work != nullptr && count >= kOrderMin) {
    // d_work = int32[count + kOrderBuckets]: the launch order, then the bucket counters
    BGR_CUDA(launch_order_by_speed(d_params, count, d_work, d_work + count, stream));
    count_launch(3);
    return launch_range(plan, id_offset, count, d_params, d_init, d_metrics, d_status, d_traj, S, d_work, nullptr, stream);
  }
  if (split_steps <= 0 || split_steps >= S || d_work == nullptr)
    return launch_range(plan, id_offset, count, d_params, d_init, d_metrics, d_status, d_traj, S, nullptr, nullptr, stream,
                        d_queue);
  int rc = launch_range(plan, id_offset, count, d_params, d_init, d_metrics, d_status, d_traj, split_steps, nullptr,
                        nullptr, stream);
  if (rc != BGR_OK) return rc;
  int32_t* d_n = d_work + count;
  BGR_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int32_t), stream));
  BGR_CUDA(launch_compact_survivors(d_status, count, d_work, d_n, stream));
  count_launch();
  return launch_range(plan, id_offset, count, d_params, d_init, d_metrics, d_status, d_traj, S, d_work, d_n, stream);
}

int check_rollout_buffers(int64_t n_episodes, const void* d_params, const void* d_params64, const void* d_init,
                          const void* d_metrics, const void* d_status) {
  if (n_episodes < 0 || n_episodes > (1LL << 31) * kRolloutThreads) {
    set_error("bgr_rollout: n_episodes %lld out of range", static_cast<long long>(n_episodes));
    return BGR_ERR_BAD_ARG;
  }
  if (n_episodes == 0) return BGR_OK;
  if ((d_params == nullptr && d_params64 == nullptr) || d_init == nullptr || d_metrics == nullptr ||
      d_status == nullptr) {
    set_error("bgr_rollout: d_params / d_init_state / d_metrics / d_status is NULL");
    return BGR_ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(d_params) & 15) || (reinterpret_cast<uintptr_t>(d_metrics) & 15) ||
      (reinterpret_cast<uintptr_t>(d_init) & 7)) {
    set_error("bgr_rollout: d_params and d_metrics must be 16-byte aligned, d_init_state 8-byte aligned");
    return BGR_ERR_BAD_ARG;
  }
  return BGR_OK;
}

int empty_aggregate(double* d_aggregate, cudaStream_t stream) {
  if (d_aggregate != nullptr) {
    const double zero[BGR_N_AGG] = {0, 0, 0, 0, 0, 1e300, 0, 0};
    BGR_CUDA(cudaMemcpyAsync(d_aggregate, zero, sizeof(zero), cudaMemcpyHostToDevice, stream));
    BGR_CUDA(cudaStreamSynchronize(stream));
  }
  return BGR_OK;
}

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

int rollout_impl(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes, const float* d_params,
                 const double* d_init, float* d_metrics, int32_t* d_status, double* d_traj, double* d_aggregate,
                 const StepPlumbing& sp, cudaStream_t stream) {
  int rc = validate_cfg(track, cfg, "bgr_rollout");
  if (rc != BGR_OK) return rc;
  rc = check_rollout_buffers(n_episodes, d_params, sp.params64, d_init, d_metrics, d_status);
  if (rc != BGR_OK) return rc;
  DeviceGuard guard(track->device);
  if (n_episodes == 0) return empty_aggregate(d_aggregate, stream);
  Plan plan;
  rc = make_plan(track, cfg, sp, d_traj != nullptr, stream, &plan);
  if (rc != BGR_OK) return rc;
  const int grid = static_cast<int>((n_episodes + kRolloutThreads - 1) / kRolloutThreads);
  const int split = sp.force_full ? 0 : static_cast<int>(cfg->split_steps);
  double* d_partials = nullptr;
  int32_t* d_work = nullptr;
  if (d_aggregate != nullptr)
    BGR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_partials), static_cast<size_t>(grid) * BGR_N_AGG * sizeof(double),
                             stream));
  if (split > 0 && split < cfg->max_steps)
    BGR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_work), static_cast<size_t>(n_episodes + 1) * sizeof(int32_t),
                             stream));
  else if (plan.order && n_episodes >= kOrderMin)
    BGR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_work),
                             static_cast<size_t>(n_episodes + kOrderBuckets) * sizeof(int32_t), stream));
  int32_t* d_queue = nullptr;
  if (plan.dynamic && n_episodes > plan.wave_dyn)
    BGR_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_queue), sizeof(int32_t), stream));
  rc = ensure_events(track->device);
  if (rc == BGR_OK) rc = check_cuda(cudaEventRecord(g_last.start, stream), "cudaEventRecord");
  if (rc == BGR_OK)
    rc = launch_rollout_phases(plan, split, 0, n_episodes, d_params, d_init, d_metrics, d_status, d_traj, d_work, stream,
                               d_queue);
  if (rc == BGR_OK) {
    rc = check_cuda(cudaEventRecord(g_last.stop, stream), "cudaEventRecord");
    g_last.valid = rc == BGR_OK;
  }
  if (rc == BGR_OK && d_partials != nullptr) {
    rc = check_cuda(launch_aggregate_partials(d_metrics, d_status, n_episodes, d_partials, stream), "aggregate launch");
    if (rc == BGR_OK) rc = check_cuda(launch_reduce_partials(d_partials, grid, d_aggregate, stream), "aggregate reduction launch");
    if (rc == BGR_OK) count_launch(2);
  }
  if (d_partials != nullptr) cudaFreeAsync(d_partials, stream);
  if (d_work != nullptr) cudaFreeAsync(d_work, stream);
  if (d_queue != nullptr) cudaFreeAsync(d_queue, stream);
  return rc;
}

}  // namespace

extern "C" int bgr_rollout(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                           const float* d_params, const double* d_init_state, float* d_metrics, int32_t* d_status,
                           double* d_traj, double* d_aggregate, void* stream) {
  StepPlumbing sp;
  return rollout_impl(track, cfg, n_episodes, d_params, d_init_state, d_metrics, d_status, d_traj, d_aggregate, sp,
                      static_cast<cudaStream_t>(stream));
}

namespace {

// Stream-ordered scratch that is released when the scope ends.
struct Scratch {
  cudaStream_t stream;
  std::vector<void*> ptrs;
  explicit Scratch(cudaStream_t s) : stream(s) {}
  ~Scratch() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
  }
  template <typename T>
  int get(T** out, size_t count) {
    *out = nullptr;
    if (count == 0) count = 1;
    void* p = nullptr;
    int rc = check_cuda(cudaMallocAsync(&p, count * sizeof(T), stream), "cudaMallocAsync");
    if (rc != BGR_OK) return rc;
    ptrs.push_back(p);
    *out = static_cast<T*>(p);
    return BGR_OK;
  }
};

struct DeviceScope {
  int prev = 0;
  explicit DeviceScope(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DeviceScope() { cudaSetDevice(prev); }
};

#define BGR_TRY(expr)            \
  do {                           \
    int rc_ = (expr);            \
    if (rc_ != BGR_OK) return rc_; \
  } while (0)

}  // namespace

namespace {

// The host-buffer pipeline of bgr_rollout_host / bgr_rollout_host_async, per host thread:
//   * one UPLOAD stream, two COMPUTE streams, one DOWNLOAD stream, tied together by events.  Uploads of all chunks are
//     queued back to back (nothing but PCIe paces them), chunk c's kernel waits for its own upload only, its download
//     waits for its own kernel only: no copy ever sits between two kernels of the same stream.
//   * chunks are sized in WAVES.  Every CTA of a rollout runs the full episode length, so a chunk kernel lasts one CTA
//     lifetime whatever its size: a chunk that does not fill the machine wastes the rest of it for that long.  The
//     first wave is split 1/4 + 3/4 over the two compute streams (the first kernel starts after a quarter of the first
//     wave's upload), then whole waves follow; kernels of consecutive chunks overlap at their edges because they
//     alternate between the two compute streams.
//   * device buffers live in a small ring of SLOTS that persist across calls (grown on demand, never shrunk): no
//     stream-ordered allocation per call, whose cross-stream reuse would serialise consecutive asynchronous calls, and
//     up to kSlots calls in flight per thread (bgr_rollout_host_async).
constexpr int kSlots = 3;
struct Slot {
  size_t cap = 0, cap_partials = 0, cap_work = 0;   // capacities in episodes / rows / ints
  float* d_params = nullptr;
  double* d_init = nullptr;
  float* d_metrics = nullptr;
  int32_t* d_status = nullptr;
  double* d_partials = nullptr;
  double* d_agg = nullptr;
  int32_t* d_work = nullptr;
  int32_t* d_queue = nullptr;   // [kMaxQueueChunks] row counters of dynamically scheduled chunk launches
  cudaEvent_t done = nullptr;
  bool used = false;
};
constexpr int kMaxQueueChunks = 256;
struct Pipeline {
  int device = -1;
  cudaStream_t up = nullptr, down = nullptr, comp[2] = {nullptr, nullptr};
  cudaEvent_t ready = nullptr;
  std::vector<cudaEvent_t> ev;   // [2 c] = upload of chunk c done, [2 c + 1] = kernel of chunk c done (re-recorded per call)
  Slot slot[kSlots];
  unsigned next = 0;
  // No destructor on purpose: a thread_local destructor runs at thread / process exit, possibly after the CUDA runtime
  // has begun to unload, where any CUDA call is undefined behaviour (observed: SIGSEGV at interpreter shutdown when
  // worker threads owned a pipeline).  The streams and buffers of a thread's pipeline live until the process ends (the
  // driver reclaims them); release() is for switching devices while the context is alive.
  void release() {
    if (device < 0) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    for (cudaStream_t st : {up, down, comp[0], comp[1]})
      if (st != nullptr) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
      }
    if (ready != nullptr) cudaEventDestroy(ready);
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear();
    for (Slot& sl : slot) {
      cudaFree(sl.d_params); cudaFree(sl.d_init); cudaFree(sl.d_metrics); cudaFree(sl.d_status);
      cudaFree(sl.d_partials); cudaFree(sl.d_agg); cudaFree(sl.d_work); cudaFree(sl.d_queue);
      if (sl.done != nullptr) cudaEventDestroy(sl.done);
      sl = Slot();
    }
    up = down = comp[0] = comp[1] = nullptr;
    ready = nullptr;
    cudaSetDevice(prev);
    device = -1;
  }
};
thread_local Pipeline g_pipe;

int ensure_pipeline(int device) {
  if (g_pipe.device == device) return BGR_OK;
  g_pipe.release();
  BGR_CUDA(cudaStreamCreateWithFlags(&g_pipe.up, cudaStreamNonBlocking));
  BGR_CUDA(cudaStreamCreateWithFlags(&g_pipe.down, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) BGR_CUDA(cudaStreamCreateWithFlags(&g_pipe.comp[i], cudaStreamNonBlocking));
  BGR_CUDA(cudaEventCreateWithFlags(&g_pipe.ready, cudaEventDisableTiming));
  for (Slot& sl : g_pipe.slot) BGR_CUDA(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  g_pipe.device = device;
  return BGR_OK;
}

int ensure_chunk_events(size_t n_chunks) {
  while (g_pipe.ev.size() < 2 * n_chunks) {
    cudaEvent_t e;
    BGR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g_pipe.ev.push_back(e);
  }
  return BGR_OK;
}

template <typename T>
int grow(T** p, size_t count) {
  cudaFree(*p);
  *p = nullptr;
  BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  return BGR_OK;
}

// the next slot of the ring, free of its previous user and large enough for this call
int acquire_slot(size_t E, size_t n_partials, size_t n_work, Slot** out) {
  Slot& sl = g_pipe.slot[g_pipe.next++ % kSlots];
  if (sl.used) BGR_CUDA(cudaEventSynchronize(sl.done));
  if (sl.cap < E) {
    const size_t cap = E + E / 8;
    BGR_TRY(grow(&sl.d_params, cap * BGR_N_PARAMS));
    BGR_TRY(grow(&sl.d_init, cap * BGR_N_STATE));
    BGR_TRY(grow(&sl.d_metrics, cap * BGR_N_METRICS));
    BGR_TRY(grow(&sl.d_status, cap * BGR_N_STATUS));
    sl.cap = cap;
  }
  if (sl.cap_partials < n_partials) {
    BGR_TRY(grow(&sl.d_partials, (n_partials + n_partials / 8 + 8) * BGR_N_AGG));
    sl.cap_partials = n_partials + n_partials / 8 + 8;
  }
  if (sl.d_agg == nullptr) BGR_TRY(grow(&sl.d_agg, BGR_N_AGG));
  if (sl.d_queue == nullptr) BGR_TRY(grow(&sl.d_queue, kMaxQueueChunks));
  if (sl.cap_work < n_work) {
    BGR_TRY(grow(&sl.d_work, n_work + n_work / 8 + 64));
    sl.cap_work = n_work + n_work / 8 + 64;
  }
  sl.used = true;
  *out = &sl;
  return BGR_OK;
}

// Chunk sizes (episodes) for n episodes on a machine that holds `wave` of them at once.  BGR_HOST_CHUNK=<episodes>
// overrides the wave size for tuning.
std::vector<long long> host_chunks(long long n_episodes, long long wave) {
  static const long long forced = [] {
    const char* e = std::getenv("BGR_HOST_CHUNK");
    return e != nullptr ? std::atoll(e) : 0LL;
  }();
  if (forced >= kRolloutThreads) wave = forced;
  wave = std::max<long long>(kRolloutThreads, wave / kRolloutThreads * kRolloutThreads);
  std::vector<long long> out;
  long long left = n_episodes;
  const long long first_wave = std::min(left, wave);
  long long head = std::max<long long>(kRolloutThreads, first_wave / 4 / kRolloutThreads * kRolloutThreads);
  head = std::min(head, first_wave);
  out.push_back(head);
  if (first_wave > head) out.push_back(first_wave - head);
  left -= first_wave;
  while (left > 0) {
    long long c = std::min(wave, left);
    if (left - c > 0 && left - c < wave / 8) c = left;      // no sliver at the end
    out.push_back(c);
    left -= c;
  }
  return out;
}

int rollout_host_impl(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                      const float* h_params, const double* h_init_state, float* h_metrics,
                      int32_t* h_status, double* h_traj, double* h_aggregate, void* stream_, bool synchronous) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BGR_TRY(validate_cfg(track, cfg, "bgr_rollout_host"));
  if (n_episodes < 0) {
    set_error("bgr_rollout_host: n_episodes < 0");
    return BGR_ERR_BAD_ARG;
  }
  if (n_episodes > 0 && (h_params == nullptr || h_init_state == nullptr || h_metrics == nullptr || h_status == nullptr)) {
    set_error("bgr_rollout_host: h_params / h_init_state / h_metrics / h_status is NULL");
    return BGR_ERR_BAD_ARG;
  }
  if (n_episodes == 0) {
    if (h_aggregate != nullptr) {
      const double zero[BGR_N_AGG] = {0, 0, 0, 0, 0, 1e300, 0, 0};
      std::memcpy(h_aggregate, zero, sizeof(zero));
    }
    return BGR_OK;
  }
  DeviceScope dev(track->device);
  StepPlumbing sp;
  Plan plan;
  BGR_TRY(make_plan(track, cfg, sp, h_traj != nullptr, stream, &plan));
  BGR_TRY(ensure_pipeline(track->device));
  BGR_TRY(ensure_events(track->device));
  const size_t E = static_cast<size_t>(n_episodes);
  const size_t S = static_cast<size_t>(cfg->max_steps);
  const bool two_phase = cfg->split_steps > 0 && static_cast<int>(cfg->split_steps) < cfg->max_steps;
  // a dynamically scheduled kernel wants MANY episodes per launch (its lanes refill from the queue): four waves per
  // chunk instead of one
  // Rollouts whose episodes end at different times (speed-ordered or dynamically scheduled plans) are packed best by ONE
  // launch -- chunk boundaries would cut the order into pieces and leave a tail per chunk -- and their copies are small
  // next to their compute: up to kSingleShotBytes of input they go up in one piece.
  constexpr size_t kSingleShotBytes = 64u << 20;
  const bool single_shot = (plan.order || plan.dynamic) &&
                           static_cast<size_t>(n_episodes) * (BGR_N_PARAMS * sizeof(float) + BGR_N_STATE * sizeof(double)) <=
                               kSingleShotBytes;
  const std::vector<long long> chunks = single_shot ? std::vector<long long>{static_cast<long long>(n_episodes)}
                                                    : host_chunks(n_episodes, plan.dynamic ? 4 * plan.wave_dyn : plan.wave);
  const int grid_total = static_cast<int>((n_episodes + kRolloutThreads - 1) / kRolloutThreads);
  BGR_TRY(ensure_chunk_events(chunks.size()));
  Slot* sl = nullptr;
  BGR_TRY(acquire_slot(E, h_aggregate != nullptr ? static_cast<size_t>(grid_total) : 0,
                       two_phase ? E + chunks.size() + 64 : plan.order ? E + chunks.size() * kOrderBuckets : 0, &sl));
  // the trajectory dump (a debugging aid, synchronous calls only) is the one stream-ordered allocation left
  Scratch sc(stream);
  double* d_traj = nullptr;
  if (h_traj != nullptr) {
    BGR_TRY(sc.get(&d_traj, E * S * BGR_N_TRAJ));
    BGR_CUDA(cudaMemsetAsync(d_traj, 0, E * S * BGR_N_TRAJ * sizeof(double), stream));
  }

  // If anything below fails half way, the internal streams may still be working on buffers the caller is about to
  // reuse or free: drain them first.
  struct Drain {
    bool armed = true;
    ~Drain() {
      if (armed)
        for (cudaStream_t st : {g_pipe.up, g_pipe.comp[0], g_pipe.comp[1], g_pipe.down}) cudaStreamSynchronize(st);
    }
  } drain;
  // everything queued on the caller's stream so far happens before the first upload
  BGR_CUDA(cudaEventRecord(g_pipe.ready, stream));
  BGR_CUDA(cudaStreamWaitEvent(g_pipe.up, g_pipe.ready, 0));
  if (d_traj != nullptr)
    for (int i = 0; i < 2; ++i) BGR_CUDA(cudaStreamWaitEvent(g_pipe.comp[i], g_pipe.ready, 0));
  long long e0 = 0;
  for (size_t c = 0; c < chunks.size(); ++c) {
    const long long cnt = chunks[c];
    const size_t n = static_cast<size_t>(cnt), o = static_cast<size_t>(e0);
    cudaEvent_t up_done = g_pipe.ev[2 * c], k_done = g_pipe.ev[2 * c + 1];
    cudaStream_t cs = g_pipe.comp[c & 1];
    BGR_CUDA(cudaMemcpyAsync(sl->d_params + o * BGR_N_PARAMS, h_params + o * BGR_N_PARAMS,
                             n * BGR_N_PARAMS * sizeof(float), cudaMemcpyHostToDevice, g_pipe.up));
    BGR_CUDA(cudaMemcpyAsync(sl->d_init + o * BGR_N_STATE, h_init_state + o * BGR_N_STATE,
                             n * BGR_N_STATE * sizeof(double), cudaMemcpyHostToDevice, g_pipe.up));
    BGR_CUDA(cudaEventRecord(up_done, g_pipe.up));
    BGR_CUDA(cudaStreamWaitEvent(cs, up_done, 0));
    const bool last = c + 1 == chunks.size();
    if (last) BGR_CUDA(cudaEventRecord(g_last.start, cs));
    BGR_TRY(launch_rollout_phases(plan, static_cast<int>(cfg->split_steps), e0, cnt, sl->d_params + o * BGR_N_PARAMS,
                                  sl->d_init + o * BGR_N_STATE, sl->d_metrics + o * BGR_N_METRICS,
                                  sl->d_status + o * BGR_N_STATUS,
                                  d_traj != nullptr ? d_traj + o * S * BGR_N_TRAJ : nullptr,
                                  two_phase ? sl->d_work + o + c : plan.order ? sl->d_work + o + c * kOrderBuckets : nullptr,
                                  cs,
                                  c < static_cast<size_t>(kMaxQueueChunks) ? sl->d_queue + c : nullptr));
    if (last) {
      BGR_CUDA(cudaEventRecord(g_last.stop, cs));
      g_last.valid = true;
    }
    BGR_CUDA(cudaEventRecord(k_done, cs));
    BGR_CUDA(cudaStreamWaitEvent(g_pipe.down, k_done, 0));
    BGR_CUDA(cudaMemcpyAsync(h_metrics + o * BGR_N_METRICS, sl->d_metrics + o * BGR_N_METRICS,
                             n * BGR_N_METRICS * sizeof(float), cudaMemcpyDeviceToHost, g_pipe.down));
    BGR_CUDA(cudaMemcpyAsync(h_status + o * BGR_N_STATUS, sl->d_status + o * BGR_N_STATUS,
                             n * BGR_N_STATUS * sizeof(int32_t), cudaMemcpyDeviceToHost, g_pipe.down));
    if (h_traj != nullptr)
      BGR_CUDA(cudaMemcpyAsync(h_traj + o * S * BGR_N_TRAJ, d_traj + o * S * BGR_N_TRAJ,
                               n * S * BGR_N_TRAJ * sizeof(double), cudaMemcpyDeviceToHost, g_pipe.down));
    e0 += cnt;
  }
  // the download stream has now waited for every chunk's kernel: the aggregate is computed from the complete rows
  if (h_aggregate != nullptr) {
    BGR_CUDA(launch_aggregate_partials(sl->d_metrics, sl->d_status, n_episodes, sl->d_partials, g_pipe.down));
    BGR_CUDA(launch_reduce_partials(sl->d_partials, grid_total, sl->d_agg, g_pipe.down));
    count_launch(2);
    BGR_CUDA(cudaMemcpyAsync(h_aggregate, sl->d_agg, BGR_N_AGG * sizeof(double), cudaMemcpyDeviceToHost, g_pipe.down));
  }
  BGR_CUDA(cudaEventRecord(sl->done, g_pipe.down));
  BGR_CUDA(cudaStreamWaitEvent(stream, sl->done, 0));
  // the caller's stream has waited for the whole call: synchronising it completes it (and the trajectory scratch is
  // freed in stream order behind that wait)
  if (synchronous) BGR_CUDA(cudaStreamSynchronize(stream));
  drain.armed = false;
  return BGR_OK;
}

}  // namespace

extern "C" int bgr_rollout_host(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                                const float* h_params, const double* h_init_state, float* h_metrics,
                                int32_t* h_status, double* h_traj, double* h_aggregate, void* stream) {
  return rollout_host_impl(track, cfg, n_episodes, h_params, h_init_state, h_metrics, h_status, h_traj, h_aggregate,
                           stream, true);
}

extern "C" int bgr_rollout_host_async(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg, int64_t n_episodes,
                                      const float* h_params, const double* h_init_state, float* h_metrics,
                                      int32_t* h_status, double* h_aggregate, void* stream) {
  return rollout_host_impl(track, cfg, n_episodes, h_params, h_init_state, h_metrics, h_status, nullptr, h_aggregate,
                           stream, false);
}

extern "C" int bgr_step_host(const bgr_track_t* track, const bgr_rollout_cfg_t* cfg_in, int32_t phases, int64_t n,
                             const double* h_params, double* h_state, int32_t* h_target_ind, double* h_ctrl,
                             const double* h_cmd, double* h_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BGR_TRY(validate_cfg(track, cfg_in, "bgr_step_host"));
  if (n < 1 || h_params == nullptr || h_state == nullptr) {
    set_error("bgr_step_host: need n >= 1, h_params and h_state");
    return BGR_ERR_BAD_ARG;
  }
  const int32_t all = BGR_PHASE_STEER | BGR_PHASE_ACCEL | BGR_PHASE_MODEL | BGR_PHASE_ERRORS | BGR_PHASE_KAPPA_GIVEN;
  if ((phases & ~all) != 0 || (phases & all) == 0) {
    set_error("bgr_step_host: phases 0x%x is not a combination of BGR_PHASE_*", phases);
    return BGR_ERR_BAD_ARG;
  }
  const bool need_cmd = !(phases & BGR_PHASE_STEER) || !(phases & BGR_PHASE_ACCEL) || (phases & BGR_PHASE_KAPPA_GIVEN);
  if (need_cmd && h_cmd == nullptr) {
    set_error("bgr_step_host: h_cmd is required when a law phase is off or the curvature is given");
    return BGR_ERR_BAD_ARG;
  }
  bgr_rollout_cfg_t cfg = *cfg_in;
  cfg.max_steps = 1;
  cfg.laps_target = 0;
  DeviceScope dev(track->device);
  Scratch sc(stream);
  const size_t E = static_cast<size_t>(n);
  double* d_params;
  double *d_state, *d_state_out, *d_ctrl_in, *d_ctrl_out, *d_cmd = nullptr, *d_out;
  float* d_metrics;
  int32_t *d_status, *d_ti_in, *d_ti_out;
  BGR_TRY(sc.get(&d_params, E * BGR_N_PARAMS));
  BGR_TRY(sc.get(&d_state, E * BGR_N_STATE));
  BGR_TRY(sc.get(&d_state_out, E * BGR_N_STATE));
  BGR_TRY(sc.get(&d_ctrl_in, E * BGR_N_CTRL));
  BGR_TRY(sc.get(&d_ctrl_out, E * BGR_N_CTRL));
  BGR_TRY(sc.get(&d_out, E * BGR_N_OUT));
  BGR_TRY(sc.get(&d_metrics, E * BGR_N_METRICS));
  BGR_TRY(sc.get(&d_status, E * BGR_N_STATUS));
  BGR_TRY(sc.get(&d_ti_in, E));
  BGR_TRY(sc.get(&d_ti_out, E));
  BGR_CUDA(cudaMemcpyAsync(d_params, h_params, E * BGR_N_PARAMS * sizeof(double), cudaMemcpyHostToDevice, stream));
  BGR_CUDA(cudaMemcpyAsync(d_state, h_state, E * BGR_N_STATE * sizeof(double), cudaMemcpyHostToDevice, stream));
  if (h_ctrl != nullptr)
    BGR_CUDA(cudaMemcpyAsync(d_ctrl_in, h_ctrl, E * BGR_N_CTRL * sizeof(double), cudaMemcpyHostToDevice, stream));
  else
    BGR_CUDA(cudaMemsetAsync(d_ctrl_in, 0, E * BGR_N_CTRL * sizeof(double), stream));
  if (h_target_ind != nullptr)
    BGR_CUDA(cudaMemcpyAsync(d_ti_in, h_target_ind, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  else
    BGR_CUDA(cudaMemsetAsync(d_ti_in, 0, E * sizeof(int32_t), stream));
  if (h_cmd != nullptr) {
    BGR_TRY(sc.get(&d_cmd, E * BGR_N_CMD));
    BGR_CUDA(cudaMemcpyAsync(d_cmd, h_cmd, E * BGR_N_CMD * sizeof(double), cudaMemcpyHostToDevice, stream));
  }
  StepPlumbing sp;
  sp.phases = phases;
  sp.ctrl_in = d_ctrl_in;
  sp.ctrl_out = d_ctrl_out;
  sp.ti_in = d_ti_in;
  sp.ti_out = d_ti_out;
  sp.state_out = d_state_out;
  sp.cmd = d_cmd;
  sp.step_out = d_out;
  sp.params64 = d_params;
  sp.force_full = true;
  BGR_TRY(rollout_impl(track, &cfg, n, nullptr, d_state, d_metrics, d_status, nullptr, nullptr, sp, stream));
  if (phases & BGR_PHASE_MODEL)
    BGR_CUDA(cudaMemcpyAsync(h_state, d_state_out, E * BGR_N_STATE * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (h_ctrl != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_ctrl, d_ctrl_out, E * BGR_N_CTRL * sizeof(double), cudaMemcpyDeviceToHost, stream));
  if (h_target_ind != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_target_ind, d_ti_out, E * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  if (h_out != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_out, d_out, E * BGR_N_OUT * sizeof(double), cudaMemcpyDeviceToHost, stream));
  BGR_CUDA(cudaStreamSynchronize(stream));
  return BGR_OK;
}

// ================================================================================================
// MPC candidate evaluation
// ================================================================================================
static int validate_mpc(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes, int32_t K,
                        const char* who) {
  if (track == nullptr || cfg == nullptr) {
    set_error("%s: track / cfg is NULL", who);
    return BGR_ERR_BAD_ARG;
  }
  if (cfg->horizon < 1 || cfg->horizon > 64 || !(cfg->dt > 0.0) || !(cfg->wheelbase > 0.0) || cfg->precision < 0 ||
      cfg->precision > 1) {
    set_error("%s: need 1 <= horizon <= 64, dt > 0, wheelbase > 0 (got %d, %g, %g)", who, cfg->horizon, cfg->dt,
              cfg->wheelbase);
    return BGR_ERR_BAD_ARG;
  }
  for (int i = 0; i < 4; ++i) {
    if (cfg->q[i] < 0.0 || (i < 2 && cfg->r[i] < 0.0)) {
      set_error("%s: Q and R weights must be >= 0 (costs are compared through their bit patterns)", who);
      return BGR_ERR_BAD_ARG;
    }
  }
  if (n_episodes < 0 || K < 1) {
    set_error("%s: need n_episodes >= 0 and n_candidates >= 1", who);
    return BGR_ERR_BAD_ARG;
  }
  return BGR_OK;
}

// Device-resident candidate bank (bgr_mpc_bank_t): the bank itself plus, per cost configuration it has been scored
// under, the five prefix-sum statistics per candidate that the factored evaluator needs.
struct bgr_mpc_bank {
  int device = 0;
  int32_t K = 0, H = 0;
  float* d_bank = nullptr;
  struct Stats {
    double dt, r0, r1, q3;
    void* d_cand;
  };
  std::mutex mu;
  std::vector<Stats> stats;
  // device arena of the HOST entry point (states, outputs, evaluator scratch of one call), grown on demand: a call
  // through the handle makes no allocation at all; host calls on one handle are serialised by `call_mu`
  std::mutex call_mu;
  // Two arenas: a synchronous call uses slot 0; asynchronous calls (bgr_mpc_eval_bank_host_async) alternate, so that the
  // upload of call i + 1 can run under the kernels of call i.  A slot remembers the stream of its last call and waits
  // for it when it comes back on another stream.
  struct Arena {
    unsigned char* mem = nullptr;
    size_t bytes = 0;
    cudaStream_t last = nullptr;
    bool used = false;
  };
  Arena arenas[2];
  unsigned next_async = 0;
};

namespace {

// The evaluation proper, shared by the pointer and the handle entry points.  `bank_handle` (may be NULL) supplies
// cached candidate statistics; without it they are recomputed from d_bank inside the call.
int mpc_eval_core(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes, const double* d_state,
                  const int32_t* d_ref_start, const int32_t* d_ref_len, const float* d_bank, int32_t n_candidates,
                  bgr_mpc_bank* bank_handle, float* d_best_u, float* d_best_cost, int32_t* d_best_idx, float* d_costs,
                  cudaStream_t stream, const char* who, unsigned char* evaluator_scratch = nullptr) {
  BGR_TRY(validate_mpc(track, cfg, n_episodes, n_candidates, who));
  if (n_episodes == 0) return BGR_OK;
  if (d_state == nullptr || d_ref_start == nullptr || d_bank == nullptr) {
    set_error("%s: state / ref_start / bank is NULL", who);
    return BGR_ERR_BAD_ARG;
  }
  if (reinterpret_cast<uintptr_t>(d_bank) & 7) {
    set_error("%s: the bank must be 8-byte aligned", who);
    return BGR_ERR_BAD_ARG;
  }
  DeviceScope dev(track->device);
  DeviceInfo di;
  BGR_TRY(device_info(track->device, &di));
  Scratch sc(stream);
  const bool explicit_rollout = cfg->precision == BGR_PRECISION_FP64 || (cfg->flags & BGR_MPC_EXPLICIT_ROLLOUT);
  MpcArgs A;
  std::memset(&A, 0, sizeof(A));
  A.trk = track->view();
  A.state = d_state;
  A.ref_start = d_ref_start;
  A.ref_len = d_ref_len;
  A.bank = d_bank;
  A.costs = d_costs;
  A.n_episodes = n_episodes;
  A.K = n_candidates;
  A.H = cfg->horizon;
  A.dt = cfg->dt;
  A.wheelbase = cfg->wheelbase;
  A.target_speed = cfg->target_speed;
  A.q0 = cfg->q[0];
  A.q1 = cfg->q[1];
  A.q2 = cfg->q[2];
  A.q3 = cfg->q[3];
  A.r0 = cfg->r[0];
  A.r1 = cfg->r[1];
  // cached candidate statistics of a bank handle: computed once per (dt, r0, r1, q3), kept until the handle dies
  void* d_cand_cached = nullptr;
  if (bank_handle != nullptr && !explicit_rollout) {
    std::lock_guard<std::mutex> lock(bank_handle->mu);
    for (const auto& st : bank_handle->stats)
      if (st.dt == A.dt && st.r0 == A.r0 && st.r1 == A.r1 && st.q3 == A.q3) d_cand_cached = st.d_cand;
    if (d_cand_cached == nullptr) {
      void* p = nullptr;
      BGR_CUDA(cudaMalloc(&p, mpc_cand_bytes(n_candidates)));
      cudaError_t err = launch_mpc_cand_stats(A, p, stream);
      if (err == cudaSuccess) err = cudaStreamSynchronize(stream);   // (once per configuration)
      if (err != cudaSuccess) {
        cudaFree(p);
        return check_cuda(err, "candidate statistics");
      }
      count_launch();
      bank_handle->stats.push_back({A.dt, A.r0, A.r1, A.q3, p});
      d_cand_cached = p;
    }
  }
  // one scratch block: [packed argmin accumulators | candidate statistics | episode coefficients]
  const size_t b_packed = (static_cast<size_t>(n_episodes) * sizeof(unsigned long long) + 255) & ~size_t(255);
  const size_t b_cand = (explicit_rollout || d_cand_cached != nullptr) ? 0 : (mpc_cand_bytes(n_candidates) + 255) & ~size_t(255);
  const size_t b_epi = explicit_rollout ? 0 : mpc_epi_bytes(n_episodes);
  unsigned char* d_scratch = evaluator_scratch;      // (sized by mpc_scratch_bytes; only with cached statistics)
  if (d_scratch == nullptr || b_cand != 0) BGR_TRY(sc.get(&d_scratch, b_packed + b_cand + b_epi));
  unsigned long long* d_packed = reinterpret_cast<unsigned long long*>(d_scratch);
  if (explicit_rollout)
    BGR_CUDA(cudaMemsetAsync(d_packed, 0xff, static_cast<size_t>(n_episodes) * sizeof(unsigned long long), stream));
  A.packed = d_packed;
  BGR_TRY(ensure_events(track->device));
  BGR_CUDA(cudaEventRecord(g_last.start, stream));
  if (explicit_rollout) {
    // audit: explicit stage-by-stage rollout, reference operation order (fp64), or the fp32 explicit rollout
    BGR_CUDA(launch_mpc(A, cfg->precision, stream, di.sm_count));
    count_launch();
  } else {
    BGR_CUDA(launch_mpc_factored(A, d_cand_cached != nullptr ? d_cand_cached : d_scratch + b_packed,
                                 d_scratch + b_packed + b_cand, stream, di.sm_count, d_cand_cached != nullptr));
    count_launch(2);
  }
  BGR_CUDA(cudaEventRecord(g_last.stop, stream));
  g_last.valid = true;
  BGR_CUDA(launch_mpc_finalize(d_packed, d_bank, cfg->horizon, n_episodes, static_cast<float>(cfg->max_accel),
                               static_cast<float>(cfg->max_decel), static_cast<float>(cfg->max_steer), d_best_u,
                               d_best_cost, d_best_idx, stream));
  count_launch();
  return BGR_OK;
}

// host buffers around mpc_eval_core; `h_bank` is uploaded per call unless a handle supplies the resident bank
int mpc_eval_host_core(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes, const double* h_state,
                       const int32_t* h_ref_start, const int32_t* h_ref_len, const float* h_bank, int32_t n_candidates,
                       bgr_mpc_bank* bank_handle, float* h_best_u, float* h_best_cost, int32_t* h_best_idx,
                       float* h_costs, cudaStream_t stream, const char* who, bool synchronous = true) {
  BGR_TRY(validate_mpc(track, cfg, n_episodes, n_candidates, who));
  if (n_episodes == 0) return BGR_OK;
  if (h_state == nullptr || h_ref_start == nullptr || (h_bank == nullptr && bank_handle == nullptr)) {
    set_error("%s: h_state / h_ref_start / bank is NULL", who);
    return BGR_ERR_BAD_ARG;
  }
  for (int64_t e = 0; e < n_episodes; ++e) {
    if (h_ref_start[e] < 0 || h_ref_start[e] >= track->n || (h_ref_len != nullptr && h_ref_len[e] < 1)) {
      set_error("%s: episode %lld has ref_start %d (track has %d waypoints)%s", who, static_cast<long long>(e),
                h_ref_start[e], track->n, h_ref_len != nullptr && h_ref_len[e] < 1 ? " / ref_len < 1" : "");
      return BGR_ERR_BAD_ARG;
    }
  }
  DeviceScope dev(track->device);
  Scratch sc(stream);
  const size_t E = static_cast<size_t>(n_episodes), K = static_cast<size_t>(n_candidates);
  const size_t bank_count = K * static_cast<size_t>(cfg->horizon) * 2;
  double* d_state;
  int32_t *d_start, *d_len = nullptr, *d_idx;
  float *d_bank = bank_handle != nullptr ? bank_handle->d_bank : nullptr, *d_u, *d_cost, *d_costs = nullptr;
  unsigned char* evaluator_scratch = nullptr;
  std::unique_lock<std::mutex> call_lock;
  if (bank_handle != nullptr) {
    // everything but the optional cost matrix comes out of the handle's arena: [state | start | len | u | cost | idx |
    // evaluator scratch], 256-byte aligned pieces
    call_lock = std::unique_lock<std::mutex>(bank_handle->call_mu);
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    const size_t o_state = 0, o_start = o_state + up(E * 4 * sizeof(double)), o_len = o_start + up(E * sizeof(int32_t)),
                 o_u = o_len + up(E * sizeof(int32_t)), o_cost = o_u + up(E * 2 * sizeof(float)),
                 o_idx = o_cost + up(E * sizeof(float)), o_eval = o_idx + up(E * sizeof(int32_t)),
                 total = o_eval + up(E * sizeof(unsigned long long)) + up(mpc_epi_bytes(n_episodes)) + 256;
    bgr_mpc_bank::Arena& ar = bank_handle->arenas[synchronous ? 0 : 1 & bank_handle->next_async++];
    if (ar.used && ar.last != stream) BGR_CUDA(cudaStreamSynchronize(ar.last));   // the slot's previous call, elsewhere
    if (ar.bytes < total) {
      if (ar.used) BGR_CUDA(cudaStreamSynchronize(ar.last));
      cudaFree(ar.mem);
      ar.mem = nullptr;
      ar.bytes = 0;
      BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(&ar.mem), total + total / 4));
      ar.bytes = total + total / 4;
    }
    ar.last = stream;
    ar.used = true;
    unsigned char* a = ar.mem;
    d_state = reinterpret_cast<double*>(a + o_state);
    d_start = reinterpret_cast<int32_t*>(a + o_start);
    if (h_ref_len != nullptr) d_len = reinterpret_cast<int32_t*>(a + o_len);
    d_u = reinterpret_cast<float*>(a + o_u);
    d_cost = reinterpret_cast<float*>(a + o_cost);
    d_idx = reinterpret_cast<int32_t*>(a + o_idx);
    evaluator_scratch = a + o_eval;
  } else {
    BGR_TRY(sc.get(&d_state, E * 4));
    BGR_TRY(sc.get(&d_start, E));
    BGR_TRY(sc.get(&d_bank, bank_count));
    BGR_TRY(sc.get(&d_u, E * 2));
    BGR_TRY(sc.get(&d_cost, E));
    BGR_TRY(sc.get(&d_idx, E));
  }
  BGR_CUDA(cudaMemcpyAsync(d_state, h_state, E * 4 * sizeof(double), cudaMemcpyHostToDevice, stream));
  BGR_CUDA(cudaMemcpyAsync(d_start, h_ref_start, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  if (bank_handle == nullptr)
    BGR_CUDA(cudaMemcpyAsync(d_bank, h_bank, bank_count * sizeof(float), cudaMemcpyHostToDevice, stream));
  if (h_ref_len != nullptr) {
    if (d_len == nullptr) BGR_TRY(sc.get(&d_len, E));
    BGR_CUDA(cudaMemcpyAsync(d_len, h_ref_len, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  }
  if (h_costs != nullptr) BGR_TRY(sc.get(&d_costs, E * K));
  BGR_TRY(mpc_eval_core(track, cfg, n_episodes, d_state, d_start, d_len, d_bank, n_candidates, bank_handle, d_u, d_cost,
                        d_idx, d_costs, stream, who, evaluator_scratch));
  if (h_best_u != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_best_u, d_u, E * 2 * sizeof(float), cudaMemcpyDeviceToHost, stream));
  if (h_best_cost != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_best_cost, d_cost, E * sizeof(float), cudaMemcpyDeviceToHost, stream));
  if (h_best_idx != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_best_idx, d_idx, E * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  if (h_costs != nullptr)
    BGR_CUDA(cudaMemcpyAsync(h_costs, d_costs, E * K * sizeof(float), cudaMemcpyDeviceToHost, stream));
  if (synchronous) BGR_CUDA(cudaStreamSynchronize(stream));
  return BGR_OK;
}

int check_bank_handle(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, const bgr_mpc_bank* bank, const char* who) {
  if (bank == nullptr || track == nullptr || cfg == nullptr) {
    set_error("%s: track / cfg / bank is NULL", who);
    return BGR_ERR_BAD_ARG;
  }
  if (bank->device != track->device || bank->H != cfg->horizon) {
    set_error("%s: the bank lives on device %d with horizon %d; the call needs device %d, horizon %d", who, bank->device,
              bank->H, track->device, cfg->horizon);
    return BGR_ERR_BAD_ARG;
  }
  return BGR_OK;
}

}  // namespace

extern "C" int bgr_mpc_eval(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                            const double* d_state, const int32_t* d_ref_start, const int32_t* d_ref_len,
                            const float* d_bank, int32_t n_candidates, float* d_best_u, float* d_best_cost,
                            int32_t* d_best_idx, float* d_costs, void* stream) {
  return mpc_eval_core(track, cfg, n_episodes, d_state, d_ref_start, d_ref_len, d_bank, n_candidates, nullptr, d_best_u,
                       d_best_cost, d_best_idx, d_costs, static_cast<cudaStream_t>(stream), "bgr_mpc_eval");
}

extern "C" int bgr_mpc_eval_host(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                                 const double* h_state, const int32_t* h_ref_start, const int32_t* h_ref_len,
                                 const float* h_bank, int32_t n_candidates, float* h_best_u, float* h_best_cost,
                                 int32_t* h_best_idx, float* h_costs, void* stream) {
  return mpc_eval_host_core(track, cfg, n_episodes, h_state, h_ref_start, h_ref_len, h_bank, n_candidates, nullptr,
                            h_best_u, h_best_cost, h_best_idx, h_costs, static_cast<cudaStream_t>(stream),
                            "bgr_mpc_eval_host");
}

extern "C" int bgr_mpc_bank_create(int device, const float* h_bank, int32_t n_candidates, int32_t horizon,
                                   bgr_mpc_bank_t** out_bank) {
  if (out_bank == nullptr) {
    set_error("bgr_mpc_bank_create: out_bank is NULL");
    return BGR_ERR_BAD_ARG;
  }
  *out_bank = nullptr;
  if (h_bank == nullptr || n_candidates < 1 || horizon < 1 || horizon > 64) {
    set_error("bgr_mpc_bank_create: need h_bank, n_candidates >= 1 and 1 <= horizon <= 64");
    return BGR_ERR_BAD_ARG;
  }
  int n_dev = 0;
  BGR_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    set_error("bgr_mpc_bank_create: device %d out of range (%d CUDA devices)", device, n_dev);
    return BGR_ERR_BAD_ARG;
  }
  bgr_mpc_bank* b = new (std::nothrow) bgr_mpc_bank();
  if (b == nullptr) {
    set_error("bgr_mpc_bank_create: out of host memory");
    return BGR_ERR_ALLOC;
  }
  b->device = device;
  b->K = n_candidates;
  b->H = horizon;
  DeviceScope dev(device);
  const size_t bytes = static_cast<size_t>(n_candidates) * horizon * 2 * sizeof(float);
  cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&b->d_bank), bytes);
  if (err == cudaSuccess) err = cudaMemcpy(b->d_bank, h_bank, bytes, cudaMemcpyHostToDevice);
  if (err != cudaSuccess) {
    cudaFree(b->d_bank);
    delete b;
    return check_cuda(err, "bgr_mpc_bank_create");
  }
  *out_bank = b;
  return BGR_OK;
}

extern "C" void bgr_mpc_bank_destroy(bgr_mpc_bank_t* bank) {
  if (bank == nullptr) return;
  DeviceScope dev(bank->device);
  cudaFree(bank->d_bank);
  for (auto& ar : bank->arenas) {
    if (ar.used) cudaStreamSynchronize(ar.last);
    cudaFree(ar.mem);
  }
  for (auto& st : bank->stats) cudaFree(st.d_cand);
  delete bank;
}

extern "C" int bgr_mpc_eval_bank(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                                 const double* d_state, const int32_t* d_ref_start, const int32_t* d_ref_len,
                                 bgr_mpc_bank_t* bank, float* d_best_u, float* d_best_cost, int32_t* d_best_idx,
                                 float* d_costs, void* stream) {
  BGR_TRY(check_bank_handle(track, cfg, bank, "bgr_mpc_eval_bank"));
  return mpc_eval_core(track, cfg, n_episodes, d_state, d_ref_start, d_ref_len, bank->d_bank, bank->K, bank, d_best_u,
                       d_best_cost, d_best_idx, d_costs, static_cast<cudaStream_t>(stream), "bgr_mpc_eval_bank");
}

extern "C" int bgr_mpc_eval_bank_host(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                                      const double* h_state, const int32_t* h_ref_start, const int32_t* h_ref_len,
                                      bgr_mpc_bank_t* bank, float* h_best_u, float* h_best_cost, int32_t* h_best_idx,
                                      float* h_costs, void* stream) {
  BGR_TRY(check_bank_handle(track, cfg, bank, "bgr_mpc_eval_bank_host"));
  return mpc_eval_host_core(track, cfg, n_episodes, h_state, h_ref_start, h_ref_len, nullptr, bank->K, bank, h_best_u,
                            h_best_cost, h_best_idx, h_costs, static_cast<cudaStream_t>(stream), "bgr_mpc_eval_bank_host");
}

extern "C" int bgr_mpc_eval_bank_host_async(const bgr_track_t* track, const bgr_mpc_cfg_t* cfg, int64_t n_episodes,
                                            const double* h_state, const int32_t* h_ref_start, const int32_t* h_ref_len,
                                            bgr_mpc_bank_t* bank, float* h_best_u, float* h_best_cost,
                                            int32_t* h_best_idx, void* stream) {
  BGR_TRY(check_bank_handle(track, cfg, bank, "bgr_mpc_eval_bank_host_async"));
  return mpc_eval_host_core(track, cfg, n_episodes, h_state, h_ref_start, h_ref_len, nullptr, bank->K, bank, h_best_u,
                            h_best_cost, h_best_idx, nullptr, static_cast<cudaStream_t>(stream),
                            "bgr_mpc_eval_bank_host_async", false);
}

// ================================================================================================
// batched system identification
// ================================================================================================
extern "C" int bgr_identify_dynamics(int64_t n_episodes, int32_t max_rows, const double* d_traj, const int32_t* d_rows,
                                     double* d_AB, int32_t* d_ok, void* stream_) {
  if (n_episodes < 0 || max_rows < 2 || n_episodes > 0x7fffffffLL) {
    set_error("bgr_identify_dynamics: need 0 <= n_episodes < 2^31 and max_rows >= 2");
    return BGR_ERR_BAD_ARG;
  }
  if (n_episodes == 0) return BGR_OK;
  if (d_traj == nullptr || d_rows == nullptr || d_AB == nullptr || d_ok == nullptr) {
    set_error("bgr_identify_dynamics: NULL buffer");
    return BGR_ERR_BAD_ARG;
  }
  BGR_CUDA(launch_sysid(d_traj, d_rows, max_rows, n_episodes, d_AB, d_ok, static_cast<cudaStream_t>(stream_)));
  count_launch();
  return BGR_OK;
}

extern "C" int bgr_identify_dynamics_host(int64_t n_episodes, int32_t max_rows, const double* h_traj,
                                          const int32_t* h_rows, double* h_AB, int32_t* h_ok, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_episodes < 0 || max_rows < 2) {
    set_error("bgr_identify_dynamics_host: need n_episodes >= 0 and max_rows >= 2");
    return BGR_ERR_BAD_ARG;
  }
  if (n_episodes == 0) return BGR_OK;
  if (h_traj == nullptr || h_rows == nullptr || h_AB == nullptr || h_ok == nullptr) {
    set_error("bgr_identify_dynamics_host: NULL buffer");
    return BGR_ERR_BAD_ARG;
  }
  Scratch sc(stream);
  const size_t E = static_cast<size_t>(n_episodes);
  const size_t n_traj = E * static_cast<size_t>(max_rows) * BGR_N_TRAJ;
  double *d_traj, *d_AB;
  int32_t *d_rows, *d_ok;
  BGR_TRY(sc.get(&d_traj, n_traj));
  BGR_TRY(sc.get(&d_rows, E));
  BGR_TRY(sc.get(&d_AB, E * 6 * BGR_SYSID_COLS));
  BGR_TRY(sc.get(&d_ok, E));
  BGR_CUDA(cudaMemcpyAsync(d_traj, h_traj, n_traj * sizeof(double), cudaMemcpyHostToDevice, stream));
  BGR_CUDA(cudaMemcpyAsync(d_rows, h_rows, E * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
  BGR_TRY(bgr_identify_dynamics(n_episodes, max_rows, d_traj, d_rows, d_AB, d_ok, stream));
  BGR_CUDA(cudaMemcpyAsync(h_AB, d_AB, E * 6 * BGR_SYSID_COLS * sizeof(double), cudaMemcpyDeviceToHost, stream));
  BGR_CUDA(cudaMemcpyAsync(h_ok, d_ok, E * sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
  BGR_CUDA(cudaStreamSynchronize(stream));
  return BGR_OK;
}

// ================================================================================================
// FP64 peak probe (DFMA chains): the measured roofline denominator of the MPC evaluator
// ================================================================================================
extern "C" int bgr_fp64_peak_probe(int device, double* out_tflops) {
  int n_dev = 0;
  BGR_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    set_error("bgr_fp64_peak_probe: device %d out of range", device);
    return BGR_ERR_BAD_ARG;
  }
  DeviceScope dev(device);
  DeviceInfo di;
  BGR_TRY(device_info(device, &di));
  const int grid = di.sm_count * 8;
  const int iters = 1 << 12;
  double* d_out = nullptr;
  BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_out), static_cast<size_t>(grid) * 256 * sizeof(double)));
  cudaEvent_t a, b;
  BGR_CUDA(cudaEventCreate(&a));
  BGR_CUDA(cudaEventCreate(&b));
  double best_ms = 1e30;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a, nullptr);
    cudaError_t err = launch_fp64_probe(d_out, grid, iters, nullptr);
    cudaEventRecord(b, nullptr);
    if (err != cudaSuccess || cudaEventSynchronize(b) != cudaSuccess) {
      cudaFree(d_out);
      return check_cuda(err != cudaSuccess ? err : cudaGetLastError(), "fp64 probe");
    }
    count_launch();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0) best_ms = std::min(best_ms, static_cast<double>(ms));
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_out);
  const double fmas = static_cast<double>(grid) * 256.0 * iters * 8.0;   // 8 independent DFMA chains per thread
  if (out_tflops != nullptr) *out_tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
  return BGR_OK;
}

// ================================================================================================
// FP32 peak probe
// ================================================================================================
extern "C" int bgr_fp32_peak_probe(int device, double* out_tflops, double* out_ginst_per_s) {
  int n_dev = 0;
  BGR_CUDA(cudaGetDeviceCount(&n_dev));
  if (device < 0 || device >= n_dev) {
    set_error("bgr_fp32_peak_probe: device %d out of range", device);
    return BGR_ERR_BAD_ARG;
  }
  DeviceScope dev(device);
  DeviceInfo di;
  BGR_TRY(device_info(device, &di));
  const int grid = di.sm_count * 8;
  const int iters = 1 << 14;
  float* d_out = nullptr;
  BGR_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_out), static_cast<size_t>(grid) * 256 * sizeof(float)));
  cudaEvent_t a, b;
  BGR_CUDA(cudaEventCreate(&a));
  BGR_CUDA(cudaEventCreate(&b));
  double best_ms = 1e30;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(a, nullptr);
    cudaError_t err = launch_fp32_probe(d_out, grid, iters, nullptr);
    cudaEventRecord(b, nullptr);
    if (err != cudaSuccess || cudaEventSynchronize(b) != cudaSuccess) {
      cudaFree(d_out);
      return check_cuda(err != cudaSuccess ? err : cudaGetLastError(), "fp32 probe");
    }
    count_launch();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    if (rep > 0) best_ms = std::min(best_ms, static_cast<double>(ms));
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d_out);
  // 16 independent FMA chains per thread per iteration
  const double fmas = static_cast<double>(grid) * 256.0 * iters * 16.0;
  if (out_tflops != nullptr) *out_tflops = 2.0 * fmas / (best_ms * 1e-3) / 1e12;
  if (out_ginst_per_s != nullptr) *out_ginst_per_s = fmas / (best_ms * 1e-3) / 1e9;
  return BGR_OK;
}
