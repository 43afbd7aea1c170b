// Host-side track handle: derived per-waypoint tables + device residency.
#pragma once

#include <mutex>
#include <vector>

#include "bgr_device.cuh"

struct bgr_track {
  int device = 0;
  int n = 0, n_pad = 0, closed = 0, n_cones = 0;
  double cone_reach = 0.0;
  // device tables
  float2* xy_f = nullptr;
  float4* attr_f = nullptr;
  double2* xy_d = nullptr;
  bgr::D4* attr_d = nullptr;
  float* guard2 = nullptr;
  int2* rival_range = nullptr;     // per waypoint (first, count) into rival_list; count < 0: no usable list
  int2* rival_list = nullptr;      // (waypoint, float bits of the squared distance from w_j at which it can first win)
  int2* twin = nullptr;            // per waypoint (twin waypoint t or -1, float bits of the squared twin guard)
  float4* scan_blk = nullptr;      // bounding circles of blocks of 32 waypoints (exact-scan pruning)
  int32_t* cone_range = nullptr;
  float* cone_near = nullptr;   // [n] distance from waypoint j to its closest listed cone (rounded down)
  int32_t* cone_list = nullptr;
  double2* cones = nullptr;
  float2* cones_f = nullptr;
  // LQG Kalman gain sequences, uploaded on first use per (dt, length); freed with the handle
  struct KfTable {
    double dt = 0;
    int len = 0;
    double K[2] = {0, 0};   // LQR gain of control.dlqr for this dt
    double* d_gain = nullptr;
    float2* d_gain_f = nullptr;
  };
  mutable std::mutex kf_mu;
  mutable std::vector<KfTable> kf_tables;
  // host copies
  std::vector<double> x, y, kappa, path_yaw, guard_radius;
  double ds_min = 0, ds_max = 0, guard_min = 0, guard_median = 0;
  double guard_tight_frac = 0;     // share of waypoints whose exactness guard is under kTightGuard (lines driven twice, crossings)
  long long rival_total = 0;
  int rival_max = 0, rival_unlisted = 0, twin_count = 0;

  bgr::TrackView view() const {
    bgr::TrackView v;
    v.xy_f = xy_f; v.attr_f = attr_f; v.xy_d = xy_d; v.attr_d = attr_d; v.guard2 = guard2;
    v.rival_range = rival_range; v.rival_list = rival_list; v.twin = twin; v.scan_blk = scan_blk;
    v.cone_range = cone_range; v.cone_near = cone_near; v.cone_list = cone_list; v.cones = cones; v.cones_f = cones_f;
    v.n = n; v.n_pad = n_pad; v.closed = closed; v.n_cones = n_cones;
    v.ds_max = static_cast<float>(ds_max);
    return v;
  }
};

namespace bgr {
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t err, const char* what);
void count_launch(int n = 1);
}  // namespace bgr
