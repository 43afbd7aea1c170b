/*
 * c_abi_demo.c -- the C ABI of libbgr_rollout.so used from plain C (no Python, no torch, no CUDA headers).
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -o c_abi_demo -Lbgr_pathplanning_control_b200 -lbgr_rollout -lm \
 *       -Wl,-rpath,$PWD/bgr_pathplanning_control_b200
 *   ./c_abi_demo [episodes] [steps]
 *
 * Builds an elliptic centreline (the recipe of the reference-frozen golden vectors: a = 40, b = 20), sweeps the
 * pure-pursuit look-ahead over the episodes, runs one batched closed-loop rollout through bgr_rollout_host and prints
 * the aggregate row plus the first episodes' lap metrics -- the numbers preformance_analysis/metrics_calculator.py
 * would compute from each episode's log.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bgr_rollout.h"

#define CHECK(call)                                                            \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ != BGR_OK) {                                                       \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, bgr_last_error_string()); \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  const long E = argc > 1 ? atol(argv[1]) : 4096;
  const int S = argc > 2 ? atoi(argv[2]) : 800;
  const int n = 720;
  const double PI = 3.14159265358979323846;
  double* x = malloc(sizeof(double) * n), *y = malloc(sizeof(double) * n), *kappa = malloc(sizeof(double) * n);
  for (int i = 0; i < n; ++i) {
    const double th = 2.0 * PI * i / n;
    x[i] = 40.0 * cos(th);
    y[i] = 20.0 * sin(th);
    kappa[i] = 40.0 * 20.0 / pow(40.0 * 40.0 * sin(th) * sin(th) + 20.0 * 20.0 * cos(th) * cos(th), 1.5);
  }
  bgr_track_desc_t desc;
  memset(&desc, 0, sizeof(desc));
  desc.h_x = x; desc.h_y = y; desc.h_kappa = kappa; desc.n = n; desc.closed = 1; desc.cone_reach = 4.0;
  bgr_track_t* track = NULL;
  CHECK(bgr_track_create(&desc, 0, &track));

  bgr_rollout_cfg_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.steer_kind = BGR_STEER_PURE_PURSUIT; cfg.accel_kind = BGR_ACCEL_PID; cfg.model_kind = BGR_MODEL_KINEMATIC;
  cfg.precision = BGR_PRECISION_FP32; cfg.max_steps = S; cfg.laps = 4;
  cfg.dt = 0.05; cfg.wheelbase = 1.8; cfg.l_f = 0.9; cfg.l_r = 0.9;                 /* vehicle_config.py:9-16,39 */
  cfg.max_steer = 30.0 * PI / 180.0; cfg.max_accel = 1.5; cfg.max_decel = -2.0;     /* :30-35 */

  float* params = calloc((size_t)E * BGR_N_PARAMS, sizeof(float));
  double* init = calloc((size_t)E * BGR_N_STATE, sizeof(double));
  for (long e = 0; e < E; ++e) {
    float* p = params + e * BGR_N_PARAMS;
    p[BGR_P_LF] = 3.0f + 5.0f * (float)e / (float)(E > 1 ? E - 1 : 1);              /* look-ahead sweep 3..8 m */
    p[BGR_P_KP] = 0.35f; p[BGR_P_KI] = 0.4f; p[BGR_P_KD] = 0.3f;                   /* :19-21 */
    p[BGR_P_TARGET_SPEED] = (float)(25.0 / 3.6); p[BGR_P_K_EXPO] = 0.05f;          /* :14, :24 */
    p[BGR_P_K_STANLEY] = 1.0f; p[BGR_P_K_SOFT] = 0.1f;
    p[BGR_P_MU] = 1.0f; p[BGR_P_C_F] = 16000.f; p[BGR_P_C_R] = 17000.f; p[BGR_P_MASS] = 255.f; p[BGR_P_I_Z] = 1700.f;
    double* s = init + e * BGR_N_STATE;
    s[BGR_X] = 40.0; s[BGR_Y] = 0.0; s[BGR_YAW] = PI / 2.0; s[BGR_V] = 0.0;
  }
  float* metrics = malloc(sizeof(float) * (size_t)E * BGR_N_METRICS);
  int32_t* status = malloc(sizeof(int32_t) * (size_t)E * BGR_N_STATUS);
  double agg[BGR_N_AGG];
  CHECK(bgr_rollout_host(track, &cfg, E, params, init, metrics, status, NULL, agg, NULL));
  double ms = 0.0;
  CHECK(bgr_last_kernel_ms(&ms));

  printf("episodes %ld steps %.0f kernel_ms_last_chunk %.3f launches %lld\n", E, agg[BGR_A_STEPS], ms,
         (long long)bgr_launch_count());
  printf("aggregate sum_lat_mean %.9g sum_lat_std %.9g sum_lap_time %.9g finished %.0f\n", agg[BGR_A_SUM_LAT_MEAN],
         agg[BGR_A_SUM_LAT_STD], agg[BGR_A_SUM_LAP_TIME], agg[BGR_A_FINISHED]);
  for (long e = 0; e < E && e < 3; ++e) {
    const float* m = metrics + e * BGR_N_METRICS;
    printf("episode %ld Lf %.4f lat_mean %.9g lat_std %.9g head_mean %.9g head_std %.9g lap_time %.6f laps %d\n", e,
           params[e * BGR_N_PARAMS + BGR_P_LF], m[BGR_M_LAT_MEAN], m[BGR_M_LAT_STD], m[BGR_M_HEAD_MEAN],
           m[BGR_M_HEAD_STD], m[BGR_M_LAP_TIME], status[e * BGR_N_STATUS + BGR_S_LAPS]);
  }
  bgr_track_destroy(track);
  free(x); free(y); free(kappa); free(params); free(init); free(metrics); free(status);
  return 0;
}
