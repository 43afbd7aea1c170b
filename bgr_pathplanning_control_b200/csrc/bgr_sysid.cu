// Batched system identification: for every episode, the least-squares fit of
//     x_{k+1} = A x_k + B u_k + c          x = [x, y, yaw, v, yaw_rate, beta],  u = [accel, steering]
// over its logged rows -- identify_dynamics (old/test_runs_system_id.py:84-111: sklearn LinearRegression with an
// intercept on np.hstack([X, U]) -> Y), the one place on this side of the reference where a small dense solve appears.
// One CTA of 64 threads per episode: column means (pass 1), centred Gram matrix Z^T Z (8 x 8) and cross matrix
// Z^T Y (8 x 6) accumulated in fp64 registers, one entry per thread (pass 2), then an 8 x 8 Cholesky solve of the
// column-scaled normal equations.  No tensor cores: 8 x 8 per episode is far below any MMA tile.
// Rank-deficient logs (e.g. the kinematic model, whose yaw_rate and beta columns are constant) are flagged ok = 0 and
// return NaN: LinearRegression would return the minimum-norm solution there (documented deviation).
#include "bgr_launch.cuh"

namespace bgr {

constexpr int kSysidThreads = 64;
constexpr int kNZ = 8;   // regressors: 6 states + 2 inputs
constexpr int kNY = 6;   // targets: next state

__device__ __forceinline__ double z_of(const double* row, int c) {   // regressor c of a trajectory row
  return c < 6 ? row[c] : (c == 6 ? row[BGR_T_ACCEL] : row[BGR_T_STEER]);
}

__global__ void __launch_bounds__(kSysidThreads) sysid_kernel(const double* __restrict__ traj, const int32_t* __restrict__ rows,
                                                              int max_rows, long long n_episodes, double* __restrict__ AB,
                                                              int32_t* __restrict__ ok) {
  const long long e = blockIdx.x;
  if (e >= n_episodes) return;
  const int t = threadIdx.x;
  const int T = min(rows[e], max_rows) - 1;            // number of (x_k, u_k) -> x_{k+1} pairs
  const double* tr = traj + static_cast<size_t>(e) * max_rows * BGR_N_TRAJ;
  __shared__ double s_mean[kNZ + kNY];
  __shared__ double s_G[kNZ][kNZ];
  __shared__ double s_C[kNZ][kNY];
  __shared__ double s_scale[kNZ];
  __shared__ int s_ok;

  // ---- pass 1: column means (threads 0..13, one column each; rows are L2-resident for the second pass) ----
  if (t < kNZ + kNY) {
    double sum = 0.0;
    for (int k = 0; k < T; ++k) {
      const double* row = tr + static_cast<size_t>(k) * BGR_N_TRAJ;
      sum += t < kNZ ? z_of(row, t) : row[BGR_N_TRAJ + (t - kNZ)];   // target = state column of the NEXT row
    }
    s_mean[t] = T > 0 ? sum / T : 0.0;
  }
  __syncthreads();

  // ---- pass 2: centred Gram (thread t owns G[i][j]) and cross (threads 0..47 own C[i][c]) ----
  const int gi = t >> 3, gj = t & 7;
  const int ci = t / kNY, cc = t % kNY;
  double g = 0.0, c = 0.0;
  const double mi = s_mean[gi], mj = s_mean[gj];
  const double mci = t < kNZ * kNY ? s_mean[ci] : 0.0, mcc = t < kNZ * kNY ? s_mean[kNZ + cc] : 0.0;
  for (int k = 0; k < T; ++k) {
    const double* row = tr + static_cast<size_t>(k) * BGR_N_TRAJ;
    g = fma(z_of(row, gi) - mi, z_of(row, gj) - mj, g);
    if (t < kNZ * kNY) c = fma(z_of(row, ci) - mci, row[BGR_N_TRAJ + cc] - mcc, c);
  }
  s_G[gi][gj] = g;
  if (t < kNZ * kNY) s_C[ci][cc] = c;
  __syncthreads();

  // ---- column scaling (unit diagonal), Cholesky, solves: thread 0 factorises, threads 0..5 solve one target each ----
  if (t < kNZ) s_scale[t] = s_G[t][t] > 0.0 ? rsqrt(s_G[t][t]) : 0.0;
  __syncthreads();
  s_G[gi][gj] = s_G[gi][gj] * s_scale[gi] * s_scale[gj];
  __syncthreads();
  if (t == 0) {
    int good = T >= kNZ + 1;
    for (int j = 0; j < kNZ && good; ++j) {
      double d = s_G[j][j];
      for (int k = 0; k < j; ++k) d -= s_G[j][k] * s_G[j][k];
      if (!(d > 1e-13)) {          // scaled pivots are O(1): anything this small is a (numerically) dependent column
        good = 0;
        break;
      }
      d = sqrt(d);
      s_G[j][j] = d;
      for (int i = j + 1; i < kNZ; ++i) {
        double v = s_G[i][j];
        for (int k = 0; k < j; ++k) v -= s_G[i][k] * s_G[j][k];
        s_G[i][j] = v / d;
      }
    }
    s_ok = good;
  }
  __syncthreads();
  double* out = AB + static_cast<size_t>(e) * kNY * (kNZ + 1);   // row r = [A[r][0..5] | B[r][0..1] | c[r]]
  if (t < kNY) {
    if (s_ok) {
      double w[kNZ];
      for (int i = 0; i < kNZ; ++i) {                  // L y = D C[:, t]
        double v = s_C[i][t] * s_scale[i];
        for (int k = 0; k < i; ++k) v -= s_G[i][k] * w[k];
        w[i] = v / s_G[i][i];
      }
      for (int i = kNZ - 1; i >= 0; --i) {             // L^T u = y
        double v = w[i];
        for (int k = i + 1; k < kNZ; ++k) v -= s_G[k][i] * w[k];
        w[i] = v / s_G[i][i];
      }
      double icpt = s_mean[kNZ + t];
      for (int i = 0; i < kNZ; ++i) {
        w[i] *= s_scale[i];                            // undo the column scaling
        out[t * (kNZ + 1) + i] = w[i];
        icpt -= s_mean[i] * w[i];
      }
      out[t * (kNZ + 1) + kNZ] = icpt;                 // LinearRegression.intercept_
    } else {
      const double nan_ = __longlong_as_double(0x7ff8000000000000LL);
      for (int i = 0; i <= kNZ; ++i) out[t * (kNZ + 1) + i] = nan_;
    }
  }
  if (t == 0) ok[e] = s_ok;
}

cudaError_t launch_sysid(const double* traj, const int32_t* rows, int max_rows, long long n_episodes, double* AB,
                         int32_t* ok, cudaStream_t stream) {
  sysid_kernel<<<static_cast<unsigned>(n_episodes), kSysidThreads, 0, stream>>>(traj, rows, max_rows, n_episodes, AB, ok);
  return cudaGetLastError();
}

}  // namespace bgr
