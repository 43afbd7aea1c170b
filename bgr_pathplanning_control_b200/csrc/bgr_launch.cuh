// Declarations of the per-configuration rollout launchers (defined in bgr_rollout_inst_*.cu).
#pragma once

#include "bgr_device.cuh"

namespace bgr {

template <typename Real, int STEER, int ACCEL, int MODEL, bool EXTRAS>
// `query_resident` != nullptr: no launch -- report how many episodes (threads) one SM holds at once with this kernel
cudaError_t launch_rollout(const RolloutArgs& args, int grid, size_t smem, int smem_optin, cudaStream_t stream,
                           int* query_resident);

#define BGR_DECLARE_LAUNCH(Real, STEER, ACCEL, MODEL, EXTRAS) \
  template <>                                                \
  cudaError_t launch_rollout<Real, STEER, ACCEL, MODEL, EXTRAS>(const RolloutArgs&, int, size_t, int, cudaStream_t, int*);

#define BGR_FOR_ALL_LAWS(M, Real, EXTRAS) \
  M(Real, 0, 0, 0, EXTRAS)                \
  M(Real, 0, 0, 1, EXTRAS)                \
  M(Real, 0, 1, 0, EXTRAS)                \
  M(Real, 0, 1, 1, EXTRAS)                \
  M(Real, 1, 0, 0, EXTRAS)                \
  M(Real, 1, 0, 1, EXTRAS)                \
  M(Real, 1, 1, 0, EXTRAS)                \
  M(Real, 1, 1, 1, EXTRAS)

BGR_FOR_ALL_LAWS(BGR_DECLARE_LAUNCH, float, false)
BGR_FOR_ALL_LAWS(BGR_DECLARE_LAUNCH, float, true)
BGR_FOR_ALL_LAWS(BGR_DECLARE_LAUNCH, double, true)

// fp32 kernels on long tracks stage only the xy table (8 B per waypoint)
__host__ __device__ constexpr size_t rollout_smem_bytes_xy_only(int n_pad) {
  return static_cast<size_t>(n_pad) * sizeof(float2) + 64 + kRolloutThreads / 32 * BGR_N_AGG * sizeof(double);
}
constexpr size_t kSmemAllTablesLimit = 56 * 1024;   // up to here 4 CTAs per SM still fit with every table staged

template <typename Real>
__host__ __device__ constexpr size_t rollout_smem_bytes(int n_pad) {
  return static_cast<size_t>(n_pad) * (sizeof(typename RT<Real>::R2) + sizeof(typename RT<Real>::R4) + sizeof(float)) +
         64 + kRolloutThreads / 32 * BGR_N_AGG * sizeof(double);  // mbarrier slot + block-reduction scratch
}

// ---- MPC candidate evaluation (bgr_mpc.cu) ----
constexpr int kMpcThreads = 128;
constexpr int kMpcMaxH = 64;

struct MpcArgs {
  TrackView trk;
  const double* state;       // [E][4]
  const int32_t* ref_start;  // [E]
  const int32_t* ref_len;    // [E] or nullptr
  const float* bank;         // [K][H][2]
  unsigned long long* packed;  // [E], pre-set to ~0
  float* costs;              // nullptr or [E][K]
  long long n_episodes;
  int32_t K, H;
  double dt, wheelbase, target_speed;
  double q0, q1, q2, q3, r0, r1;
};

cudaError_t launch_mpc(const MpcArgs& A, int precision, cudaStream_t stream, int sm_count);
// factored evaluator (production): candidate statistics + episode coefficients + 4 FMAs per pair
cudaError_t launch_mpc_factored(const MpcArgs& A, void* scratch_cand, void* scratch_epi, cudaStream_t stream,
                                int sm_count, bool cand_ready = false);
cudaError_t launch_mpc_cand_stats(const MpcArgs& A, void* cand_out, cudaStream_t stream);
size_t mpc_cand_bytes(int K);
size_t mpc_epi_bytes(long long E);
cudaError_t launch_mpc_finalize(const unsigned long long* packed, const float* bank, int H, long long n_episodes,
                                float max_accel, float max_decel, float max_steer, float* best_u, float* best_cost,
                                int32_t* best_idx, cudaStream_t stream);

// ---- batched system identification (bgr_sysid.cu) ----
cudaError_t launch_sysid(const double* traj, const int32_t* rows, int max_rows, long long n_episodes, double* AB,
                         int32_t* ok, cudaStream_t stream);

// ---- small kernels (bgr_misc.cu) ----
cudaError_t launch_reduce_partials(const double* partials, int n_blocks, double* out, cudaStream_t stream);
cudaError_t launch_aggregate_partials(const float* metrics, const int32_t* status, long long n_episodes,
                                      double* partials, cudaStream_t stream);
cudaError_t launch_compact_survivors(const int32_t* status, long long n_episodes, int32_t* idx, int32_t* count,
                                     cudaStream_t stream);
// launch order by target speed: perm[count] = the rows, slowest first; hist = kOrderBuckets ints of scratch
constexpr int kOrderBuckets = 32768;
cudaError_t launch_order_by_speed(const float* params, long long n_episodes, int32_t* perm, int32_t* hist,
                                  cudaStream_t stream);
cudaError_t launch_fp32_probe(float* out, int grid, int iters, cudaStream_t stream);
cudaError_t launch_fp64_probe(double* out, int grid, int iters, cudaStream_t stream);

}  // namespace bgr
