// Small kernels: the fixed-order reduction of per-CTA aggregates and the FP32 FMA-chain probe that
// measures the roofline denominator of the step kernels.
#include "bgr_launch.cuh"

namespace bgr {

// Fixed-order final reduction of the per-CTA partials (deterministic: no atomics).  One warp per
// aggregate slot: lanes stride over the CTAs in a fixed pattern, then a fixed shuffle tree.
__global__ void reduce_partials_kernel(const double* __restrict__ partials, int n_blocks, double* __restrict__ out) {
  const int a = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (a >= BGR_N_AGG) return;
  const bool is_min = a == BGR_A_MIN_LAP_TIME;
  double x = is_min ? 1e300 : 0.0;
  for (int b = lane; b < n_blocks; b += 32) {
    const double o = partials[static_cast<size_t>(b) * BGR_N_AGG + a];
    x = is_min ? fmin(x, o) : x + o;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_down_sync(0xffffffffu, x, off);
    x = is_min ? fmin(x, o) : x + o;
  }
  if (lane == 0) out[a] = x;
}

// Per-256-episode partial aggregates from the per-episode OUTPUT rows (fixed order: lane shuffle tree, then one
// thread per slot over the warps): deterministic, no atomics, and independent of how the rollout was launched
// (single launch, pipelined chunks, two-phase).
__global__ void __launch_bounds__(256) aggregate_partials_kernel(const float* __restrict__ metrics,
                                                                 const int32_t* __restrict__ status, long long n_episodes,
                                                                 double* __restrict__ partials) {
  __shared__ double red[8 * BGR_N_AGG];
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = e < n_episodes;
  double vals[BGR_N_AGG];
  int steps = 0, flags = 0, hits = 0;
  float lat_mean = 0.f, lat_std = 0.f, lap_time = 0.f;
  if (active) {
    const float* mrow = metrics + e * BGR_N_METRICS;
    const int32_t* srow = status + e * BGR_N_STATUS;
    lat_mean = mrow[BGR_M_LAT_MEAN]; lat_std = mrow[BGR_M_LAT_STD]; lap_time = mrow[BGR_M_LAP_TIME];
    flags = srow[BGR_S_FLAGS]; steps = srow[BGR_S_STEPS]; hits = srow[BGR_S_CONE_HITS];
  }
  const bool finished = active && (flags & (BGR_ST_PATH_END | BGR_ST_LAPS_DONE));
  vals[BGR_A_EPISODES] = active ? 1.0 : 0.0;
  vals[BGR_A_STEPS] = static_cast<double>(steps);
  vals[BGR_A_SUM_LAT_MEAN] = active ? static_cast<double>(lat_mean) : 0.0;
  vals[BGR_A_SUM_LAT_STD] = (active && steps > 1) ? static_cast<double>(lat_std) : 0.0;
  vals[BGR_A_SUM_LAP_TIME] = active ? static_cast<double>(lap_time) : 0.0;
  vals[BGR_A_MIN_LAP_TIME] = finished ? static_cast<double>(lap_time) : 1e300;
  vals[BGR_A_CONE_HITS] = static_cast<double>(hits);
  vals[BGR_A_FINISHED] = finished ? 1.0 : 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < BGR_N_AGG; ++a) {
    double x = vals[a];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_down_sync(0xffffffffu, x, off);
      x = (a == BGR_A_MIN_LAP_TIME) ? fmin(x, o) : x + o;
    }
    if (lane == 0) red[warp * BGR_N_AGG + a] = x;
  }
  __syncthreads();
  if (threadIdx.x < BGR_N_AGG) {
    const int a = threadIdx.x;
    double x = red[a];
    for (int w = 1; w < 8; ++w) {
      const double o = red[w * BGR_N_AGG + a];
      x = (a == BGR_A_MIN_LAP_TIME) ? fmin(x, o) : x + o;
    }
    partials[static_cast<size_t>(blockIdx.x) * BGR_N_AGG + a] = x;
  }
}

cudaError_t launch_aggregate_partials(const float* metrics, const int32_t* status, long long n_episodes,
                                      double* partials, cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>((n_episodes + 255) / 256);
  aggregate_partials_kernel<<<grid, 256, 0, stream>>>(metrics, status, n_episodes, partials);
  return cudaGetLastError();
}

// Survivors of a capped first phase (status TIMEOUT with steps == cap): their rows, densely packed (order is
// whatever the atomics give -- results do not depend on it), and their number.
__global__ void __launch_bounds__(256) compact_survivors_kernel(const int32_t* __restrict__ status, long long n_episodes,
                                                                int32_t* __restrict__ idx, int32_t* __restrict__ count) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool alive = e < n_episodes && status[e * BGR_N_STATUS + BGR_S_FLAGS] == BGR_ST_TIMEOUT;
  const unsigned mask = __ballot_sync(0xffffffffu, alive);
  if (mask == 0u) return;
  const int lane = threadIdx.x & 31;
  int base = 0;
  if (lane == __ffs(mask) - 1) base = atomicAdd(count, __popc(mask));
  base = __shfl_sync(0xffffffffu, base, __ffs(mask) - 1);
  if (alive) idx[base + __popc(mask & ((1u << lane) - 1u))] = static_cast<int32_t>(e);
}

cudaError_t launch_compact_survivors(const int32_t* status, long long n_episodes, int32_t* idx, int32_t* count,
                                     cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>((n_episodes + 255) / 256);
  compact_survivors_kernel<<<grid, 256, 0, stream>>>(status, n_episodes, idx, count);
  return cudaGetLastError();
}

// ---- launch order by target speed (bucket sort) ----
// An episode that stops at the end of the path or at a lap target runs for about (path length) / (target speed) steps, and
// a warp runs until its slowest lane is done: with speeds drawn at random, a quarter of the lane-steps of BASELINE config 5
// belong to lanes that have already finished.  Launching the episodes in order of target speed puts lanes of like length
// into one warp (and, since they also travel together, into the same branch of the step at the same time).  Only the
// launch order changes: thread t works on row perm[t] and writes row perm[t]; an episode is a function of its own row.
// Key = the upper 16 bits of the speed's float pattern (sign, exponent, 7 mantissa bits: 1/128 relative resolution);
// slowest first, so the longest episodes start first and the short ones fill the tail.  The order inside a bucket is
// whatever the atomics give -- results do not depend on it.
__device__ __forceinline__ int order_bucket(const float* __restrict__ params, long long e) {
  const unsigned bits = __float_as_uint(params[e * BGR_N_PARAMS + BGR_P_TARGET_SPEED]);
  return static_cast<int>(min(bits >> 16, static_cast<unsigned>(kOrderBuckets - 1)));   // (negative / NaN: last bucket)
}

__global__ void __launch_bounds__(256) order_hist_kernel(const float* __restrict__ params, long long n_episodes,
                                                         int32_t* __restrict__ hist) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n_episodes) atomicAdd(hist + order_bucket(params, e), 1);
}

// one CTA of 1024 threads: exclusive prefix sum of the kOrderBuckets counters, in place (they become the write cursors)
__global__ void __launch_bounds__(1024) order_scan_kernel(int32_t* __restrict__ hist) {
  constexpr int kPer = kOrderBuckets / 1024;
  __shared__ int warp_sums[32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  int v[kPer], sum = 0;
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    v[i] = hist[t * kPer + i];
    sum += v[i];
  }
  int inc = sum;                                      // inclusive scan of the per-thread sums across the warp ...
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) warp_sums[w] = inc;
  __syncthreads();
  if (w == 0) {                                       // ... and of the warp totals across the CTA
    int ws = warp_sums[lane];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, ws, d);
      if (lane >= d) ws += o;
    }
    warp_sums[lane] = ws;
  }
  __syncthreads();
  int run = inc - sum + (w > 0 ? warp_sums[w - 1] : 0);
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    hist[t * kPer + i] = run;
    run += v[i];
  }
}

__global__ void __launch_bounds__(256) order_scatter_kernel(const float* __restrict__ params, long long n_episodes,
                                                            int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e < n_episodes) perm[atomicAdd(cursor + order_bucket(params, e), 1)] = static_cast<int32_t>(e);
}

cudaError_t launch_order_by_speed(const float* params, long long n_episodes, int32_t* perm, int32_t* hist,
                                  cudaStream_t stream) {
  cudaError_t err = cudaMemsetAsync(hist, 0, kOrderBuckets * sizeof(int32_t), stream);
  if (err != cudaSuccess) return err;
  const unsigned grid = static_cast<unsigned>((n_episodes + 255) / 256);
  order_hist_kernel<<<grid, 256, 0, stream>>>(params, n_episodes, hist);
  order_scan_kernel<<<1, 1024, 0, stream>>>(hist);
  order_scatter_kernel<<<grid, 256, 0, stream>>>(params, n_episodes, hist, perm);
  return cudaGetLastError();
}

cudaError_t launch_reduce_partials(const double* partials, int n_blocks, double* out, cudaStream_t stream) {
  reduce_partials_kernel<<<1, 32 * BGR_N_AGG, 0, stream>>>(partials, n_blocks, out);
  return cudaGetLastError();
}

// 16 independent FFMA chains per thread: enough ILP to saturate both FP32 datapaths of an SM
// sub-partition at 8 resident warps.
__global__ void __launch_bounds__(256) fp32_probe_kernel(float* __restrict__ out, int iters) {
  float acc[16];
  const float a = 1.0f + 1e-7f * static_cast<float>(threadIdx.x), b = 1e-9f * static_cast<float>(blockIdx.x);
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = static_cast<float>(i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

cudaError_t launch_fp32_probe(float* out, int grid, int iters, cudaStream_t stream) {
  fp32_probe_kernel<<<grid, 256, 0, stream>>>(out, iters);
  return cudaGetLastError();
}

// the same with DFMA chains: the measured FP64 CUDA-core peak (roofline denominator of the MPC evaluator)
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* __restrict__ out, int iters) {
  double acc[8];
  const double a = 1.0 + 1e-12 * static_cast<double>(threadIdx.x), b = 1e-15 * static_cast<double>(blockIdx.x);
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = static_cast<double>(i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x] = s;
}

cudaError_t launch_fp64_probe(double* out, int grid, int iters, cudaStream_t stream) {
  fp64_probe_kernel<<<grid, 256, 0, stream>>>(out, iters);
  return cudaGetLastError();
}

}  // namespace bgr
