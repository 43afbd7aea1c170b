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

}  // namespace bgr
