// Issue-rate probe for the packed fp32 instructions of sm_100a (FFMA2 / FADD2 / FMUL2) next to their scalar forms.
// Question it answers: does one packed instruction cost one issue slot (then packing the x / y halves of the waypoint
// distance evaluations halves their issue cost in the instruction-issue-bound step kernel) or two?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

__device__ __forceinline__ uint64_t pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack(uint64_t r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }

// MODE 0: 2 x kChains scalar FFMA per iteration; 1: kChains FFMA2; 2: 2 x kChains FADD; 3: kChains FADD2;
// 4: kChains FFMA2 + kChains IMAD (do packed and integer instructions share issue slots one for one?)
// 5: 2 x kChains FFMA + kChains IMAD; 6 / 7: the same two with LOP3 (ALU pipe, not the FMA pipe IMAD shares)
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, float a, float b, int ia) {
  float x[2 * kChains];
  uint64_t p[kChains];
  int z[kChains];
#pragma unroll
  for (int i = 0; i < 2 * kChains; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < kChains; ++i) { p[i] = pack(x[2 * i], x[2 * i + 1]); z[i] = threadIdx.x + i; }
  const uint64_t pa = pack(a, a), pb = pack(b, b);
  for (int k = 0; k < kIters; ++k) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (MODE == 0 || MODE == 5 || MODE == 7) { x[2 * i] = fmaf(x[2 * i], a, b); x[2 * i + 1] = fmaf(x[2 * i + 1], a, b); }
      if (MODE == 1 || MODE == 4 || MODE == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
      if (MODE == 2) { x[2 * i] = __fadd_rn(x[2 * i], a); x[2 * i + 1] = __fadd_rn(x[2 * i + 1], a); }
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pa));
      if (MODE == 4 || MODE == 5) z[i] = z[i] * ia + k;
      if (MODE == 6 || MODE == 7) z[i] = (z[i] ^ ia) | k;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < kChains; ++i) {
    float u, v;
    unpack(p[i], u, v);
    s += u + v + x[2 * i] + x[2 * i + 1] + z[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double inst_per_iter, float* d) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(d, 0.999f, 1e-3f, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  const double warp_inst = static_cast<double>(grid) * 8 * kIters * inst_per_iter;   // 8 warps per CTA
  printf("%-28s %8.3f ms  %7.3f warp-inst/ns  = %5.2f warp-inst/clk/SM at the max clock %d MHz\n", name, best,
         warp_inst / (best * 1e6), warp_inst / (best * 1e-3) / (clk * 1e3) / sms, clk / 1000);
}

int main() {
  float* d;
  cudaMalloc(&d, 148 * 8 * 256 * 4 * 2);
  run<0>("16 FFMA", 16, d);
  run<1>("8 FFMA2", 8, d);
  run<2>("16 FADD", 16, d);
  run<3>("8 FADD2", 8, d);
  run<5>("16 FFMA + 8 IMAD", 24, d);
  run<4>("8 FFMA2 + 8 IMAD", 16, d);
  run<7>("16 FFMA + 8 LOP3", 24, d);
  run<6>("8 FFMA2 + 8 LOP3", 16, d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
