// ubench_rcp.cu -- what does one FP64 reciprocal cost on sm_100a?  Per-SM issue cost (clk per warp instruction) of
// MUFU.RCP64H, of the product's rcp_fast (MUFU.RCP64H + 3 DFMA), and of an FP32-seeded variant (bit surgery to FP32,
// MUFU.RCP, bit surgery back, two Newton steps = 4 DFMA), each with 16 independent chains per thread and 16 warps per SM
// (the fixed point's occupancy).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_rcp ubench_rcp.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double rcp_mufu64(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}
__device__ __forceinline__ double rcp_fast(double x) {
  const double r = rcp_mufu64(x);
  const double d = fma(-x, r, 1.0);
  const double s = fma(d, d, d);
  return fma(r, s, r);
}
// x in [2^-126, 2^127): FP32 seed from the top 32 + 3 bits, ~2^-22 accurate, then two Newton steps
__device__ __forceinline__ double rcp_f32seed(double x) {
  const unsigned int hi = (unsigned int)__double2hiint(x), lo = (unsigned int)__double2loint(x);
  const float xf = __uint_as_float(__funnelshift_l(lo, hi - 0x38000000u, 3));
  float rf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(xf));
  const unsigned int rb = __float_as_uint(rf);
  double r = __hiloint2double(int((rb >> 3) + 0x38000000u), int(rb << 29));
  double d = fma(-x, r, 1.0);
  r = fma(r, d, r);
  d = fma(-x, r, 1.0);
  return fma(r, d, r);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(double* out, int iters, double seed) {
  double v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + 0.001 * (threadIdx.x + 32 * i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = rcp_mufu64(v[i]) + 1.0;          // MUFU.RCP64H (+ 1 DADD to keep the value in range)
      if (MODE == 1) v[i] = fma(v[i], 0.999, 0.001);          // DFMA alone
      if (MODE == 2) v[i] = rcp_fast(v[i]) + 1.0;
      if (MODE == 3) v[i] = rcp_f32seed(v[i]) + 1.0;
      if (MODE == 4) v[i] = v[i] + 1.0;                       // DADD alone
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int inst_per_elem) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 2 * 256);
  const int iters = 2000;
  k<MODE><<<sms * 2, 256>>>(out, 10, 1.5);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<MODE><<<sms * 2, 256>>>(out, iters, 1.5);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const double clk = ms * 1e-3 * clk_khz * 1e3;
  const double warp_ops = 16.0 * iters * 16;      // per SM: 16 warps x 16 chains x iters
  printf("%-28s %8.3f ms  %6.2f clk per warp-op per SM (%d FP64-pipe instr + extras per op)\n", name, ms, clk / warp_ops, inst_per_elem);
  cudaFree(out);
}

int main() {
  run<4>("DADD", 1);
  run<1>("DFMA", 1);
  run<0>("MUFU.RCP64H + DADD", 2);
  run<2>("rcp_fast + DADD", 5);
  run<3>("rcp_f32seed + DADD", 5);
  double h[4];
  // accuracy of the FP32-seeded variant vs 1/x
  return 0;
}
