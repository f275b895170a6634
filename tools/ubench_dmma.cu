// ubench_dmma.cu -- microbenchmark: FP64 tensor (DMMA.8x8x4) and FP64 vector (DFMA) pipe rates on B200 and
// whether they overlap.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dmma ubench_dmma.cu
// Output: cycles per DMMA per SM, per DFMA warp-instruction per SM, for several warps/SM and mixes.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NMMA independent DMMA accumulators and NFMA independent DFMA accumulators per loop trip.
template <int NMMA, int NFMA, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_mix(int iters, double seed, double* out, long long* cyc) {
  double acc[2 * (NMMA > 0 ? NMMA : 1)], f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < 2 * NMMA; ++i) acc[i] = 0.0;
#pragma unroll
  for (int i = 0; i < NFMA; ++i) f[i] = seed * i;
  double a = seed + threadIdx.x * 1e-9, b = seed * 0.5;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < (NMMA > NFMA ? NMMA : NFMA); ++i) {
      if (i < NMMA) dmma884(acc[2 * i], acc[2 * i + 1], a, b);
      if (i < NFMA) f[i] = fma(f[i], a, b);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 2 * NMMA; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < NFMA; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int NMMA, int NFMA, int MAXT>
void run(int warps, int iters, const char* name) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 148 * 8);
  k_mix<NMMA, NFMA, MAXT><<<148, warps * 32>>>(10, 1.0, out, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_mix<NMMA, NFMA, MAXT><<<148, warps * 32>>>(iters, 1.0, out, cyc);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double mma = double(NMMA) * iters * warps, fm = double(NFMA) * iters * warps;
  printf("%-28s warps/SM=%2d  cycles=%.0f  ms=%.3f  clk/DMMA/SM=%6.2f  clk/DFMAwarp/SM=%6.2f  DMMA TF=%.1f DFMA TF=%.1f err=%s\n",
         name, warps, c, ms, mma > 0 ? c / mma : 0.0, fm > 0 ? c / fm : 0.0,
         mma * 512 * 148 / (ms * 1e-3) / 1e12, fm * 64 * 148 / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int it = 4000;
  for (int w : {4, 8}) run<36, 0, 256>(w, it, "DMMA x36 only");
  for (int w : {4, 8, 16, 32}) run<8, 0, 1024>(w, it, "DMMA x8 only");
  for (int w : {4, 8, 16}) run<18, 0, 512>(w, it, "DMMA x18 only");
  for (int w : {4, 8}) run<0, 36, 256>(w, it, "DFMA x36 only");
  for (int w : {4, 8, 16, 32}) run<0, 8, 1024>(w, it, "DFMA x8 only");
  for (int w : {4, 8}) run<36, 36, 256>(w, it, "DMMA x36 + DFMA x36");
  for (int w : {4, 8}) run<36, 8, 256>(w, it, "DMMA x36 + DFMA x8");
  for (int w : {4, 8}) run<18, 36, 256>(w, it, "DMMA x18 + DFMA x36");
  for (int w : {8, 16, 32}) run<8, 8, 1024>(w, it, "DMMA x8 + DFMA x8");
  for (int w : {8, 16, 32}) run<4, 12, 1024>(w, it, "DMMA x4 + DFMA x12");
  return 0;
}
