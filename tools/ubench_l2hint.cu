// ubench_l2hint.cu -- does pinning the head of e[] in L2 pay for the single-GPU fixed point?  
//
// The fixed point re-reads the same 512 MiB vector ~29 times; B200's L2 holds 126 MB.  This micro-benchmark streams
// a vector of N doubles PASSES times through the same per-warp cp.async.bulk ring the product kernel uses
// (fixed_point.cu: 4 KiB slots, depth 3) and sums it, with the first HEAD MiB loaded under an L2 `evict_last`
// policy and the rest under `evict_first` (createpolicy + .L2::cache_hint on the bulk copy), and reports the time
// per pass and the effective GB/s for HEAD = 0 (no hints), 0 (all evict_first), 32, 64, 80, 96 MiB.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../rlvi_b200/csrc -I../include \
//        ubench_l2hint.cu -o ubench_l2hint && ./ubench_l2hint [log2n=26] [passes=29]
//
// Not part of the library; nothing links against it.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tma.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kDepth = 3;
constexpr int kSlotDoubles = 512;                       // 4 KiB
constexpr int kSlotBytes = kSlotDoubles * 8;

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}

// mode 0: plain bulk copies; mode 1: hinted (chunks below head_chunks evict_last, the rest evict_first)
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) pass_kernel(const double* __restrict__ e, int64_t nchunks,
                                                           int64_t head_chunks, double* __restrict__ partials) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double (*ring)[kDepth][kSlotDoubles] = reinterpret_cast<double (*)[kDepth][kSlotDoubles]>(smem_raw);
  uint64_t (*full)[kDepth] = reinterpret_cast<uint64_t (*)[kDepth]>(smem_raw + size_t(kWarps) * kDepth * kSlotBytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0)
    for (int s = 0; s < kDepth; ++s) mbar_init(&full[warp][s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int64_t gw = int64_t(blockIdx.x) * kWarps + warp, nw = int64_t(gridDim.x) * kWarps;
  const int64_t mine = (nchunks > gw) ? (nchunks - gw + nw - 1) / nw : 0;
  uint64_t pol_last = 0, pol_first = 0;
  if (MODE == 1) {
    pol_last = policy_evict_last();
    pol_first = policy_evict_first();
  }
  auto issue = [&](int64_t k) {
    const int64_t c = gw + k * nw;
    const int s = int(k % kDepth);
    mbar_arrive_expect_tx(&full[warp][s], kSlotBytes);
    if (MODE == 1)
      bulk_g2s_hint(ring[warp][s], e + c * kSlotDoubles, kSlotBytes, &full[warp][s], c < head_chunks ? pol_last : pol_first);
    else
      bulk_g2s(ring[warp][s], e + c * kSlotDoubles, kSlotBytes, &full[warp][s]);
  };
  if (lane == 0)
    for (int64_t k = 0; k < kDepth && k < mine; ++k) issue(k);
  double acc = 0.0;
  for (int64_t k = 0; k < mine; ++k) {
    const int s = int(k % kDepth);
    mbar_wait(&full[warp][s], uint32_t(k / kDepth) & 1u);
    const double2* v = reinterpret_cast<const double2*>(ring[warp][s]);
#pragma unroll
    for (int j = 0; j < kSlotDoubles / 64; ++j) {
      const double2 x = v[j * 32 + lane];
      acc += x.x + x.y;
    }
    __syncwarp();
    if (lane == 0 && k + kDepth < mine) issue(k + kDepth);
  }
  acc = warp_sum(acc);
  if (lane == 0) partials[gw] = acc;
}

}  // namespace

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      fprintf(stderr, "%s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

int main(int argc, char** argv) {
  const int log2n = argc > 1 ? atoi(argv[1]) : 26;
  const int passes = argc > 2 ? atoi(argv[2]) : 29;
  const int64_t n = int64_t(1) << log2n;
  const int64_t nchunks = n / kSlotDoubles;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount * 2;
  double *e = nullptr, *partials = nullptr, *flush = nullptr;
  CK(cudaMalloc(&e, size_t(n) * 8));
  CK(cudaMalloc(&partials, size_t(grid) * kWarps * 8));
  const size_t flush_bytes = size_t(256) << 20;
  CK(cudaMalloc(&flush, flush_bytes));
  std::vector<double> h(1 << 20);
  for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 / double(1 + i % 977);
  for (int64_t off = 0; off < n; off += int64_t(h.size()))
    CK(cudaMemcpy(e + off, h.data(), size_t(n - off < int64_t(h.size()) ? n - off : int64_t(h.size())) * 8,
                  cudaMemcpyHostToDevice));
  const int smem = kWarps * kDepth * kSlotBytes + kWarps * kDepth * 8;
  CK(cudaFuncSetAttribute(pass_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CK(cudaFuncSetAttribute(pass_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  printf("n = 2^%d doubles (%.0f MiB), %d passes, grid %d x %d threads, L2 %d MB\n", log2n, double(n) * 8 / 1048576.0,
         passes, grid, kThreads, prop.l2CacheSize >> 20);
  const int heads_mib[] = {-1, 0, 32, 64, 80, 96};
  for (int hm : heads_mib) {
    const int64_t head_chunks = hm < 0 ? 0 : (int64_t(hm) << 20) / kSlotBytes;
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaMemset(flush, rep, flush_bytes));              // start every repetition from a flushed L2
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a));
      for (int p = 0; p < passes; ++p) {
        if (hm < 0)
          pass_kernel<0><<<grid, kThreads, smem>>>(e, nchunks, 0, partials);
        else
          pass_kernel<1><<<grid, kThreads, smem>>>(e, nchunks, head_chunks, partials);
      }
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      CK(cudaGetLastError());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (ms < best) best = ms;
    }
    const double per_pass_us = best * 1e3 / passes;
    char label[16];
    if (hm < 0)
      snprintf(label, sizeof(label), "none");
    else
      snprintf(label, sizeof(label), "%d", hm);
    printf("head %4s MiB evict_last: %8.3f ms total, %7.2f us/pass, %7.1f GB/s effective\n", label, best, per_pass_us,
           double(n) * 8 / (per_pass_us * 1e-6) / 1e9);
  }
  return 0;
}
