// bf16_debug.cu -- single-CTA probe of tcgen05.mma kind::f16 (BF16 operands) MN-major shared-memory descriptors (bring-up
// tool, not part of the library): one M = 128, N = 128, K = 16 instruction per variant; operands placed by plain stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bf16_debug tools/bf16_debug.cu && tools/bf16_debug
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

struct Variant {
  int layout_type;        // descriptor bits 61..63: 0 none, 1 SW128 base 32 B, 2 SW128, 4 SW64, 6 SW32
  int lbo, sbo;           // descriptor fields (bytes)
  int placement;          // see place()
  int lp, sp;             // strides used when placing: between MN atoms / between K groups (bytes)
};

// byte offset of element (mn, k) of a 128 (mn) x 16 (k) BF16 operand stored MN-major
__device__ __host__ inline int place(const Variant& v, int mn, int k) {
  switch (v.placement) {
    case 0:   // SW128: atom = 64 mn (128 B) x 8 k rows, 16-byte unit XOR (k & 7)
      return (mn / 64) * v.lp + (k / 8) * v.sp + (k % 8) * 128 + ((((mn % 64) / 8) ^ (k % 8)) * 16) + (mn % 8) * 2;
    case 1:   // SW128 base 32 B: atom = 64 mn (128 B) x 4 k rows, 32-byte unit XOR (k & 3)
      return (mn / 64) * v.lp + (k / 4) * v.sp + (k % 4) * 128 + ((((mn % 64) / 16) ^ (k % 4)) * 32) + (mn % 16) * 2;
    case 2:   // no swizzle: core matrix = 8 k rows x 16 B (8 mn)
      return (mn / 8) * v.lp + (k / 8) * v.sp + (k % 8) * 16 + (mn % 8) * 2;
    case 3:   // SW64: atom = 32 mn (64 B) x 8 k rows, 16-byte unit XOR ((k >> 1) & 3)
      return (mn / 32) * v.lp + (k / 8) * v.sp + (k % 8) * 64 + ((((mn % 32) / 8) ^ ((k % 8) >> 1)) * 16) + (mn % 8) * 2;
    default:  // SW32: atom = 16 mn (32 B) x 8 k rows, 16-byte unit XOR ((k >> 2) & 1)
      return (mn / 16) * v.lp + (k / 8) * v.sp + (k % 8) * 32 + ((((mn % 16) / 8) ^ ((k % 8) >> 2)) * 16) + (mn % 8) * 2;
  }
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, Variant v, unsigned int* flag) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  for (int idx = threadIdx.x; idx < 128 * 16; idx += blockDim.x) {
    const int mn = idx / 16, k = idx % 16;
    *reinterpret_cast<__nv_bfloat16*>(sA + place(v, mn, k)) = __float2bfloat16(A[idx]);
    *reinterpret_cast<__nv_bfloat16*>(sB + place(v, mn, k)) = __float2bfloat16(B[idx]);
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(slot);
  if (threadIdx.x == 0) {
    auto desc = [&](uint32_t addr) {
      return uint64_t((addr & 0x3FFFFu) >> 4) | (uint64_t(uint32_t(v.lbo) >> 4) << 16) | (uint64_t(uint32_t(v.sbo) >> 4) << 32) |
             (uint64_t(1) << 46) | (uint64_t(v.layout_type) << 61);
    };
    const uint64_t ad = desc(smem_u32(sA)), bd = desc(smem_u32(sB));
    // D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | (uint32_t(128 >> 3) << 17) |
                           (uint32_t(128 >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
        "l"(ad), "l"(bd), "r"(idesc), "r"(0)
        : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
      if (!ok && clock64() - t0 > 2000000000LL) {
        if (threadIdx.x == 0) *flag = 1;
        break;
      }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 8; ++c) {
    uint32_t r[16];
    const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16) + uint32_t(c * 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * 128 + c * 16 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128) : "memory");
}

int main() {
  static float hA[128 * 16], hB[128 * 16], hD[128 * 128], ref[128 * 128];
  srand(1);
  for (int i = 0; i < 128 * 16; ++i) {
    hA[i] = float(rand() % 9 - 4);
    hB[i] = float(rand() % 7 - 3);
  }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < 16; ++k) s += hA[m * 16 + k] * hB[n * 16 + k];
      ref[m * 128 + n] = s;
    }
  float *dA, *dB, *dD;
  unsigned int* dflag;
  cudaMalloc(&dA, sizeof(hA));
  cudaMalloc(&dB, sizeof(hB));
  cudaMalloc(&dD, sizeof(hD));
  cudaMalloc(&dflag, 4);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024);
  struct Named { const char* name; Variant v; };
  Named vs[] = {
      {"SW128      place(L=2048,S=1024) desc lbo=2048 sbo=1024", {2, 2048, 1024, 0, 2048, 1024}},
      {"SW128      place(L=2048,S=1024) desc lbo=1024 sbo=2048", {2, 1024, 2048, 0, 2048, 1024}},
      {"SW128      place(L=1024,S=2048) desc lbo=1024 sbo=2048", {2, 1024, 2048, 0, 1024, 2048}},
      {"SW128      place(L=1024,S=2048) desc lbo=2048 sbo=1024", {2, 2048, 1024, 0, 1024, 2048}},
      {"SW128_32B  place(L=2048,S=512)  desc lbo=2048 sbo=512", {1, 2048, 512, 1, 2048, 512}},
      {"SW128_32B  place(L=2048,S=512)  desc lbo=512  sbo=2048", {1, 512, 2048, 1, 2048, 512}},
      {"none       place(L=256,S=128)   desc lbo=256 sbo=128 (cores mn-adjacent... )", {0, 256, 128, 2, 256, 128}},
      {"none       place(L=256,S=128)   desc lbo=128 sbo=256", {0, 128, 256, 2, 256, 128}},
      {"none       place(L=128,S=2048)  desc lbo=128 sbo=2048", {0, 128, 2048, 2, 128, 2048}},
      {"none       place(L=128,S=2048)  desc lbo=2048 sbo=128", {0, 2048, 128, 2, 128, 2048}},
      {"SW64       place(L=1024,S=512)  desc lbo=1024 sbo=512", {4, 1024, 512, 3, 1024, 512}},
      {"SW64       place(L=1024,S=512)  desc lbo=512 sbo=1024", {4, 512, 1024, 3, 1024, 512}},
      {"SW32       place(L=512,S=256)   desc lbo=512 sbo=256", {6, 512, 256, 4, 512, 256}},
      {"SW32       place(L=512,S=256)   desc lbo=256 sbo=512", {6, 256, 512, 4, 512, 256}},
  };
  for (const Named& nv : vs) {
    cudaMemset(dD, 0xff, sizeof(hD));
    cudaMemset(dflag, 0, 4);
    probe<<<1, 128, 70 * 1024>>>(dA, dB, dD, nv.v, dflag);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-62s CUDA error: %s\n", nv.name, cudaGetErrorString(e));
      return 1;
    }
    unsigned int flag = 0;
    cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
    double maxerr = 0, maxabs = 0;
    int nz = 0, nanc = 0;
    for (int i = 0; i < 128 * 128; ++i) {
      if (hD[i] != hD[i]) { ++nanc; continue; }
      maxerr = fmax(maxerr, fabs(double(hD[i]) - ref[i]));
      maxabs = fmax(maxabs, fabs(double(hD[i])));
      if (hD[i] != 0.f) ++nz;
    }
    printf("%-62s timeout=%u maxerr=%g max|D|=%g nonzero=%d nan=%d  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n", nv.name, flag, maxerr,
           maxabs, nz, nanc, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
