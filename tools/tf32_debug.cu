// tf32_debug.cu -- single-CTA probe of tcgen05.mma kind::tf32 operand descriptors (bring-up tool, not part of the
// library): one M = 128, N = 128, K = 8 instruction per variant, operands placed in shared memory by plain stores
// according to the variant's layout function, result read back with tcgen05.ld and compared with the host.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tf32_debug tools/tf32_debug.cu && tools/tf32_debug
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

struct Variant {
  int layout_type;   // descriptor bits 61..63
  int lbo, sbo;      // bytes
  int a_major, b_major;   // idesc bits 15, 16
  int placement;     // 0: MN-major SW128, 1: K-major no swizzle, 2: MN-major no swizzle, 3: K-major SW128 (K = 8 -> 32 B rows)
  int use_mask_form; // 1: the 4-register disable-output-lane form of the instruction
  int lbo_place, sbo_place;   // strides used when PLACING the data (bytes)
};

// byte offset of element (mn, k) of a 128 x 8 operand for the variant's placement
__device__ __host__ inline int place(const Variant& v, int mn, int k) {
  switch (v.placement) {
    case 0:   // MN-major, 128B swizzle: atom = 32 mn (128 B) x 8 k rows; 16-byte unit index XOR (k & 7)
      return (mn / 32) * v.lbo_place + (k / 8) * v.sbo_place + (k % 8) * 128 + ((((mn % 32) / 4) ^ (k % 8)) * 16) + (mn % 4) * 4;
    case 1:   // K-major, no swizzle: core matrix = 8 mn rows x 16 B (4 k); LBO between core matrices along K, SBO along MN
      return (k / 4) * v.lbo_place + (mn / 8) * v.sbo_place + (mn % 8) * 16 + (k % 4) * 4;
    case 2:   // MN-major, no swizzle: core matrix = 8 k rows x 16 B (4 mn)
      return (mn / 4) * v.sbo_place + (k / 8) * v.lbo_place + (k % 8) * 16 + (mn % 4) * 4;
    case 4:   // MN-major, 128B swizzle with 32-byte atoms (Swizzle<2,5,2>): atom = 32 mn (128 B) x 4 k rows
      return (mn / 32) * v.lbo_place + (k / 4) * v.sbo_place + (k % 4) * 128 + ((((mn % 32) / 8) ^ (k % 4)) * 32) + (mn % 8) * 4;
    default:  // K-major, 128B swizzle with 32-byte rows?  rows of 128 B hold 32 k; only k < 8 used
      return (mn / 8) * v.sbo_place + (mn % 8) * 128 + ((((k % 32) / 4) ^ (mn % 8)) * 16) + (k % 4) * 4;
  }
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, Variant v, unsigned int* flag,
                                                const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int use_tma) {
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
  if (!use_tma) {
    for (int idx = threadIdx.x; idx < 128 * 8; idx += blockDim.x) {
      const int mn = idx / 8, k = idx % 8;
      *reinterpret_cast<float*>(sA + place(v, mn, k)) = A[idx];
      *reinterpret_cast<float*>(sB + place(v, mn, k)) = B[idx];
    }
  }
  uint64_t* tbar = bar + 1;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(tbar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (use_tma) {
    // global operands are [k = 8][mn = 128] row-major (MN-major); four boxes of 32 floats x 8 rows per operand
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(tbar)), "r"(2 * 4096) : "memory");
      for (int c = 0; c < 4; ++c) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(sA + c * v.lbo_place)), "l"(&mapA), "r"(c * 32), "r"(0), "r"(smem_u32(tbar)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(sB + c * v.lbo_place)), "l"(&mapB), "r"(c * 32), "r"(0), "r"(smem_u32(tbar)) : "memory");
      }
    }
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(tbar)), "r"(0) : "memory");
      if (!ok && clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) *flag = 2; break; }
    }
  }
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
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(v.a_major) << 15) | (uint32_t(v.b_major) << 16) |
                           (uint32_t(128 >> 3) << 17) | (uint32_t(128 >> 4) << 24);
    if (v.use_mask_form) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(idesc), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    } else {
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(ad), "l"(bd), "r"(idesc), "r"(0)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // bounded wait
  {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(bar)), "r"(0)
          : "memory");
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
  float hA[128 * 8], hB[128 * 8], hD[128 * 128], ref[128 * 128];
  srand(1);
  for (int i = 0; i < 128 * 8; ++i) {
    hA[i] = float(rand() % 9 - 4);
    hB[i] = float(rand() % 7 - 3);
  }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      float s = 0;
      for (int k = 0; k < 8; ++k) s += hA[m * 8 + k] * hB[n * 8 + k];
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
  // transposed copies [k][mn] for the TMA variants
  float hAt[8 * 128], hBt[8 * 128];
  for (int mn = 0; mn < 128; ++mn)
    for (int k = 0; k < 8; ++k) {
      hAt[k * 128 + mn] = hA[mn * 8 + k];
      hBt[k * 128 + mn] = hB[mn * 8 + k];
    }
  float *dAt, *dBt;
  cudaMalloc(&dAt, sizeof(hAt));
  cudaMalloc(&dBt, sizeof(hBt));
  cudaMemcpy(dAt, hAt, sizeof(hAt), cudaMemcpyHostToDevice);
  cudaMemcpy(dBt, hBt, sizeof(hBt), cudaMemcpyHostToDevice);
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
  EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fp);
  auto make_map = [&](float* ptr, CUtensorMapSwizzle sw) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {128, 8};
    const cuuint64_t gstride[1] = {128 * 4};
    const cuuint32_t box[2] = {32, 8};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) printf("encode failed %d\n", int(cr));
    return m;
  };
  CUtensorMap mapA32 = make_map(dAt, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), mapB32 = make_map(dBt, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  struct Named { const char* name; Variant v; int tma; };
  Named vs[] = {
      {"K-major none  lbo=128 sbo=256 (core matrices: K then MN)", {0, 128, 256, 0, 0, 1, 0, 128, 256}, 0},
      {"K-major none  mask-form", {0, 128, 256, 0, 0, 1, 1, 128, 256}, 0},
      {"K-major none  lbo/sbo swapped in descriptor", {0, 256, 128, 0, 0, 1, 0, 128, 256}, 0},
      {"MN-major SW128 lbo=2048(chunk) sbo=1024 [library]", {2, 2048, 1024, 1, 1, 0, 0, 2048, 1024}, 0},
      {"MN-major SW128 descriptor lbo/sbo swapped", {2, 1024, 2048, 1, 1, 0, 0, 2048, 1024}, 0},
      {"MN-major SW128 lbo=1024 (dense chunks) sbo=4096", {2, 1024, 4096, 1, 1, 0, 0, 1024, 4096}, 0},
      {"MN-major SW128 dense, descriptor swapped", {2, 4096, 1024, 1, 1, 0, 0, 1024, 4096}, 0},
      {"MN-major none  lbo=128 sbo=... (8k x 4mn cores, mn-adjacent)", {0, 4096, 128, 1, 1, 2, 0, 4096, 128}, 0},
      {"MN-major none  descriptor swapped", {0, 128, 4096, 1, 1, 2, 0, 4096, 128}, 0},
      {"K-major SW128 32B rows sbo=1024", {2, 16, 1024, 0, 0, 3, 0, 16, 1024}, 0},
      {"MN-major SW128_32B lbo=2048 sbo=512 [new library layout]", {1, 2048, 512, 1, 1, 4, 0, 2048, 512}, 0},
      {"MN-major SW128_32B descriptor lbo/sbo swapped", {1, 512, 2048, 1, 1, 4, 0, 2048, 512}, 0},
      {"MN-major SW128_32B lbo=1024 (dense) sbo=512", {1, 1024, 512, 1, 1, 4, 0, 1024, 512}, 0},
      {"MN-major SW128_32B dense, descriptor swapped", {1, 512, 1024, 1, 1, 4, 0, 1024, 512}, 0},
      {"MN-major SW128_32B via TMA (SWIZZLE_128B_ATOM_32B) lbo=2048 sbo=512", {1, 2048, 512, 1, 1, 4, 0, 2048, 512}, 1},
      {"MN-major SW128_32B via TMA, A MN-major x B MN-major, lbo=1024", {1, 1024, 512, 1, 1, 4, 0, 1024, 512}, 1},
  };
  for (const Named& nv : vs) {
    cudaMemset(dD, 0xff, sizeof(hD));
    cudaMemset(dflag, 0, 4);
    probe<<<1, 128, 70 * 1024>>>(dA, dB, dD, nv.v, dflag, mapA32, mapB32, nv.tma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-70s CUDA error: %s\n", nv.name, cudaGetErrorString(e));
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
    printf("%-70s timeout=%u maxerr=%g max|D|=%g nonzero=%d nan=%d  D[0][0..3]=%g %g %g %g ref=%g %g %g %g\n", nv.name, flag,
           maxerr, maxabs, nz, nanc, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
  }
  return 0;
}
