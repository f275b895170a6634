// gram_tf32.cu -- pi-weighted M-step statistics of FP32-STORED samples on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM): the FP32 mode of SURVEY.md section 8(d), config C3
// (robust PCA, N = 2^24, d = 512).  Replaces, for float32 X, the contractions of
//   standard-learning/utils.py:82-84   PCA of the rows pi_i x_i      (power = 2:  G = sum pi_i^2 x_i x_i^T)
//   standard-learning/rlvi.py:70-71    sqrt(pi)-scaled least squares (power = 1:  G = sum pi_i x_i x_i^T)
//   standard-learning/utils.py:36-38,103-105                         (power = 1)
//
// Formulation.  With s_i = pi_i (power 2) or sqrt(pi_i) (power 1) and z_i = s_i x_i -- the very rows the reference
// forms -- G = Z^T Z.  The d features are cut into nb = ceil(d/128) blocks; the nb(nb+1)/2 upper-triangular
// 128 x 128 block pairs are dealt two at a time to `ngroups` groups, and CTA c (persistent, one per SM) serves
// group c % ngroups on the row tiles c / ngroups, + nslots, + 2 nslots, ... (16 rows each), so the CTAs of
// different groups sweep the same rows at the same time and X comes from HBM once and from L2 for the others.
//
// Per CTA, warp-specialised (16 warps):
//   warp 0      TMA producer: per tile and loaded 128-feature block four cp.async.bulk.tensor.2d boxes
//               (32 floats x 16 rows, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) -> exactly the one canonical MN-major
//               layout tcgen05 accepts for 32-bit operands, SWIZZLE_128B_BASE32B (Swizzle<2,5,2>: atoms of 128 B x
//               4 rows, the 32-byte unit index XORed with row mod 4), LBO = 2048 B (next 32 features), SBO = 512 B
//               (next 4 rows).  The plain SWIZZLE_128B / no-swizzle MN-major layouts silently give zeros for
//               kind::tf32 (measured with tools/tf32_debug.cu, profiles/r02_tf32_descriptor_probe.txt);
//   warps 4-7   transform (CUDA cores, in place on the landed tile): z = s_i x rounded to TF32 (cvt.rna) = Z_hi,
//               and for the 3xTF32 mode the remainder Z_lo = tf32(z - Z_hi) into a second buffer; the column
//               sums S1 = X^T pi, Sy = X^T (w y), S0, sum w y ride along; fence.proxy.async, mbarrier arrive;
//   warp 1      MMA issuer (one thread): D[128 x 128 per pair] += Z_hi[a]^T Z_hi[b] (+ Z_lo[a]^T Z_hi[b]
//               + Z_hi[a]^T Z_lo[b]), M = 128, N = 128 or 256 (two pairs sharing A), K = 8 per instruction, both
//               operands MN-major straight from shared memory; tcgen05.commit frees the stage / publishes the chunk;
//   warps 8-15  accumulators: the tensor core adds in FP32 (round-toward-zero inside the datapath), so a TMEM
//               accumulator only lives for `tiles_per_chunk` tiles (128 rows); it is then read back
//               (tcgen05.ld 32x32b) and added, round-to-nearest, into 128 FP32 registers per thread, which are
//               themselves flushed into this CTA's FP64 partial in global memory every `chunks_per_flush` chunks.
//               Two TMEM buffers (2 x 256 columns) alternate so the read-back overlaps the next chunk's MMAs.
// A finalize kernel adds the per-CTA partials in slot order (deterministic) and mirrors the upper triangle.
//
// Accuracy: 3xTF32 keeps ~2^-21 per product, the three-level accumulation keeps the sums at FP32 level over
// any N -> statistics agree with the FP64 oracle on the same float32 samples to ~1e-6 (tested at 1e-5);
// single-pass TF32 (precision = RLVI_TF32X1) rounds operands to 11 bits (zero-mean error): ~1e-3/sqrt(rows).
//
// Shapes outside the tensor path (d > 512, d % 4 != 0, unaligned X) are converted to FP64 chunk by chunk and
// handed to rlvi_weighted_moments_f64.
#include <cuda.h>
#include <math.h>

#include "tma.cuh"

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn32() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

constexpr int kMB = 128;                     // feature block edge = UMMA M
constexpr int kR = 16;                       // rows per tile (two K = 8 steps)
constexpr int kChunkBytes = kR * 128;        // one TMA box: 32 floats x 16 rows
constexpr int kBlkBytes = 4 * kChunkBytes;   // one 128-feature block of one tile: 8 KiB
constexpr int kMaxStages = 8;
constexpr int kMaxFb = 3;                    // distinct feature blocks a group can touch
constexpr int kThreads = 512;
constexpr int kXformWarp0 = 4, kEpiWarp0 = 8;
constexpr int kSmemBudget = 200 * 1024;      // stages
constexpr int kTailBytes = 8 * 1024;         // barriers, TMEM slot, reduction scratch
constexpr int kSmemTotal = kSmemBudget + kTailBytes + 1024;   // + alignment slack
constexpr int kFlushTiles = 16;              // column sums: FP32 per thread for 16 tiles, then FP64

struct Tf32Params {
  const double* w;        // pi [n]
  const double* y;        // [n] or null
  int64_t n;
  int d;
  int power;
  int nb;                 // feature blocks
  int npairs;             // nb (nb + 1) / 2
  int ngroups;            // ceil(npairs / 2)
  int nslots;             // CTAs per group; grid = ngroups * nslots
  int tiles_per_chunk;    // TMEM (level 1) accumulation length, in tiles
  int chunks_per_flush;   // register (level 2) accumulation length, in chunks
  float* gpart;           // unused (kept zero): see gpart64
  double* gpart64;        // [grid][2][128][128]
  double* spart;          // [grid][2][128]   S1, Sy of the block this group owns
  double* s0part;         // [grid][2]        S0, Swy (group 0 only)
  unsigned int* err;      // device flag: a bounded wait expired
};

__host__ __device__ __forceinline__ void pair_of(int idx, int nb, int* a, int* b) {
  int i = 0;
  while (idx >= nb - i) {
    idx -= nb - i;
    ++i;
  }
  *a = i;
  *b = i + idx;
}
__host__ __device__ __forceinline__ int pair_index(int a, int b, int nb) { return a * nb - a * (a - 1) / 2 + (b - a); }

// What one group works on: its (up to) two block pairs, the distinct feature blocks they touch (the block whose
// column sums this group owns -- the one of its diagonal pair -- first, the others ascending), and the slots of
// the pairs' operands in that list.
struct GroupPlan {
  int npairs;
  int nfb;
  int fb[kMaxFb];
  int ia[2], ib[2];
  int owned;      // 1 if fb[0] is owned (its column sums are this group's job)
  int merge;      // 1 if the two pairs share A and their B blocks are adjacent in the list: one N = 256 MMA
};

__host__ __device__ inline GroupPlan make_plan(int g, int nb, int npairs_total) {
  GroupPlan pl;
  int pa[2] = {0, 0}, pb[2] = {0, 0};
  pl.npairs = (2 * g + 1 < npairs_total) ? 2 : 1;
  for (int k = 0; k < pl.npairs; ++k) pair_of(2 * g + k, nb, &pa[k], &pb[k]);
  int own = -1;
  for (int k = 0; k < pl.npairs; ++k)
    if (pa[k] == pb[k]) own = pa[k];
  pl.nfb = 0;
  pl.owned = own >= 0 ? 1 : 0;
  if (own >= 0) pl.fb[pl.nfb++] = own;
  for (int f = 0; f < nb; ++f) {
    if (f == own) continue;
    bool used = false;
    for (int k = 0; k < pl.npairs; ++k) used = used || pa[k] == f || pb[k] == f;
    if (used) pl.fb[pl.nfb++] = f;
  }
  for (int k = pl.nfb; k < kMaxFb; ++k) pl.fb[k] = 0;
  for (int k = 0; k < 2; ++k) {
    pl.ia[k] = pl.ib[k] = 0;
    if (k >= pl.npairs) continue;
    for (int j = 0; j < pl.nfb; ++j) {
      if (pl.fb[j] == pa[k]) pl.ia[k] = j;
      if (pl.fb[j] == pb[k]) pl.ib[k] = j;
    }
  }
  pl.merge = (pl.npairs == 2 && pl.ia[0] == pl.ia[1] && pl.ib[1] == pl.ib[0] + 1) ? 1 : 0;
  return pl;
}

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d_f32(uint32_t dst_smem, const CUtensorMap* tmap, int x, int y, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          dst_smem),
      "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread (thread l = TMEM lane base + l)
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B_BASE32B: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type 1 << 61.  LBO = byte stride between 32-feature column groups, SBO = between 4-row atoms.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(kChunkBytes >> 4) << 16) | (uint64_t(512 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(1) << 61);
}
// instruction descriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), A and B MN-major (bits 15, 16),
// N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc(int n_cols) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | (uint32_t(n_cols >> 3) << 17) |
         (uint32_t(kMB >> 4) << 24);
}

// Bounded mbarrier wait: a protocol error must end in an error flag, never in a hung GPU.
__device__ __forceinline__ bool wait_or_abort(uint64_t* bar, uint32_t parity, unsigned int* err) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0u) {
      if (clock64() - t0 > 6000000000LL) {
        atomicExch(err, 1u);
        return false;
      }
      if (*reinterpret_cast<volatile unsigned int*>(err) != 0u) return false;
    }
  }
  return true;
}

template <int NSPLIT, bool HAS_Y>
__global__ void __launch_bounds__(kThreads, 1)
    gram_tf32_kernel(const __grid_constant__ CUtensorMap tmap, const Tf32Params p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  unsigned char* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const uint32_t smem_base = smem_u32(smem);
  unsigned char* tail = smem + kSmemBudget;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);          // [kMaxStages] TMA -> transform
  uint64_t* ready_bar = full_bar + kMaxStages;                      // [kMaxStages] transform -> MMA
  uint64_t* empty_bar = ready_bar + kMaxStages;                     // [kMaxStages] MMA -> TMA
  uint64_t* tfull_bar = empty_bar + kMaxStages;                     // [2] MMA -> accumulator warps
  uint64_t* tempty_bar = tfull_bar + 2;                             // [2] accumulator warps -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* red = reinterpret_cast<float*>(tail + 512);                // [4 warps][2][128] column-sum exchange
  double* dred = reinterpret_cast<double*>(tail + 512 + 4096);      // [16][2] S0 / Swy exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x % p.ngroups, slot = blockIdx.x / p.ngroups;
  const GroupPlan pl = make_plan(group, p.nb, p.npairs);
  const int stage_bytes = pl.nfb * kBlkBytes * (NSPLIT == 3 ? 2 : 1);
  int nst = kSmemBudget / stage_bytes;
  if (nst > kMaxStages) nst = kMaxStages;
  const int64_t ntiles = (p.n + kR - 1) / kR;
  const int64_t my_tiles = (ntiles > slot) ? (ntiles - slot + p.nslots - 1) / p.nslots : 0;
  const int tpc = p.tiles_per_chunk;
  const int64_t my_chunks = (my_tiles + tpc - 1) / tpc;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&ready_bar[s], 128);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 256);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: all 512 columns (two 256-column accumulator buffers); one CTA per SM
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp < kXformWarp0) {
    // ===== warpgroup 0: TMA producer (warp 0) and MMA issuer (warp 1) ==========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        if (!wait_or_abort(&empty_bar[s], ph ^ 1u, p.err)) break;
        const int row0 = int((slot + it * p.nslots) * kR);
        const uint32_t sb = smem_base + uint32_t(s) * uint32_t(stage_bytes);
        mbar_arrive_expect_tx(&full_bar[s], uint32_t(pl.nfb * kBlkBytes));
        for (int i = 0; i < pl.nfb; ++i)
#pragma unroll
          for (int c = 0; c < 4; ++c)
            tma_load_2d_f32(sb + uint32_t(i * kBlkBytes + c * kChunkBytes), &tmap, pl.fb[i] * kMB + c * 32, row0,
                            &full_bar[s]);
        if (++s == nst) {
          s = 0;
          ph ^= 1u;
        }
      }
    } else if (warp == 1 && lane == 0) {
      const uint32_t idesc128 = umma_idesc(128), idesc256 = umma_idesc(256);
      const uint32_t lo_off = uint32_t(pl.nfb * kBlkBytes);   // Z_lo blocks follow the Z_hi blocks of a stage
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      for (int64_t it = 0; it < my_tiles && ok; ++it) {
        const int64_t ch = it / tpc;
        const int tin = int(it - ch * tpc);
        const int buf = int(ch & 1);
        if (tin == 0) {
          ok = wait_or_abort(&tempty_bar[buf], (uint32_t(ch >> 1) & 1u) ^ 1u, p.err);
          if (!ok) break;
          tc_fence_after();
        }
        ok = wait_or_abort(&ready_bar[s], ph, p.err);
        if (!ok) break;
        tc_fence_after();
        const uint32_t sb = smem_base + uint32_t(s) * uint32_t(stage_bytes);
        const uint32_t dcol = tmem_base + uint32_t(buf * 256);
#pragma unroll
        for (int ks = 0; ks < kR / 8; ++ks) {
          const uint32_t first = (tin == 0 && ks == 0) ? 0u : 1u;
          if (pl.merge) {
            const uint64_t a_hi = umma_desc(sb + uint32_t(pl.ia[0] * kBlkBytes + ks * 1024));
            const uint64_t b_hi = umma_desc(sb + uint32_t(pl.ib[0] * kBlkBytes + ks * 1024));
            tc_mma_tf32(dcol, a_hi, b_hi, idesc256, first);
            if (NSPLIT == 3) {
              const uint64_t a_lo = umma_desc(sb + lo_off + uint32_t(pl.ia[0] * kBlkBytes + ks * 1024));
              const uint64_t b_lo = umma_desc(sb + lo_off + uint32_t(pl.ib[0] * kBlkBytes + ks * 1024));
              tc_mma_tf32(dcol, a_lo, b_hi, idesc256, 1u);
              tc_mma_tf32(dcol, a_hi, b_lo, idesc256, 1u);
            }
          } else {
            for (int k = 0; k < pl.npairs; ++k) {
              const uint64_t a_hi = umma_desc(sb + uint32_t(pl.ia[k] * kBlkBytes + ks * 1024));
              const uint64_t b_hi = umma_desc(sb + uint32_t(pl.ib[k] * kBlkBytes + ks * 1024));
              tc_mma_tf32(dcol + uint32_t(k * 128), a_hi, b_hi, idesc128, first);
              if (NSPLIT == 3) {
                const uint64_t a_lo = umma_desc(sb + lo_off + uint32_t(pl.ia[k] * kBlkBytes + ks * 1024));
                const uint64_t b_lo = umma_desc(sb + lo_off + uint32_t(pl.ib[k] * kBlkBytes + ks * 1024));
                tc_mma_tf32(dcol + uint32_t(k * 128), a_lo, b_hi, idesc128, 1u);
                tc_mma_tf32(dcol + uint32_t(k * 128), a_hi, b_lo, idesc128, 1u);
              }
            }
          }
        }
        tc_commit(&empty_bar[s]);                                        // stage free once these MMAs have read it
        if (tin == tpc - 1 || it == my_tiles - 1) tc_commit(&tfull_bar[buf]);   // chunk complete in TMEM
        if (++s == nst) {
          s = 0;
          ph ^= 1u;
        }
      }
    }
  } else if (warp < kEpiWarp0) {
    // ===== warpgroup 1: transform ==============================================================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
    const int tt = threadIdx.x - kXformWarp0 * 32;       // 0..127
    const int q = tt & 7, rr = tt >> 3;                   // logical 16-byte unit, row of the tile
    // SWIZZLE_128B_ATOM_32B: the 32-byte unit q >> 1 of row rr sits at unit (q >> 1) ^ (rr & 3)
    const uint32_t off = uint32_t(rr * 128 + (((((q >> 1) ^ (rr & 3)) << 1) | (q & 1)) << 4));
    const uint32_t lo_off = uint32_t(pl.nfb * kBlkBytes);
    const bool own = pl.owned != 0;
    const bool own_s0 = (group == 0) && (q == 0);
    float s1acc[4][4], syacc[4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) s1acc[c][k] = syacc[c][k] = 0.f;
    double s1d = 0.0, syd = 0.0, s0 = 0.0, swy = 0.0;

    auto flush = [&]() {
      if (own) {
        const int wq = tt >> 5;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = s1acc[c][k];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 8) red[(wq * 2 + 0) * 128 + c * 32 + lane * 4 + k] = v;
            s1acc[c][k] = 0.f;
            if (HAS_Y) {
              float u = syacc[c][k];
              u += __shfl_xor_sync(0xffffffffu, u, 8);
              u += __shfl_xor_sync(0xffffffffu, u, 16);
              if (lane < 8) red[(wq * 2 + 1) * 128 + c * 32 + lane * 4 + k] = u;
              syacc[c][k] = 0.f;
            }
          }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (own) {
        s1d += double((red[0 * 128 + tt] + red[2 * 128 + tt]) + (red[4 * 128 + tt] + red[6 * 128 + tt]));
        if (HAS_Y) syd += double((red[1 * 128 + tt] + red[3 * 128 + tt]) + (red[5 * 128 + tt] + red[7 * 128 + tt]));
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    };

    auto load_row = [&](int64_t it, double& pid, double& yd) {
      const int64_t row = (slot + it * p.nslots) * kR + rr;
      pid = (it < my_tiles && row < p.n) ? p.w[row] : 0.0;
      yd = (HAS_Y && it < my_tiles && row < p.n) ? p.y[row] : 0.0;
    };

    double pid, yd;
    load_row(0, pid, yd);
    int s = 0;
    uint32_t ph = 0;
    bool ok = true;
    for (int64_t it = 0; it < my_tiles; ++it) {
      double pid_n, yd_n;
      load_row(it + 1, pid_n, yd_n);                       // prefetch the next tile's row coefficients
      const double wd = (p.power == 2) ? pid * pid : pid;
      const float sc = (p.power == 2) ? float(pid) : float(sqrt(pid));
      const float c1 = float(pid);
      const float cy = HAS_Y ? float(wd * yd) : 0.f;
      if (own_s0) {
        s0 += wd;
        if (HAS_Y) swy = fma(wd, yd, swy);
      }
      ok = wait_or_abort(&full_bar[s], ph, p.err);
      if (!ok) break;
      unsigned char* sb = smem + size_t(s) * stage_bytes;
#pragma unroll
      for (int i = 0; i < kMaxFb; ++i) {
        if (i < pl.nfb) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            float4* ptr = reinterpret_cast<float4*>(sb + i * kBlkBytes + c * kChunkBytes + off);
            const float4 x = *ptr;
            if (i == 0 && own) {
              s1acc[c][0] = fmaf(c1, x.x, s1acc[c][0]);
              s1acc[c][1] = fmaf(c1, x.y, s1acc[c][1]);
              s1acc[c][2] = fmaf(c1, x.z, s1acc[c][2]);
              s1acc[c][3] = fmaf(c1, x.w, s1acc[c][3]);
              if (HAS_Y) {
                syacc[c][0] = fmaf(cy, x.x, syacc[c][0]);
                syacc[c][1] = fmaf(cy, x.y, syacc[c][1]);
                syacc[c][2] = fmaf(cy, x.z, syacc[c][2]);
                syacc[c][3] = fmaf(cy, x.w, syacc[c][3]);
              }
            }
            const float z0 = sc * x.x, z1 = sc * x.y, z2 = sc * x.z, z3 = sc * x.w;
            const uint32_t h0 = to_tf32(z0), h1 = to_tf32(z1), h2 = to_tf32(z2), h3 = to_tf32(z3);
            *reinterpret_cast<uint4*>(ptr) = make_uint4(h0, h1, h2, h3);
            if (NSPLIT == 3) {
              const uint32_t l0 = to_tf32(z0 - __uint_as_float(h0)), l1 = to_tf32(z1 - __uint_as_float(h1)),
                             l2 = to_tf32(z2 - __uint_as_float(h2)), l3 = to_tf32(z3 - __uint_as_float(h3));
              *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(ptr) + lo_off) = make_uint4(l0, l1, l2, l3);
            }
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the tensor core
      mbar_arrive(&ready_bar[s]);
      if (((it + 1) % kFlushTiles) == 0) flush();
      pid = pid_n;
      yd = yd_n;
      if (++s == nst) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (ok) {
      flush();
      if (own) {
        p.spart[(size_t(blockIdx.x) * 2 + 0) * 128 + tt] = s1d;
        p.spart[(size_t(blockIdx.x) * 2 + 1) * 128 + tt] = syd;
      }
      if (group == 0) {
        if (q == 0) {
          dred[rr * 2 + 0] = s0;
          dred[rr * 2 + 1] = swy;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tt == 0) {
          double a = 0.0, b = 0.0;
          for (int r = 0; r < kR; ++r) {
            a += dred[r * 2 + 0];
            b += dred[r * 2 + 1];
          }
          p.s0part[size_t(blockIdx.x) * 2 + 0] = a;
          p.s0part[size_t(blockIdx.x) * 2 + 1] = b;
        }
      }
    }
  } else {
    // ===== warpgroups 2, 3: accumulators (TMEM -> FP32 registers -> FP64 partial) ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    const int e = warp - kEpiWarp0;                // 0..7
    const int quarter = warp & 3;                  // TMEM lanes 32 quarter .. + 31 (hardware: warp id % 4)
    const int h = e >> 2;                          // which pair of the group (columns h * 128 ..)
    const bool active = h < pl.npairs;
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    double* gp = p.gpart64 + ((size_t(blockIdx.x) * 2 + h) * 128 + size_t(quarter * 32 + lane)) * 128;
    bool flushed = false;
    bool ok = true;

    auto flush = [&]() {
      if (active) {
#pragma unroll
        for (int i = 0; i < 128; i += 2) {
          double2 v = make_double2(double(acc[i]), double(acc[i + 1]));
          if (flushed) {
            const double2 o = *reinterpret_cast<const double2*>(gp + i);
            v.x += o.x;
            v.y += o.y;
          }
          *reinterpret_cast<double2*>(gp + i) = v;
          acc[i] = 0.f;
          acc[i + 1] = 0.f;
        }
      }
      flushed = true;
    };

    for (int64_t ch = 0; ch < my_chunks; ++ch) {
      const int buf = int(ch & 1);
      ok = wait_or_abort(&tfull_bar[buf], uint32_t(ch >> 1) & 1u, p.err);
      if (!ok) break;
      tc_fence_after();
      if (active) {
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(buf * 256 + h * 128);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float v[16];
          tc_ld16(taddr + uint32_t(c * 16), v);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[c * 16 + j] += v[j];
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
      if (((ch + 1) % p.chunks_per_flush) == 0) flush();
    }
    if (ok) flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// Sum the per-CTA partials in slot order, mirror the upper triangle, write [S0, Swy, S1, Sy, G] (FP64).
__global__ void __launch_bounds__(256) gram_tf32_finalize_kernel(const Tf32Params p, int has_y, int want_gram, double* out) {
  const int d = p.d;
  const int64_t gtotal = int64_t(p.npairs) * kMB * kMB;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool bad = *p.err != 0u;
  if (idx < gtotal) {
    if (!want_gram) return;
    const int pidx = int(idx / (kMB * kMB));
    const int i = int((idx / kMB) % kMB), j = int(idx % kMB);
    int a, b;
    pair_of(pidx, p.nb, &a, &b);
    const int fi = a * kMB + i, fj = b * kMB + j;
    if (fi >= d || fj >= d) return;
    if (a == b && i > j) return;
    const int g = pidx >> 1, h = pidx & 1;
    double s = 0.0;
    for (int sl = 0; sl < p.nslots; ++sl)
      s += p.gpart64[((size_t(sl) * p.ngroups + g) * 2 + h) * (kMB * kMB) + size_t(i) * kMB + j];
    if (bad) s = nan("");
    double* G = out + 2 + 2 * d;
    G[size_t(fi) * d + fj] = s;
    G[size_t(fj) * d + fi] = s;
    return;
  }
  const int64_t k = idx - gtotal;
  if (k < 2) {          // S0, Swy: group 0
    double s = 0.0;
    for (int sl = 0; sl < p.nslots; ++sl) s += p.s0part[(size_t(sl) * p.ngroups) * 2 + k];
    out[k] = (bad || (k == 1 && !has_y)) ? (bad ? nan("") : 0.0) : s;
    return;
  }
  const int64_t f2 = k - 2;
  if (f2 < 2 * int64_t(d)) {
    const int which = int(f2 / d), f = int(f2 % d);
    const int blk = f / kMB, fin = f % kMB;
    const int g = pair_index(blk, blk, p.nb) >> 1;
    double s = 0.0;
    if (which == 0 || has_y)
      for (int sl = 0; sl < p.nslots; ++sl) s += p.spart[((size_t(sl) * p.ngroups + g) * 2 + which) * 128 + fin];
    out[2 + which * d + f] = bad ? nan("") : s;
  }
}

__global__ void __launch_bounds__(256) f32_to_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t count) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x)
    dst[i] = double(src[i]);
}
__global__ void __launch_bounds__(256) add_f64_kernel(double* __restrict__ dst, const double* __restrict__ src, int count, int first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = first ? src[i] : dst[i] + src[i];
}

}  // namespace

// Shapes the tensor path does not cover: convert row chunks to FP64 and reuse the FP64 statistics kernels.
static int moments_f32_via_f64(rlvi_ctx* ctx, const float* X, const double* y, const double* weights, int64_t n, int d,
                               int power, int want_gram, double* out, cudaStream_t st) {
  const int nm = rlvi_moments_out_doubles(d);
  int64_t chunk = (int64_t(64) << 20) / (int64_t(d) * 8);   // 64 MiB of converted rows at a time
  if (chunk < 64) chunk = 64;
  if (chunk > n) chunk = n;
  const size_t need = size_t(chunk) * d * 8 + size_t(nm) * 8 + 256;
  if (need > ctx->big_bytes) {
    RLVI_CUDA(cudaDeviceSynchronize());
    if (ctx->big) cudaFree(ctx->big);
    ctx->big = nullptr;
    ctx->big_bytes = 0;
    if (cudaMalloc(&ctx->big, need) != cudaSuccess) {
      rlvi_set_error("cudaMalloc of %zu bytes for the FP32 -> FP64 staging buffer failed", need);
      return RLVI_ERR_NOMEM;
    }
    ctx->big_bytes = need;
  }
  double* xd = static_cast<double*>(ctx->big);
  double* tmp = xd + size_t(chunk) * d;
  bool first = true;
  for (int64_t r0 = 0; r0 < n; r0 += chunk) {
    const int64_t rows = (n - r0 < chunk) ? (n - r0) : chunk;
    f32_to_f64_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(X + r0 * d, xd, rows * d);
    RLVI_LAUNCH_CHECK(ctx);
    RLVI_CUDA(cudaMemsetAsync(tmp, 0, size_t(nm) * 8, st));
    const int rc = rlvi_weighted_moments_f64(ctx, xd, y ? y + r0 : nullptr, weights + r0, rows, d, power, want_gram, tmp, st);
    if (rc != RLVI_OK) return rc;
    add_f64_kernel<<<(nm + 255) / 256, 256, 0, st>>>(out, tmp, nm, first ? 1 : 0);
    RLVI_LAUNCH_CHECK(ctx);
    first = false;
  }
  return RLVI_OK;
}

extern "C" int rlvi_weighted_moments_f32(rlvi_ctx* ctx, const float* X, const double* y, const double* weights, int64_t n,
                                         int d, int power, int want_gram, int precision, double* out, void* stream) {
  RLVI_REQUIRE(ctx && X && weights && out, "null pointer");
  RLVI_REQUIRE(n > 0 && d > 0, "n and d must be positive");
  RLVI_REQUIRE(power == 1 || power == 2, "power must be 1 or 2");
  RLVI_REQUIRE(precision == RLVI_TF32X3 || precision == RLVI_TF32X1, "precision must be RLVI_TF32X3 or RLVI_TF32X1");
  if (d > 1024) {
    rlvi_set_error("rlvi_weighted_moments_f32: d = %d > 1024 is not supported", d);
    return RLVI_ERR_UNSUPPORTED;
  }
  RlviDeviceGuard guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EncodeTiledFn encode = encode_tiled_fn32();
  if (!encode || d > 4 * kMB || d % 4 != 0 || n >= (int64_t(1) << 31) - kR || !rlvi_aligned16(X))
    return moments_f32_via_f64(ctx, X, y, weights, n, d, power, want_gram, out, st);

  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {cuuint64_t(d), cuuint64_t(n)};
  const cuuint64_t gstride[1] = {cuuint64_t(d) * 4};
  const cuuint32_t box[2] = {32, cuuint32_t(kR)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(X), gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    rlvi_set_error("cuTensorMapEncodeTiled (FP32) failed (%d)", int(cr));
    return RLVI_ERR_CUDA;
  }
  Tf32Params p;
  memset(&p, 0, sizeof(p));
  p.w = weights;
  p.y = y;
  p.n = n;
  p.d = d;
  p.power = power;
  p.nb = (d + kMB - 1) / kMB;
  p.npairs = p.nb * (p.nb + 1) / 2;
  p.ngroups = (p.npairs + 1) / 2;
  const int64_t ntiles = (n + kR - 1) / kR;
  int nslots = ctx->sm_count / p.ngroups;
  if (nslots > ntiles) nslots = int(ntiles);
  if (nslots < 1) nslots = 1;
  p.nslots = nslots;
  p.tiles_per_chunk = 8;        // 128 rows per TMEM accumulation
  p.chunks_per_flush = 256;     // 32 Ki rows per register accumulation
  const int grid = p.ngroups * p.nslots;
  const size_t gbytes = size_t(grid) * 2 * kMB * kMB * sizeof(double);
  const size_t sbytes = size_t(grid) * 2 * 128 * sizeof(double);
  const size_t s0bytes = size_t(grid) * 2 * sizeof(double);
  void* scratch = nullptr;
  const int rc = rlvi_scratch(ctx, 4096 + gbytes + sbytes + s0bytes, &scratch);
  if (rc != RLVI_OK) return rc;
  char* base = static_cast<char*>(scratch);
  p.err = reinterpret_cast<unsigned int*>(base + 2048);
  p.gpart64 = reinterpret_cast<double*>(base + 4096);
  p.spart = reinterpret_cast<double*>(base + 4096 + gbytes);
  p.s0part = reinterpret_cast<double*>(base + 4096 + gbytes + sbytes);
  RLVI_CUDA(cudaMemsetAsync(p.err, 0, 4, st));

#define RLVI_TF32_LAUNCH(NS, HY)                                                                                   \
  {                                                                                                                \
    RLVI_CUDA(cudaFuncSetAttribute(gram_tf32_kernel<NS, HY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal)); \
    gram_tf32_kernel<NS, HY><<<grid, kThreads, kSmemTotal, st>>>(tmap, p);                                         \
  }
  if (precision == RLVI_TF32X1) {
    if (y) RLVI_TF32_LAUNCH(1, true) else RLVI_TF32_LAUNCH(1, false)
  } else {
    if (y) RLVI_TF32_LAUNCH(3, true) else RLVI_TF32_LAUNCH(3, false)
  }
#undef RLVI_TF32_LAUNCH
  RLVI_LAUNCH_CHECK(ctx);
  const int64_t total = int64_t(p.npairs) * kMB * kMB + 2 + 2 * int64_t(d);
  gram_tf32_finalize_kernel<<<int((total + 255) / 256), 256, 0, st>>>(p, y ? 1 : 0, want_gram, out);
  RLVI_LAUNCH_CHECK(ctx);
  return RLVI_OK;
}
